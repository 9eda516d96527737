#!/usr/bin/env python
"""Symbolic execution of a straight-line scalar/packed SSE float32 block of an objdump listing.

Development aid used once to pin the operation ORDER of the float32 4-point solver inside the installed cv2's RHO
estimator (third-party binary, no source in the image): every arithmetic instruction becomes one SSA statement
`tN = a op b` in float32, so the restatement in oracle/rho.py rounds exactly where the library rounds.  Not used
by tests or the product.

    python symexec.py listing.asm START_ADDR END_ADDR cv2.so > out.py
"""
import re, struct, sys

lst, start, end, so = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16), sys.argv[4]
blob = open(so, 'rb')

def rodata_f32(addr):
    blob.seek(addr); return struct.unpack('<I', blob.read(4))[0]

nodes = []  # (op, a, b)
def new(op, a=None, b=None):
    nodes.append((op, a, b)); return len(nodes) - 1
ZERO = new('const', 0)
xmm = {i: [ZERO] * 4 for i in range(16)}
gpr = {}
mem = {}   # (base, off) -> node (4-byte slots)
base_name = {'rax': 'P'}   # rax = pkdPts until reloaded
inputs = {}
def load(base, off):
    key = (base, off)
    if key not in mem:
        if base == 'rsp':
            raise SystemExit(f'uninitialised stack slot {off:#x}')
        mem[key] = new('in', base, off); inputs[key] = mem[key]
    return mem[key]

def parse_mem(s):
    m = re.fullmatch(r'(-?0x[0-9a-f]+)?\(%(\w+)\)', s)
    if not m: return None
    off = int(m.group(1), 16) if m.group(1) else 0
    reg = m.group(2)
    return (base_name.get(reg, reg), off)

outs = {}
for r, o in ((5, 0), (3, 8), (0, 16), (10, 24)):   # x0..x3 are loaded before the block starts
    xmm[r] = [load('P', o), ZERO, ZERO, ZERO]
for line in open(lst):
    m = re.match(r'\s*([0-9a-f]+):\s+(\w+)\s*(.*)', line)
    if not m: continue
    addr = int(m.group(1), 16)
    if addr < start or addr > end: continue
    op, args = m.group(2), m.group(3).split('#')[0].strip()
    a = [x.strip() for x in re.split(r',(?![^(]*\))', args)] if args else []
    def X(s): return int(s[4:]) if s.startswith('%xmm') else None
    if op in ('js', 'jp', 'je', 'jne', 'cvttss2si', 'xor', 'ucomiss', 'addl', 'movl', 'setnp', 'cmovne', 'cmove', 'or', 'test'):
        continue
    if op == 'mov':
        if a[0] == '0xa8(%rbx)' and a[1] == '%rax': base_name['rax'] = 'H'
        continue
    rip = re.match(r'(0x[0-9a-f]+)\(%rip\)', a[0]) if a else None
    if op == 'movss':
        s, d = a
        if X(s) is not None and X(d) is not None: xmm[X(d)][0] = xmm[X(s)][0]
        elif X(d) is not None:
            if rip:
                tgt = int(line.split('#')[1].split()[0], 16); v = new('const', rodata_f32(tgt))
            else:
                v = load(*parse_mem(s))
            xmm[X(d)] = [v, ZERO, ZERO, ZERO]
        else:
            k = parse_mem(d)
            mem[k] = xmm[X(s)][0]
            if k[0] == 'H': outs[k[1]] = xmm[X(s)][0]
    elif op == 'movaps': xmm[X(a[1])] = list(xmm[X(a[0])])
    elif op == 'movd':
        s, d = a
        if X(s) is not None: gpr[d] = xmm[X(s)][0]
        else: xmm[X(d)] = [gpr[s], ZERO, ZERO, ZERO]
    elif op == 'movq': xmm[X(a[1])] = [xmm[X(a[0])][0], xmm[X(a[0])][1], ZERO, ZERO]
    elif op in ('mulss', 'subss', 'addss', 'divss'):
        s, d = a
        sv = xmm[X(s)][0] if X(s) is not None else load(*parse_mem(s))
        xmm[X(d)][0] = new(op[:3], xmm[X(d)][0], sv)
    elif op in ('mulps', 'subps'):
        s, d = a
        xmm[X(d)] = [new(op[:3], xmm[X(d)][i], xmm[X(s)][i]) for i in range(4)]
    elif op == 'xorps':
        s, d = a
        xmm[X(d)] = [new('xor', xmm[X(d)][i], xmm[X(s)][i]) for i in range(4)]
    elif op == 'unpcklps':
        s, d = a; S, D = xmm[X(s)], xmm[X(d)]; xmm[X(d)] = [D[0], S[0], D[1], S[1]]
    elif op == 'unpckhps':
        s, d = a; S, D = xmm[X(s)], xmm[X(d)]; xmm[X(d)] = [D[2], S[2], D[3], S[3]]
    elif op == 'movlhps':
        s, d = a; S, D = xmm[X(s)], xmm[X(d)]; xmm[X(d)] = [D[0], D[1], S[0], S[1]]
    elif op == 'movsldup':
        s, d = a; S = xmm[X(s)]; xmm[X(d)] = [S[0], S[0], S[2], S[2]]
    elif op == 'movshdup':
        s, d = a; S = xmm[X(s)]; xmm[X(d)] = [S[1], S[1], S[3], S[3]]
    elif op == 'shufps':
        imm, s, d = a; imm = int(imm[1:], 16); S, D = xmm[X(s)], xmm[X(d)]
        xmm[X(d)] = [D[imm & 3], D[(imm >> 2) & 3], S[(imm >> 4) & 3], S[(imm >> 6) & 3]]
    elif op == 'movlps':
        s, d = a; k = parse_mem(d)
        for i in range(2):
            mem[(k[0], k[1] + 4 * i)] = xmm[X(s)][i]
            if k[0] == 'H': outs[k[1] + 4 * i] = xmm[X(s)][i]
    elif op == 'movups':
        s, d = a; k = parse_mem(d)
        for i in range(4):
            mem[(k[0], k[1] + 4 * i)] = xmm[X(s)][i]
            if k[0] == 'H': outs[k[1] + 4 * i] = xmm[X(s)][i]
    else:
        raise SystemExit(f'unhandled {line}')

# emit reachable nodes
need = set()
def mark(n):
    if n in need or n is None: return
    need.add(n); op, a, b = nodes[n]
    if op in ('mul', 'sub', 'add', 'div', 'xor'): mark(a); mark(b)
for v in outs.values(): mark(v)
sym = {'mul': '*', 'sub': '-', 'add': '+', 'div': '/'}
print('# generated by tools/cv2_probe/symexec.py; P[0..7] = x0,y0..x3,y3 (src), P[8..15] = X0,Y0..X3,Y3 (dst)')
for n in sorted(need):
    op, a, b = nodes[n]
    if op == 'const': print(f't{n} = CONST({a:#010x})')
    elif op == 'in': print(f't{n} = P[{b // 4}]')
    elif op == 'xor': print(f't{n} = XOR(t{a}, t{b})')
    else: print(f't{n} = t{a} {sym[op]} t{b}')
for off in sorted(outs): print(f'H[{off // 4}] = t{outs[off]}')
