#!/usr/bin/env python
"""Generic variant of symexec.py: any memory operand (base register + offset) read before it is written is an input
`M_<base>_<off>`, any xmm register read before written is an input `X<n>`; every store is recorded.  Emits SSA float32
statements for the values named on the command line (or every store outside the stack).

    python symexec2.py listing.asm START END cv2.so [want ...]      want = rcx:0x20  |  rsp:0x98  |  xmm3
"""
import re, struct, sys

lst, start, end, so = sys.argv[1], int(sys.argv[2], 16), int(sys.argv[3], 16), sys.argv[4]
wants = sys.argv[5:]
blob = open(so, 'rb')
def rodata_u32(addr):
    blob.seek(addr); return struct.unpack('<I', blob.read(4))[0]
nodes = []
def new(op, a=None, b=None):
    nodes.append((op, a, b)); return len(nodes) - 1
ZERO = new('const', 0)
xmm = {}
gpr = {}
mem = {}
stores = {}
def rx(n, lane=None):
    if n not in xmm:
        xmm[n] = [new('in', f'X{n}_{i}') for i in range(4)]
    return xmm[n] if lane is None else xmm[n][lane]
def parse_mem(s):
    m = re.fullmatch(r'(-?0x[0-9a-f]+)?\(%(\w+)\)', s)
    if not m: return None
    return (m.group(2), int(m.group(1), 16) if m.group(1) else 0)
def load(k):
    if k not in mem:
        mem[k] = new('in', f'M_{k[0]}_{k[1] & 0xffffffff:x}' if k[1] >= 0 else f'M_{k[0]}_m{-k[1]:x}')
    return mem[k]
def store(k, v):
    mem[k] = v; stores[k] = v
def X(s): return int(s[4:]) if s.startswith('%xmm') else None
SKIP = {'js','jp','je','jne','ja','jbe','jae','jb','cvttss2si','xor','ucomiss','comiss','addl','movl','setnp','cmovne','cmove','or','test','mov','cmp','nop','nopl','nopw','lea','add','sub'}
for line in open(lst):
    m = re.match(r'\s*([0-9a-f]+):\s+(\w+)\s*(.*)', line)
    if not m: continue
    addr = int(m.group(1), 16)
    if addr < start or addr > end: continue
    op, args = m.group(2), m.group(3).split('#')[0].strip()
    a = [x.strip() for x in re.split(r',(?![^(]*\))', args)] if args else []
    if op in SKIP: continue
    rip = re.match(r'(0x[0-9a-f]+)\(%rip\)', a[0]) if a else None
    def src_scalar(s):
        if X(s) is not None: return rx(X(s), 0)
        if rip: return new('const', rodata_u32(int(line.split('#')[1].split()[0], 16)))
        return load(parse_mem(s))
    if op == 'movss':
        s, d = a
        if X(s) is not None and X(d) is not None: rx(X(d))[0] = rx(X(s), 0)
        elif X(d) is not None: xmm[X(d)] = [src_scalar(s), ZERO, ZERO, ZERO]
        else: store(parse_mem(d), rx(X(s), 0))
    elif op in ('movaps', 'movdqa', 'movups', 'movdqu'):
        s, d = a
        if X(s) is not None and X(d) is not None: xmm[X(d)] = list(rx(X(s)))
        elif X(d) is not None:
            k = parse_mem(s); xmm[X(d)] = [load((k[0], k[1] + 4 * i)) for i in range(4)]
        else:
            k = parse_mem(d)
            for i in range(4): store((k[0], k[1] + 4 * i), rx(X(s), i))
    elif op == 'movd':
        s, d = a
        if X(s) is not None: gpr[d] = rx(X(s), 0)
        else: xmm[X(d)] = [gpr[s], ZERO, ZERO, ZERO]
    elif op == 'movq':
        s, d = a
        if X(s) is not None and X(d) is not None: xmm[X(d)] = [rx(X(s), 0), rx(X(s), 1), ZERO, ZERO]
        elif rip:
            t = int(line.split('#')[1].split()[0], 16)
            xmm[X(d)] = [new('const', rodata_u32(t)), new('const', rodata_u32(t + 4)), ZERO, ZERO]
        elif X(d) is not None:
            k = parse_mem(s); xmm[X(d)] = [load(k), load((k[0], k[1] + 4)), ZERO, ZERO]
        else:
            k = parse_mem(d)
            for i in range(2): store((k[0], k[1] + 4 * i), rx(X(s), i))
    elif op == 'pxor' and a[0] == a[1]: xmm[X(a[0])] = [ZERO] * 4
    elif op in ('mulss', 'subss', 'addss', 'divss'):
        s, d = a
        sv = src_scalar(s)
        rx(X(d))[0] = new(op[:3], rx(X(d), 0), sv)
    elif op == 'sqrtss':
        s, d = a; rx(X(d))[0] = new('sqrt', src_scalar(s), None)
    elif op in ('mulps', 'subps', 'addps', 'divps'):
        s, d = a
        S = rx(X(s)) if X(s) is not None else [load((parse_mem(s)[0], parse_mem(s)[1] + 4 * i)) for i in range(4)]
        xmm[X(d)] = [new(op[:3], rx(X(d), i), S[i]) for i in range(4)]
    elif op in ('xorps', 'andps'):
        s, d = a
        S = rx(X(s)) if X(s) is not None else [load((parse_mem(s)[0], parse_mem(s)[1] + 4 * i)) for i in range(4)]
        xmm[X(d)] = [new(op[:3], rx(X(d), i), S[i]) for i in range(4)]
    elif op == 'unpcklps':
        s, d = a; S, D = rx(X(s)), rx(X(d)); xmm[X(d)] = [D[0], S[0], D[1], S[1]]
    elif op == 'unpckhps':
        s, d = a; S, D = rx(X(s)), rx(X(d)); xmm[X(d)] = [D[2], S[2], D[3], S[3]]
    elif op == 'movlhps':
        s, d = a; S, D = rx(X(s)), rx(X(d)); xmm[X(d)] = [D[0], D[1], S[0], S[1]]
    elif op == 'movhlps':
        s, d = a; S, D = rx(X(s)), rx(X(d)); xmm[X(d)] = [S[2], S[3], D[2], D[3]]
    elif op == 'movsldup':
        s, d = a; S = rx(X(s)); xmm[X(d)] = [S[0], S[0], S[2], S[2]]
    elif op == 'movshdup':
        s, d = a; S = rx(X(s)); xmm[X(d)] = [S[1], S[1], S[3], S[3]]
    elif op == 'shufps':
        imm, s, d = a; imm = int(imm[1:], 16); S, D = rx(X(s)), rx(X(d))
        xmm[X(d)] = [D[imm & 3], D[(imm >> 2) & 3], S[(imm >> 4) & 3], S[(imm >> 6) & 3]]
    elif op in ('movlps', 'movhps'):
        s, d = a; lo = 0 if op == 'movlps' else 2
        if X(s) is not None:
            k = parse_mem(d)
            for i in range(2): store((k[0], k[1] + 4 * i), rx(X(s), lo + i))
        else:
            k = parse_mem(s)
            for i in range(2): rx(X(d))[lo + i] = load((k[0], k[1] + 4 * i))
    else:
        raise SystemExit(f'unhandled {line}')

targets = {}
if wants:
    for w in wants:
        if w.startswith('xmm'): targets[w] = rx(int(w[3:]), 0)
        else:
            b, o = w.split(':'); o = int(o, 16)
            targets[w] = mem[(b, o)]
else:
    for k, v in stores.items():
        if k[0] != 'rsp': targets[f'{k[0]}:{k[1]:#x}'] = v
need = set()
def mark(n):
    stack = [n]
    while stack:
        n = stack.pop()
        if n in need or n is None: continue
        need.add(n); op, a, b = nodes[n]
        if op in ('mul', 'sub', 'add', 'div', 'xor', 'and'): stack += [a, b]
        elif op == 'sqrt': stack.append(a)
for v in targets.values(): mark(v)
sym = {'mul': '*', 'sub': '-', 'add': '+', 'div': '/'}
for n in sorted(need):
    op, a, b = nodes[n]
    if op == 'const': print(f't{n} = CONST({a:#010x})')
    elif op == 'in': print(f't{n} = IN("{a}")')
    elif op in ('xor', 'and'): print(f't{n} = {op.upper()}(t{a}, t{b})')
    elif op == 'sqrt': print(f't{n} = SQRT(t{a})')
    else: print(f't{n} = t{a} {sym[op]} t{b}')
for k, v in targets.items(): print(f'OUT["{k}"] = t{v}')
