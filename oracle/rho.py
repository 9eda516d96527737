"""Oracle: OpenCV's RHO estimator (``cv2.findHomography(..., cv2.RHO)``) and the LMEDS leg, restated (test infrastructure).

coordinate_model.py:354-357 falls through RANSAC -> RHO -> LMEDS; the arithmetic lives in opencv-python (calib3d: rho.cpp,
ptsetreg.cpp, fundam.cpp), which is not under /root/reference.  This file restates the published algorithm
(Bilaniuk's RHO: PROSAC sampling on an xorshift128+ stream, SPRT evaluation, non-randomness bound, float32 LM polish)
with the control flow and the float32 rounding points pinned against the cv2 binary of this image (4.13.0):
the 4-point solver ``hfunc`` lists its operations in the order the library executes them (recovered with
tools/cv2_probe/symexec.py), everything else follows the algorithm and is held to live cv2 by
tests/test_oracle_cascade.py (H bit for bit, masks, None decisions).
"""
from __future__ import annotations

import math

import numpy as np

from . import homography as hg

f32 = np.float32
FLT_EPSILON = f32(np.finfo(np.float32).eps)
MASK64 = (1 << 64) - 1


class XorShift128Plus:
    """RHO_HEST_REFC::fastSeed / fastRandom: xorshift128+ (23, 17, 26), 20 warm-up draws."""

    def __init__(self, seed: int = MASK64):
        self.s0 = seed & MASK64
        self.s1 = (~seed) & MASK64
        for _ in range(20):
            self.next()

    def next(self) -> int:
        x, y = self.s0, self.s1
        x ^= (x << 23) & MASK64
        x ^= x >> 17
        x ^= y ^ (y >> 26)
        self.s0, self.s1 = y, x
        return (x + y) & MASK64

    def uniform(self) -> float:
        return float(self.next()) * 5.421010862427522e-20   # 2**-64; uint64 -> double rounds to nearest


def rnd_smpl(rng: XorShift128Plus, sample_size: int, data_size: int):
    """rndSmpl: selection sampling when sample_size*2 > data_size, else draws until distinct."""
    out = []
    if sample_size * 2 > data_size:
        i = 0
        while len(out) < sample_size:
            u = rng.uniform()
            if float(data_size - i) * u < float(sample_size - len(out)):
                out.append(i)
            i += 1
    else:
        for _ in range(sample_size):
            while True:
                v = int(rng.uniform() * float(data_size)) & 0xFFFFFFFF
                if v not in out:
                    break
            out.append(v)
    return out


def _f2i(v) -> int:
    """(int) of a float as cvttss2si does it: truncation, INT_MIN when out of range / NaN."""
    v = float(v)
    if v != v or v >= 2147483648.0 or v <= -2147483649.0:
        return -(1 << 31)
    return int(v)


def sample_degenerate(P) -> bool:
    """isSampleDegenerate on the packed sample P = [x0,y0,..,x3,y3, X0,Y0,..,X3,Y3] (float32)."""
    x = [P[0], P[2], P[4], P[6]]; y = [P[1], P[3], P[5], P[7]]
    X = [P[8], P[10], P[12], P[14]]; Y = [P[9], P[11], P[13], P[15]]
    for a in range(4):
        for b in range(a + 1, 4):
            if x[a] == x[b] or y[a] == y[b]:
                return True

    def side(a, b, c, px, py):
        c0 = py[a] - py[b]
        c1 = px[b] - px[a]
        c2 = px[a] * py[b] - px[b] * py[a]
        return (px[c] * c0 + py[c] * c1) + c2

    for a, b, c in ((0, 1, 2), (0, 1, 3), (2, 3, 0), (2, 3, 1)):
        if (_f2i(side(a, b, c, x, y)) ^ _f2i(side(a, b, c, X, Y))) < 0:
            return True
    return False


def hfunc(P):
    """hFuncRefC: the float32 4-point homography, operations in the library's order (see module docstring)."""
    H = [None] * 9
    with np.errstate(all="ignore"):
        t1 = P[0]
        t2 = P[2]
        t3 = P[4]
        t4 = P[6]
        t5 = P[1]
        t6 = P[3]
        t7 = P[5]
        t8 = P[7]
        t9 = P[9]
        t11 = P[10]
        t16 = P[11]
        t18 = P[8]
        t20 = P[12]
        t23 = P[13]
        t34 = P[14]
        t35 = P[15]
        t45 = t4 - t3
        t48 = t34 - t20
        t69 = t3 * t20
        t70 = t18 * t1
        t71 = t3 * t23
        t72 = t7 * t20
        t73 = t69 - t70
        t74 = t7 * t23
        t75 = t5 - t7
        t76 = t1 * t9
        t77 = t1 - t3
        t78 = t2 - t3
        t79 = t71 - t76
        t80 = t18 * t5
        t81 = t5 * t9
        t82 = t9 - t23
        t83 = t73 * t78
        t84 = t72 - t80
        t85 = t74 - t81
        t86 = t78 * t75
        t87 = t18 - t20
        t88 = t6 - t7
        t89 = t88 * t77
        t90 = t89 - t86
        t91 = t11 * t2
        t92 = t69 - t91
        t93 = t11 * t6
        t94 = t92 * t77
        t95 = t72 - t93
        t96 = t94 - t83
        t97 = t95 * t77
        t98 = t84 * t78
        t99 = t11 - t20
        t100 = t97 - t98
        t101 = t99 * t77
        t102 = t87 * t78
        t103 = t101 - t102
        t104 = t2 * t16
        t105 = t6 * t16
        t106 = t71 - t104
        t107 = t78 * t79
        t108 = t106 * t77
        t109 = t108 - t107
        t110 = t74 - t105
        t111 = t110 * t77
        t112 = t85 * t78
        t113 = t78 * t82
        t114 = t111 - t112
        t115 = t16 - t23
        t116 = t115 * t77
        t117 = t116 - t113
        t118 = t8 - t7
        t119 = t45 * t75
        t120 = t118 * t77
        t121 = t120 - t119
        t122 = t34 * t4
        t123 = t4 * t35
        t124 = t69 - t122
        t125 = t73 * t45
        t126 = t124 * t77
        t127 = t126 - t125
        t128 = t96 * t121
        t129 = t127 * t90
        t130 = t129 - t128
        t131 = t71 - t123
        t132 = t45 * t79
        t133 = t131 * t77
        t134 = t79 * t90
        t135 = t133 - t132
        t136 = t109 * t121
        t137 = t135 * t90
        t138 = t137 - t136
        t139 = t77 * t90
        t140 = t139 - t86
        t141 = f32(1.0)
        t142 = t141 / t140
        t143 = t73 * t90
        t144 = t75 * t96
        t145 = t143 - t144
        t146 = t84 * t90
        t147 = t145 * t142
        t148 = t100 * t75
        t149 = t146 - t148
        t150 = t87 * t90
        t151 = t149 * t142
        t152 = t114 * t75
        t153 = t103 * t75
        t154 = t150 - t153
        t155 = t85 * t90
        t156 = t154 * t142
        t157 = t75 * t109
        t158 = t75 * t117
        t159 = t134 - t157
        t160 = t159 * t142
        t161 = t155 - t152
        t162 = t161 * t142
        t163 = t82 * t90
        t164 = t82 * t45
        t165 = t163 - t158
        t166 = t165 * t142
        t167 = t141 / t90
        t168 = t100 * t167
        t169 = t96 * t167
        t170 = t109 * t167
        t171 = t103 * t167
        t172 = t114 * t167
        t173 = t167 * t117
        t174 = None
        t175 = t147 * t3
        t176 = -t69
        t180 = t169 * t7
        t181 = t175 + t180
        t182 = t160 * t3
        t183 = t176 - t181
        t184 = -t71
        t188 = t7 * t170
        t189 = t182 + t188
        t190 = t184 - t189
        t191 = t35 * t8
        t192 = t74 - t191
        t193 = t85 * t45
        t194 = t192 * t77
        t195 = t194 - t193
        t196 = t35 - t23
        t197 = t114 * t121
        t198 = t195 * t90
        t199 = t196 * t77
        t200 = t198 - t197
        t201 = t200 / t138
        t202 = t199 - t164
        t203 = t202 * t90
        t204 = t117 * t121
        t205 = t203 - t204
        t206 = t205 / t138
        t207 = t8 * t34
        t208 = t48 * t77
        t209 = t87 * t45
        t210 = t45 * t84
        t211 = t208 - t209
        t212 = t103 * t121
        t213 = t121 * t100
        t214 = t211 * t90
        t215 = t214 - t212
        t216 = t130 * t206
        t217 = t215 - t216
        t218 = t72 - t207
        t219 = t218 * t77
        t220 = t219 - t210
        t221 = t220 * t90
        t222 = t173 * t7
        t223 = t221 - t213
        t224 = t170 * t206
        t225 = t130 * t201
        t226 = t170 * t201
        t227 = t223 - t225
        t228 = t217 / t227
        t229 = t173 - t224
        t230 = t166 * t3
        t231 = t230 + t222
        t232 = t190 * t206
        t233 = t23 - t231
        t234 = t162 * t3
        t235 = t233 - t232
        t236 = -t74
        t240 = t172 - t226
        t241 = t172 * t7
        t242 = t241 + t234
        t243 = t236 - t242
        t244 = t190 * t201
        t245 = t171 * t7
        t246 = t243 - t244
        t247 = t156 * t3
        t248 = t247 + t245
        t249 = t240 * t228
        t250 = t246 * t228
        t253 = t20 - t248
        t254 = t183 * t206
        t255 = t229 - t249
        t256 = t235 - t250
        t259 = t147 * t206
        t260 = t147 * t201
        t261 = t253 - t254
        t262 = t160 * t206
        t263 = t156 - t259
        t264 = t169 * t206
        t265 = t166 - t262
        t266 = t3 * t151
        t267 = t171 - t264
        t268 = t151 - t260
        t269 = t169 * t201
        t270 = t7 * t168
        t271 = t168 - t269
        t272 = -t72
        t276 = t270 + t266
        t277 = t272 - t276
        t278 = t183 * t201
        t279 = t277 - t278
        t280 = t160 * t201
        t281 = t201 * t228
        t282 = t162 - t280
        t283 = t206 - t281
        t284 = t228 * t268
        t285 = t228 * t271
        t286 = t228 * t279
        t287 = t228 * t282
        t288 = t263 - t284
        t289 = t267 - t285
        t290 = t261 - t286
        t291 = t265 - t287
        H[0] = t288
        H[1] = t289
        H[2] = t290
        H[3] = t291
        H[4] = t255
        H[5] = t256
        H[6] = t283
        H[7] = t228
        H[8] = t141
    return np.array(H, dtype=np.float32)


def _design_sprt(delta, eps, t_M=25.0, m_S=1.0):
    """designSPRTTest / sacDesignSPRTTest in IEEE double semantics (eps == 1 gives infinities, as in C)."""
    d = np.float64
    with np.errstate(all="ignore"):
        delta, eps = d(delta), d(eps)
        acc = delta / eps
        rej = (d(1) - delta) / (d(1) - eps)
        C = (d(1) - delta) * np.log(rej) + delta * np.log(acc)
        K = C * d(t_M) / d(m_S) + d(1)
        An = K
        for _ in range(10):
            prev = An
            An = K + np.log(An)
            if not (An - prev > 1.5e-8):
                break
    return float(An), float(acc), float(rej)


def _iter_bound(cfd, inlier_rate, max_bound):
    p = 1.0 - math.pow(inlier_rate, 4.0)
    if p >= 1.0:
        ret = max_bound
    elif p <= 0.0:
        ret = 1
    else:
        ret = int(math.ceil(math.log(1.0 - cfd) / math.log(p))) & 0xFFFFFFFF
    return min(ret, max_bound)


def nonrand_table(N, beta=0.35):
    tbl = [0] * (N + 1)
    k = math.sqrt(beta * (1.0 - beta)) * 1.645
    for n in range(5, N):
        tbl[n] = int(math.ceil(4 + n * beta + math.sqrt(float(n)) * k))
    return tbl


def _jacobian_errors(H, src, dst, inl, want):
    """sacCalcJacobianErrors: float32 sums over the inliers, lower triangle of JtJ."""
    JtJ = np.zeros((8, 8), np.float32); Jte = np.zeros(8, np.float32); S = f32(0)
    one = f32(1)
    for i in range(len(src)):
        if not inl[i]:
            continue
        x, y = src[i]; X, Y = dst[i]
        W = (H[6] * x + H[7] * y) + one
        iW = one / W if abs(W) > FLT_EPSILON else f32(0)
        rx = ((H[0] * x + H[1] * y) + H[2]) * iW
        ry = ((H[3] * x + H[4] * y) + H[5]) * iW
        eX = rx - X; eY = ry - Y
        S = S + (eX * eX + eY * eY)
        if want:
            d11 = x * iW; d12 = y * iW; d13 = iW
            d31x = -rx * x * iW; d32x = -rx * y * iW
            d31y = -ry * x * iW; d32y = -ry * y * iW
            Jte[0] += eX * d11; Jte[1] += eX * d12; Jte[2] += eX * d13
            Jte[3] += eY * d11; Jte[4] += eY * d12; Jte[5] += eY * d13
            Jte[6] += eX * d31x + eY * d31y
            Jte[7] += eX * d32x + eY * d32y
            JtJ[0, 0] += d11 * d11
            JtJ[1, 0] += d11 * d12; JtJ[1, 1] += d12 * d12
            JtJ[2, 0] += d11 * d13; JtJ[2, 1] += d12 * d13; JtJ[2, 2] += d13 * d13
            JtJ[3, 3] += d11 * d11
            JtJ[4, 3] += d11 * d12; JtJ[4, 4] += d12 * d12
            JtJ[5, 3] += d11 * d13; JtJ[5, 4] += d12 * d13; JtJ[5, 5] += d13 * d13
            JtJ[6, 0] += d11 * d31x; JtJ[6, 1] += d12 * d31x; JtJ[6, 2] += d13 * d31x
            JtJ[6, 3] += d11 * d31y; JtJ[6, 4] += d12 * d31y; JtJ[6, 5] += d13 * d31y
            JtJ[6, 6] += d31x * d31x + d31y * d31y
            JtJ[7, 0] += d11 * d32x; JtJ[7, 1] += d12 * d32x; JtJ[7, 2] += d13 * d32x
            JtJ[7, 3] += d11 * d32y; JtJ[7, 4] += d12 * d32y; JtJ[7, 5] += d13 * d32y
            JtJ[7, 6] += d31x * d32x + d31y * d32y
            JtJ[7, 7] += d32x * d32x + d32y * d32y
    return JtJ, Jte, S


def _chol_damped(A, lam):
    L = np.zeros((8, 8), np.float32)
    lp1 = f32(lam) + f32(1)
    for i in range(8):
        for j in range(i):
            x = A[i, j]
            for k in range(j):
                x = x - L[i, k] * L[j, k]
            L[i, j] = x / L[j, j]
        x = A[i, i] * lp1
        for k in range(i):
            x = x - L[i, k] * L[i, k]
        if x < 0:
            return None
        L[i, i] = np.sqrt(x)
    return L


def _lm_step(L, Jte, H):
    """sacTRInv8x8 + sacTRISolve8x8 + sacSub8x1 on the Cholesky factor L: dH = L^-T (L^-1 Jte), newH = H - dH, with the
    operations in the library's order (tools/cv2_probe/symexec2.py)."""
    newH = [None] * 8
    dH = [None] * 8
    t1 = f32(1.0)
    t2 = L[1, 1]
    t3 = t1 / t2
    t8 = -t3
    t12 = L[2, 2]
    t13 = t1 / t12
    t14 = L[1, 0]
    t15 = t8 * t14
    t16 = L[0, 0]
    t17 = t1 / t16
    t18 = L[3, 3]
    t19 = t1 / t18
    t20 = L[5, 5]
    t21 = t1 / t20
    t22 = L[4, 4]
    t23 = t1 / t22
    t24 = L[7, 7]
    t25 = t1 / t24
    t26 = L[6, 6]
    t27 = t1 / t26
    t28 = t15 * t17
    t29 = -t19
    t33 = L[3, 2]
    t34 = t29 * t33
    t35 = t34 * t13
    t36 = -t21
    t40 = L[5, 4]
    t41 = t36 * t40
    t42 = t41 * t23
    t43 = -t25
    t47 = L[7, 6]
    t48 = t43 * t47
    t49 = t48 * t27
    t50 = L[2, 1]
    t51 = t13 * t50
    t52 = t50 * t35
    t53 = L[3, 1]
    t54 = t28 * t51
    t55 = t53 * t19
    t56 = t51 * t3
    t57 = t55 + t52
    t58 = L[2, 0]
    t59 = t13 * t58
    t60 = t59 * t17
    t61 = t60 + t54
    t64 = t35 * t58
    t65 = L[3, 0]
    t66 = -t61
    t67 = -t56
    t70 = t65 * t19
    t71 = t70 + t64
    t72 = t28 * t57
    t73 = t57 * t3
    t74 = t71 * t17
    t75 = t74 + t72
    t76 = -t75
    t80 = -t73
    t84 = L[6, 4]
    t85 = L[6, 5]
    t86 = t84 * t49
    t87 = t27 * t85
    t88 = t85 * t49
    t89 = L[7, 5]
    t90 = t89 * t25
    t91 = t90 + t88
    t92 = L[7, 4]
    t93 = t92 * t25
    t94 = t91 * t42
    t95 = t91 * t21
    t96 = t93 + t86
    t97 = t96 * t23
    t98 = t42 * t87
    t99 = t87 * t21
    t100 = t94 + t97
    t101 = t84 * t27
    t102 = -t100
    t106 = t101 * t23
    t107 = t106 + t98
    t108 = L[4, 0]
    t109 = -t99
    t113 = -t107
    t117 = L[4, 1]
    t118 = t23 * t117
    t119 = -t95
    t123 = L[4, 2]
    t124 = t123 * t23
    t125 = L[4, 3]
    t126 = t125 * t23
    t127 = L[5, 0]
    t128 = t42 * t117
    t129 = L[5, 1]
    t130 = t129 * t21
    t131 = t128 + t130
    t132 = t123 * t42
    t133 = L[5, 2]
    t134 = t133 * t21
    t135 = t132 + t134
    t136 = t125 * t42
    t137 = L[5, 3]
    t138 = t137 * t21
    t139 = t136 + t138
    t140 = t113 * t117
    t141 = t109 * t129
    t142 = t117 * t102
    t143 = L[6, 1]
    t144 = t143 * t27
    t145 = t141 + t140
    t146 = t145 + t144
    t147 = t109 * t133
    t148 = t123 * t113
    t149 = L[6, 2]
    t150 = t149 * t27
    t151 = t147 + t148
    t152 = t151 + t150
    t153 = t109 * t137
    t154 = t125 * t113
    t155 = L[6, 3]
    t156 = t155 * t27
    t157 = t153 + t154
    t158 = t157 + t156
    t159 = t129 * t119
    t160 = t159 + t142
    t161 = t143 * t49
    t162 = t160 + t161
    t163 = L[7, 1]
    t164 = t163 * t25
    t165 = t162 + t164
    t166 = t102 * t123
    t167 = t133 * t119
    t168 = t167 + t166
    t169 = t149 * t49
    t170 = t168 + t169
    t171 = L[7, 2]
    t172 = t171 * t25
    t173 = t119 * t137
    t174 = t170 + t172
    t175 = t102 * t125
    t176 = t174 * t66
    t177 = t173 + t175
    t178 = t155 * t49
    t179 = t177 + t178
    t180 = L[7, 3]
    t181 = t180 * t25
    t182 = t179 + t181
    t183 = t102 * t108
    t184 = t119 * t127
    t185 = t184 + t183
    t186 = L[6, 0]
    t187 = t186 * t49
    t188 = L[7, 0]
    t189 = t188 * t25
    t190 = t187 + t185
    t191 = t190 + t189
    t192 = t165 * t28
    t193 = t191 * t17
    t194 = t193 + t192
    t195 = t176 + t194
    t196 = t182 * t76
    t197 = t118 * t28
    t198 = t195 + t196
    t199 = t124 * t66
    t200 = t108 * t23
    t201 = t200 * t17
    t202 = t197 + t201
    t203 = t199 + t202
    t204 = t76 * t126
    t205 = t203 + t204
    t206 = t118 * t3
    t207 = -t205
    t211 = t124 * t67
    t212 = t211 + t206
    t213 = t80 * t126
    t214 = t212 + t213
    t215 = -t214
    t219 = t124 * t13
    t220 = t35 * t126
    t221 = t126 * t19
    t222 = t219 + t220
    t223 = -t222
    t227 = t135 * t66
    t228 = -t221
    t232 = t108 * t42
    t233 = t127 * t21
    t234 = t232 + t233
    t235 = t131 * t28
    t236 = t234 * t17
    t237 = t235 + t236
    t238 = t131 * t3
    t239 = t227 + t237
    t240 = t76 * t139
    t241 = t239 + t240
    t242 = -t241
    t246 = t135 * t67
    t247 = t246 + t238
    t248 = t80 * t139
    t249 = t135 * t13
    t250 = t247 + t248
    t251 = -t250
    t255 = t35 * t139
    t256 = t139 * t19
    t257 = t249 + t255
    t258 = -t256
    t262 = -t257
    t266 = t127 * t109
    t267 = t108 * t113
    t268 = t186 * t27
    t269 = t266 + t267
    t270 = t146 * t28
    t271 = t269 + t268
    t272 = t271 * t17
    t273 = t270 + t272
    t274 = t66 * t152
    t275 = t274 + t273
    t276 = t76 * t158
    t277 = t275 + t276
    t278 = t152 * t13
    t279 = -t277
    t283 = t146 * t3
    t284 = t152 * t67
    t285 = t284 + t283
    t286 = t80 * t158
    t287 = t285 + t286
    t288 = -t287
    t292 = t35 * t158
    t293 = t158 * t19
    t294 = t278 + t292
    t295 = -t293
    t299 = -t294
    t303 = -t198
    t307 = t165 * t3
    t308 = t174 * t67
    t309 = t308 + t307
    t310 = t174 * t13
    t311 = t80 * t182
    t312 = t309 + t311
    t313 = t182 * t19
    t314 = -t312
    t318 = t35 * t182
    t319 = -t313
    t323 = t310 + t318
    t324 = -t323
    t328 = Jte[1]
    t329 = Jte[2]
    t330 = Jte[0]
    t331 = t328 * t3
    t332 = t28 * t330
    t333 = t76 * t330
    t334 = t332 + t331
    t335 = t328 * t67
    t336 = t66 * t330
    t337 = t335 + t336
    t338 = t13 * t329
    t339 = t328 * t80
    t340 = t337 + t338
    t341 = Jte[3]
    t342 = t339 + t333
    t343 = t329 * t35
    t344 = t341 * t19
    t345 = t343 + t342
    t346 = Jte[4]
    t347 = t345 + t344
    t348 = t328 * t215
    t349 = t329 * t223
    t350 = t207 * t330
    t351 = t348 + t350
    t352 = t341 * t228
    t353 = t349 + t351
    t354 = t352 + t353
    t355 = t329 * t262
    t356 = t346 * t23
    t357 = t354 + t356
    t358 = t242 * t330
    t359 = Jte[5]
    t360 = t328 * t251
    t361 = t360 + t358
    t362 = t341 * t258
    t363 = t355 + t361
    t364 = Jte[6]
    t365 = t362 + t363
    t366 = t329 * t299
    t367 = t346 * t42
    t368 = t367 + t365
    t369 = t359 * t21
    t370 = t368 + t369
    t371 = t279 * t330
    t372 = t328 * t288
    t373 = t372 + t371
    t374 = t341 * t295
    t375 = t366 + t373
    t376 = t374 + t375
    t377 = t346 * t113
    t378 = t364 * t27
    t379 = t377 + t376
    t380 = t109 * t359
    t381 = t380 + t379
    t382 = t381 + t378
    t383 = Jte[7]
    t384 = t328 * t314
    t385 = t303 * t330
    t386 = t329 * t324
    t387 = t359 * t119
    t388 = t384 + t385
    t389 = t341 * t319
    t390 = t386 + t388
    t391 = t389 + t390
    t392 = t346 * t102
    t393 = t364 * t49
    t394 = t392 + t391
    t395 = t207 * t357
    t396 = t387 + t394
    t397 = t383 * t25
    t398 = t393 + t396
    t399 = t330 * t17
    t400 = t397 + t398
    t401 = t399 * t17
    t402 = t28 * t334
    t403 = t66 * t340
    t404 = t402 + t401
    t405 = t403 + t404
    t406 = t76 * t347
    t407 = t406 + t405
    t408 = t407 + t395
    t409 = t242 * t370
    t410 = t409 + t408
    t411 = t279 * t382
    t412 = t411 + t410
    t413 = t303 * t400
    t414 = t412 + t413
    t415 = t334 * t3
    t416 = t67 * t340
    t417 = t416 + t415
    t418 = t80 * t347
    t419 = t23 * t357
    t420 = t42 * t370
    t421 = t251 * t370
    t422 = t13 * t340
    t423 = t418 + t417
    t424 = t215 * t357
    t425 = t21 * t370
    t426 = t420 + t419
    t427 = t102 * t400
    t428 = t423 + t424
    t429 = t288 * t382
    t430 = t421 + t428
    t431 = t429 + t430
    t432 = t314 * t400
    t433 = t432 + t431
    t434 = t35 * t347
    t435 = t223 * t357
    t436 = t434 + t422
    t437 = t262 * t370
    t438 = t436 + t435
    t439 = t437 + t438
    t440 = t299 * t382
    t441 = t324 * t400
    t442 = t440 + t439
    t443 = t442 + t441
    t444 = t228 * t357
    t445 = t347 * t19
    t446 = t258 * t370
    t447 = t444 + t445
    t448 = t446 + t447
    t449 = t295 * t382
    t450 = t319 * t400
    t451 = t449 + t448
    t452 = t451 + t450
    t453 = t113 * t382
    t454 = t27 * t382
    t455 = t453 + t426
    t456 = t455 + t427
    t457 = t109 * t382
    t458 = t457 + t425
    t459 = t119 * t400
    t460 = t49 * t400
    t461 = t458 + t459
    t462 = t400 * t25
    t463 = t454 + t460
    t464 = H[0]
    t466 = H[1]
    t467 = t464 - t414
    t468 = t466 - t433
    t469 = H[2]
    t470 = t469 - t443
    t471 = H[3]
    t472 = t471 - t452
    t473 = H[4]
    t474 = t473 - t456
    t475 = H[5]
    t476 = t475 - t461
    t477 = H[6]
    t478 = t477 - t463
    t479 = H[7]
    t480 = t479 - t462
    newH[0] = t467
    newH[1] = t468
    newH[2] = t470
    newH[3] = t472
    newH[4] = t474
    newH[5] = t476
    newH[6] = t478
    newH[7] = t480
    dH[0] = t414
    dH[1] = t433
    dH[2] = t443
    dH[3] = t452
    dH[4] = t456
    dH[5] = t461
    dH[6] = t463
    dH[7] = t462
    return np.array(newH, np.float32), np.array(dH, np.float32)


def refine(H, src, dst, inl):
    """RHO_HEST_REFC::refine: <= 100 float32 LM iterations over the 8 free entries, damped Cholesky steps."""
    H = H.copy()
    L = f32(100.0)
    with np.errstate(all="ignore"):
        JtJ, Jte, S = _jacobian_errors(H, src, dst, inl, True)
        for _ in range(100):
            while True:
                C = _chol_damped(JtJ, L)
                if C is not None:
                    break
                L = L * f32(2.0)
            n8, dH = _lm_step(C, Jte, H)
            newH = H.copy()
            newH[:8] = n8
            _, _, newS = _jacobian_errors(newH, src, dst, inl, False)
            # sacLMGain: dS / (0.5 * (lambda * |dH|^2 + dH . Jte)), sums in the library's order
            dS = S - newS
            sq = dH[0] * dH[0] + f32(0)
            for i in range(1, 8):
                sq = sq + dH[i] * dH[i]
            dL = dH[0] * Jte[0] + sq * L
            for i in range(1, 8):
                dL = dL + Jte[i] * dH[i]
            dL = dL * f32(0.5)
            gain = dS if FLT_EPSILON > abs(dL) else dS / dL
            if gain < f32(0.25):
                L = L * f32(8)
                if L > f32(1000.0) / FLT_EPSILON:
                    break
            elif gain > f32(0.75):
                L = L * f32(0.5)
            if gain > 0:
                H = newH
                JtJ, Jte, S = _jacobian_errors(H, src, dst, inl, True)
    return H


def rho_hest(img_pts, world_pts, maxD=3.0, maxI=2000, rConvg=2000, cfd=0.995, minInl=4, beta=0.35, trace=None):
    """rhoHest with flags NR | FINAL_REFINEMENT, as createAndRunRHORegistrator calls it.  -> (H float32 3x3 | None, mask)."""
    src = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    N = len(src)
    rng = XorShift128Plus()
    tbl = nonrand_table(N, beta)
    phNum, phEndI, phMax, phNumInl = 4, 1, N, 0
    phEndFpI = rConvg * 24.0 / (float(N) * float(N - 1) * float(N - 2) * float(N - 3))
    bestH = np.zeros(9, np.float32); best_inl = np.zeros(N, np.uint8); best_num = 0
    eps_, delta = 0.1, 0.01
    A, lam_acc, lam_rej = _design_sprt(delta, eps_)
    maxD2 = f32(maxD) * f32(maxD)
    one = f32(1)
    i = 0
    with np.errstate(all="ignore"):
        while i < maxI or i < 100:
            # --- hypothesize
            if i >= phEndI and phNum < phMax:
                phNum += 1
                nxt = (phEndFpI * float(phNum)) / float(phNum - 4)
                phEndI = (phEndI + int(math.ceil(nxt - phEndFpI))) & 0xFFFFFFFF
                phEndFpI = nxt
            if i > phEndI:
                smpl = rnd_smpl(rng, 4, phNum)
            else:
                smpl = rnd_smpl(rng, 3, phNum - 1) + [phNum - 1]
            P = np.concatenate([src[smpl].reshape(-1), dst[smpl].reshape(-1)])
            if sample_degenerate(P):
                i += 1; continue
            H = hfunc(P)
            if np.isnan(((((((H[0] + H[1]) + H[2]) + H[3]) + H[4]) + H[5]) + H[6]) + H[7]):
                i += 1; continue
            # --- verify: SPRT evaluation
            inl = np.zeros(N, np.uint8); num = 0; lam = 1.0; good = True; tested = 0
            for k in range(N):
                x, y = src[k]; X, Y = dst[k]
                rx = (H[0] * x + H[1] * y) + H[2]
                ry = (H[3] * x + H[4] * y) + H[5]
                rz = (H[6] * x + H[7] * y) + one
                rx = rx / rz - X; ry = ry / rz - Y
                d = rx * rx + ry * ry
                isin = bool(maxD2 >= d)
                num += isin; inl[k] = isin
                lam *= lam_acc if isin else lam_rej
                good = A >= lam
                tested = k + 1
                if not good:
                    break
            if good:
                if num > best_num:
                    eps_ = float(num) / float(N)
                    A, lam_acc, lam_rej = _design_sprt(delta, eps_)
            else:
                nd = float(num) / float(tested)
                if nd > 0 and abs(delta - nd) / delta > 0.1:
                    delta = nd
                    A, lam_acc, lam_rej = _design_sprt(delta, eps_)
            if num > best_num:
                bestH, best_inl, best_num = H, inl, num
                if trace is not None:
                    trace.append((i, list(smpl), num))
                maxI = _iter_bound(cfd, float(best_num) / float(N), maxI)
                # non-randomness (nStarOptimize)
                best_n, bn = N, best_num
                test_n, tn = N, best_num
                while test_n > 20 and tn:
                    if tn * best_n > bn * test_n:
                        if tn < tbl[test_n]:
                            break
                        best_n, bn = test_n, tn
                    tn -= 1 if best_inl[test_n - 1] else 0
                    test_n -= 1
                if bn * phMax > phNumInl * best_n:
                    phMax, phNumInl = best_n, bn
                    maxI = _iter_bound(cfd, float(phNumInl) / float(phMax), maxI)
            i += 1
        if best_num > 4:
            bestH = refine(bestH, src, dst, best_inl)
    if best_num < minInl:
        return None, np.zeros((N, 1), np.uint8)
    return bestH.reshape(3, 3), best_inl.reshape(-1, 1).copy()


def find_homography_rho(img_pts, world_pts, thresh=None):
    """cv2.findHomography(img, wor, cv2.RHO, thresh): npoints == 4 takes the plain runKernel path (fundam.cpp)."""
    src = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    if len(src) < 4:
        return None, None
    if len(src) == 4:
        H = hg.run_kernel(src, dst)
        return (None, None) if H is None else (H, np.ones((4, 1), np.uint8))
    H, mask = rho_hest(src, dst, maxD=3.0 if not thresh or thresh <= 0 else thresh)
    if H is None:
        return None, None
    return H.astype(np.float64), mask


def find_homography_lmeds(img_pts, world_pts, thresh=None, max_iters=2000, confidence=0.995):
    """cv2.findHomography(img, wor, cv2.LMEDS, thresh): LMeDSPointSetRegistrator::run + the common refit; the returned
    mask is the inlier set of the refined H at the reprojection threshold (default 3), like the RANSAC leg."""
    src = np.ascontiguousarray(img_pts, np.float32).reshape(-1, 2)
    dst = np.ascontiguousarray(world_pts, np.float32).reshape(-1, 2)
    count = len(src)
    if count < 4:
        return None, None
    if count == 4:
        H = hg.run_kernel(src, dst)
        return (None, None) if H is None else (H, np.ones((4, 1), np.uint8))
    thr = 3.0 if not thresh or thresh <= 0 else thresh
    rng = hg.CvRNG()
    niters = max(hg.ransac_update_num_iters(confidence, 0.45, 4, max_iters), 3)
    min_median, best = float("inf"), None
    for it in range(niters):
        found = False
        for _ in range(1000):   # getSubset default maxAttempts (the RANSAC leg passes 10000)
            idx = hg.draw_subset(rng, count)
            if hg.check_subset(src[idx], dst[idx]):
                found = True
                break
        if not found:
            if it == 0:
                return None, None
            break
        H = hg.run_kernel(src[idx], dst[idx])
        if H is None:
            continue
        err = hg.compute_error(src, dst, H)
        med = float(np.sort(err.view(np.int32))[count // 2: count // 2 + 1].view(np.float32)[0])  # nth_element on the int view
        if med < min_median:
            min_median, best = med, H
    if best is None:
        return None, None
    sigma = max(2.5 * 1.4826 * (1 + 5.0 / (count - 4)) * math.sqrt(min_median), 0.001)
    good, mask = hg.find_inliers(src, dst, best, sigma)
    if good < 4:
        return None, None
    sel = mask.astype(bool)
    H = best
    Hk = hg.run_kernel(src[sel], dst[sel])
    if Hk is not None:
        H = Hk
    H, _ = hg.lm_refine(H, src[sel], dst[sel], 10)
    _, mask = hg.find_inliers(src, dst, H, thr)
    return H, mask.reshape(-1, 1)


def find_homography_cascade_restated(img_pts, world_pts):
    """coordinate_model.py:354-357 with all three legs restated.  -> (H | None, mask | None, leg 0/1/2 | None)."""
    H, mask = hg.find_homography_restated(img_pts, world_pts, 5.0)
    if H is not None:
        return H, mask, 0
    H, mask = find_homography_rho(img_pts, world_pts)
    if H is not None:
        return H, mask, 1
    H, mask = find_homography_lmeds(img_pts, world_pts)
    if H is not None:
        return H, mask, 2
    return None, None, None
