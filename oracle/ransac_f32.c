/*
 * Oracle (TEST INFRASTRUCTURE) -- C restatement of the "fixed-K arithmetic" of the CUDA hypothesis
 * kernel (DESIGN.md, section "fixed-K arithmetic"), written independently from the specification so
 * that the kernel's winning hypothesis index and inlier mask can be checked BIT FOR BIT when both
 * sides are fed the same hypothesis table.  Build: gcc -O2 -ffp-contract=off (oracle/build_c.py);
 * every fused multiply-add of the specification is an explicit fmaf(), everything else is a single
 * rounded float operation.
 *
 * The algorithm follows what the reference's cv2.findHomography(RANSAC) does per iteration
 * (eagle/models/coordinate_model.py:355 -> OpenCV RANSACPointSetRegistrator::run): take a 4-point
 * sample, reject it by the checkSubset rules, solve the 4-point DLT, count the points whose
 * reprojection error is within thr, keep the first hypothesis with the most inliers.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

typedef struct { float cX, cY, sX, sY, cx, cy, rt; } norm_t;

static float det_rows(float Xa, float Ya, float Xb, float Yb, float Xc, float Yc)
{
    float t = fmaf(Xb, Yc, -(Xc * Yb));
    return fmaf(Xa, Yb - Yc, fmaf(-Ya, Xb - Xc, t));
}

/* sum of v[0..n), n <= 64, in the specification's tree order: 32 partial sums t[i] = v[i] + v[i+32] (absent terms are
 * +0), then five pairwise rounds combining t[i] with t[i ^ 16], t[i ^ 8], ... t[i ^ 1]; the total is t[0] */
static float tree_total(const float *v, int n)
{
    float t[32], u[32];
    int i, m;
    for (i = 0; i < 32; i++) {
        float lo = (i < n) ? v[i] : 0.0f, hi = (i + 32 < n) ? v[i + 32] : 0.0f;
        t[i] = lo + hi;
    }
    for (m = 16; m >= 1; m /= 2) {
        for (i = 0; i < 32; i++) u[i] = t[i] + t[i ^ m];
        memcpy(t, u, sizeof(t));
    }
    return t[0];
}

/* frame normalisation */
static int frame_norm(const float *X, const float *Y, const float *x, const float *y, int n, float inv_thr, norm_t *m)
{
    float dX[64], dY[64], a, b;
    int i;
    m->cX = tree_total(X, n) / (float)n; m->cY = tree_total(Y, n) / (float)n;
    m->cx = tree_total(x, n) / (float)n; m->cy = tree_total(y, n) / (float)n;
    for (i = 0; i < n; i++) { dX[i] = fabsf(X[i] - m->cX); dY[i] = fabsf(Y[i] - m->cY); }
    a = tree_total(dX, n); b = tree_total(dY, n);
    if (!(a > 0) || !(b > 0)) return 0;
    m->sX = (float)n / a; m->sY = (float)n / b; m->rt = inv_thr;
    return 1;
}

/* 4 normalised correspondences q[k] = {X,Y,x,y} -> h[8]; returns 0 if the sample is rejected */
static int hypothesis(float q[4][4], float *h)
{
    float X[4], Y[4], x[4], y[4], S[4], D[4], n[4];
    int k, neg, ok;
    for (k = 0; k < 4; k++) { X[k] = q[k][0]; Y[k] = q[k][1]; x[k] = q[k][2]; y[k] = q[k][3]; }
    /* triples (0,1,2) (1,2,3) (0,2,3) (0,1,3) */
    S[0] = det_rows(X[0], Y[0], X[1], Y[1], X[2], Y[2]); D[0] = det_rows(x[0], y[0], x[1], y[1], x[2], y[2]);
    S[1] = det_rows(X[1], Y[1], X[2], Y[2], X[3], Y[3]); D[1] = det_rows(x[1], y[1], x[2], y[2], x[3], y[3]);
    S[2] = det_rows(X[0], Y[0], X[2], Y[2], X[3], Y[3]); D[2] = det_rows(x[0], y[0], x[2], y[2], x[3], y[3]);
    S[3] = det_rows(X[0], Y[0], X[1], Y[1], X[3], Y[3]); D[3] = det_rows(x[0], y[0], x[1], y[1], x[3], y[3]);
    ok = 1;
    for (k = 1; k < 4; k++) ok = ok && fabsf(S[k]) > 1e-6f && fabsf(D[k]) > 1e-6f; /* last point vs earlier pairs */
    neg = 0;
    for (k = 0; k < 4; k++) neg += (S[k] * D[k]) < 0.0f;
    ok = ok && (neg == 0 || neg == 4);
    n[0] = S[1]; n[1] = -S[2]; n[2] = S[3]; n[3] = -S[0];
    for (k = 0; k < 3; k++) {           /* largest |n| to slot 3, compare-exchange in order k = 0,1,2 */
        if (fabsf(n[k]) > fabsf(n[3])) {
            float t;
            t = n[3]; n[3] = n[k]; n[k] = t;
            t = X[3]; X[3] = X[k]; X[k] = t;
            t = Y[3]; Y[3] = Y[k]; Y[k] = t;
            t = x[3]; x[3] = x[k]; x[k] = t;
            t = y[3]; y[3] = y[k]; y[k] = t;
        }
    }
    {
        float a11 = 0, a12 = 0, a21 = 0, a22 = 0, b1 = 0, b2 = 0, Dt, rD, h6, h7, u[3], v[3], dP, rP;
        float c0, c1, c2, d0, d1, d2, e0, e1, e2, acc;
        for (k = 0; k < 4; k++) {
            float nx = n[k] * x[k], ny = n[k] * y[k];
            a11 = fmaf(-nx, X[k], a11); a12 = fmaf(-nx, Y[k], a12); b1 = b1 + nx;
            a21 = fmaf(-ny, X[k], a21); a22 = fmaf(-ny, Y[k], a22); b2 = b2 + ny;
        }
        Dt = fmaf(a11, a22, -(a12 * a21));
        rD = 1.0f / Dt;
        h6 = fmaf(b1, a22, -(a12 * b2)) * rD;
        h7 = fmaf(a11, b2, -(b1 * a21)) * rD;
        for (k = 0; k < 3; k++) {
            float w = fmaf(h6, X[k], fmaf(h7, Y[k], 1.0f));
            u[k] = x[k] * w; v[k] = y[k] * w;
        }
        dP = det_rows(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
        rP = 1.0f / dP;
        c0 = Y[1] - Y[2]; c1 = Y[2] - Y[0]; c2 = Y[0] - Y[1];
        d0 = X[2] - X[1]; d1 = X[0] - X[2]; d2 = X[1] - X[0];
        e0 = fmaf(X[1], Y[2], -(X[2] * Y[1])); e1 = fmaf(X[2], Y[0], -(X[0] * Y[2])); e2 = fmaf(X[0], Y[1], -(X[1] * Y[0]));
        h[0] = fmaf(u[0], c0, fmaf(u[1], c1, u[2] * c2)) * rP;
        h[1] = fmaf(u[0], d0, fmaf(u[1], d1, u[2] * d2)) * rP;
        h[2] = fmaf(u[0], e0, fmaf(u[1], e1, u[2] * e2)) * rP;
        h[3] = fmaf(v[0], c0, fmaf(v[1], c1, v[2] * c2)) * rP;
        h[4] = fmaf(v[0], d0, fmaf(v[1], d1, v[2] * d2)) * rP;
        h[5] = fmaf(v[0], e0, fmaf(v[1], e1, v[2] * e2)) * rP;
        h[6] = h6; h[7] = h7;
        acc = 0;
        for (k = 0; k < 8; k++) acc = acc + fabsf(h[k]);
        return ok && (acc < INFINITY);
    }
}

static int is_inlier(const float *h, const float *q)
{
    float w = fmaf(h[6], q[0], fmaf(h[7], q[1], 1.0f));
    float ex = fmaf(-q[2], w, fmaf(h[0], q[0], fmaf(h[1], q[1], h[2])));
    float ey = fmaf(-q[3], w, fmaf(h[3], q[0], fmaf(h[4], q[1], h[5])));
    float e = fmaf(ex, ex, ey * ey);
    return fmaf(-w, w, e) <= 0.0f;
}

/*
 * One frame.  X,Y = image px, x,y = pitch m (float, n <= 64); table = K x 4 sample indices.
 * Outputs: best hypothesis index (-1 if none reaches 4 inliers), its inlier count and bit mask in the
 * normalised test, its normalised h[8] and the frame normalisation (so the caller can map it back).
 */
int orc_fixedk_frame(const float *X, const float *Y, const float *x, const float *y, int n, const uint8_t *table, int K,
                     float inv_thr, int *best_index, int *best_count, uint64_t *best_mask, float *best_h, float *norm_out)
{
    norm_t m;
    float q[64][4];
    int i, hh, best = 0, bi = -1;
    *best_index = -1; *best_count = 0; *best_mask = 0;
    if (n < 4 || n > 64) return 1;
    if (!frame_norm(X, Y, x, y, n, inv_thr, &m)) return 2;
    for (i = 0; i < n; i++) {
        q[i][0] = (X[i] - m.cX) * m.sX; q[i][1] = (Y[i] - m.cY) * m.sY;
        q[i][2] = (x[i] - m.cx) * m.rt; q[i][3] = (y[i] - m.cy) * m.rt;
    }
    for (hh = 0; hh < K; hh++) {
        const uint8_t *t = table + 4 * hh;
        float s[4][4], h[8];
        int c = 0, a, b, dup = 0;
        for (a = 0; a < 4; a++) { if (t[a] >= n) dup = 1; for (b = 0; b < a; b++) if (t[a] == t[b]) dup = 1; }
        if (dup) continue;
        for (a = 0; a < 4; a++) memcpy(s[a], q[t[a]], sizeof(s[a]));
        if (!hypothesis(s, h)) continue;
        for (i = 0; i < n; i++) c += is_inlier(h, q[i]);
        if (c > best) { best = c; bi = hh; memcpy(best_h, h, sizeof(h)); }
    }
    if (best > 3) {
        uint64_t mk = 0;
        for (i = 0; i < n; i++) if (is_inlier(best_h, q[i])) mk |= (uint64_t)1 << i;
        *best_index = bi; *best_count = best; *best_mask = mk;
    }
    norm_out[0] = m.cX; norm_out[1] = m.cY; norm_out[2] = m.sX; norm_out[3] = m.sY; norm_out[4] = m.cx; norm_out[5] = m.cy;
    return 0;
}

/* the kernel's counter-based sample generator, restated: a per-frame key (one SplitMix64 output of seed and frame)
 * split into two 32-bit keys; hypothesis hh hashes (hh ^ key) with a 32-bit mixer twice -> four 16-bit fields; field i
 * is scaled to [0, n - i) and selects the r-th index that has not been drawn yet */
static uint32_t mixer32(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    return x ^ (x >> 16);
}

void orc_seeded_table(uint64_t seed, uint64_t frame, int K, int n, uint8_t *table)
{
    uint64_t z = (seed ^ (0xD1B54A32D192ED03ull * (frame + 1))) + 0x9E3779B97F4A7C15ull;
    int hh, i, j;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    for (hh = 0; hh < K; hh++) {
        uint32_t w[2], field[4];
        int taken[4], cnt = 0;
        w[0] = mixer32((uint32_t)hh ^ (uint32_t)z);
        w[1] = mixer32((uint32_t)hh ^ (uint32_t)(z >> 32));
        field[0] = w[0] & 0xFFFFu; field[1] = w[0] >> 16; field[2] = w[1] & 0xFFFFu; field[3] = w[1] >> 16;
        for (i = 0; i < 4; i++) {
            /* r-th free index: walk the indices in increasing order, skipping the ones already taken */
            int r = (int)((field[i] * (uint32_t)(n - i)) >> 16), v;
            for (v = 0; v < n; v++) {
                int used = 0;
                for (j = 0; j < cnt; j++) used |= (taken[j] == v);
                if (used) continue;
                if (r == 0) break;
                r--;
            }
            taken[cnt++] = v;
            table[4 * hh + i] = (uint8_t)v;
        }
    }
}
