// The RHO and LMEDS legs of the reference's homography cascade
// (eagle/models/coordinate_model.py:354-357: `for method in [cv2.RANSAC, cv2.RHO, cv2.LMEDS]`), as
// __host__ __device__ scalar code: cascade_kernel (fit.cu) runs it for the frames the RANSAC leg gave up on,
// tests/native/host_check.cpp compiles the same text with g++ for the CPU suite.
//
// What is restated (opencv-python 4.13 calib3d; no source in the image, see oracle/rho.py for the pinning):
//   RHO   rho.cpp RHO_HEST_REFC::rhoHest with flags NR | FINAL_REFINEMENT, maxD = 3, maxI = rConvg = 2000,
//         cfd = 0.995, minInl = 4, beta = 0.35 (fundam.cpp createAndRunRHORegistrator): PROSAC sampling from an
//         xorshift128+ stream, the "common coordinate" and orientation sample tests, the float32 4-point solve,
//         SPRT evaluation, iteration / non-randomness bounds, float32 Levenberg-Marquardt polish (damped Cholesky).
//         The result is the float32 H and the inlier flags of the best SAMPLE model (the polish does not touch them).
//   LMEDS ptsetreg.cpp LMeDSPointSetRegistrator::run (55 samples from cv::RNG, median of the float errors, the
//         2.5*1.4826*(1+5/(n-4))*sqrt(med) inlier band) followed by findHomography's common tail
//         (refit_on_inliers), final mask at the default reprojection threshold 3.
// Float32 operations are single rounded operations in the library's order (fmul/fadd/... of geometry_core.cuh);
// rho_hfunc.inc / rho_lmstep.inc are generated statement lists (tools/cv2_probe/gen_inc.py).
#pragma once

#include "geometry_core.cuh"

namespace egl {

constexpr int kMaxPtsCascade = 64;  // = kMaxPts (common.cuh), which the host build does not include
constexpr int kLmedsMaxIters = 64;

// ---- RHO_HEST_REFC::fastSeed / fastRandom -------------------------------------------------------------------------
struct XorShift128Plus {
    uint64_t s0, s1;
    EGL_HD uint64_t next() {
        uint64_t x = s0;
        const uint64_t y = s1;
        x ^= x << 23;
        x ^= x >> 17;
        x ^= y ^ (y >> 26);
        s0 = y;
        s1 = x;
        return x + y;
    }
    EGL_HD void seed(uint64_t v) {
        s0 = v;
        s1 = ~v;
        for (int i = 0; i < 20; ++i) next();
    }
    EGL_HD double uniform() {
#if defined(__CUDA_ARCH__)
        return dmul(__ull2double_rn(next()), 5.421010862427522e-20);
#else
        return (double)next() * 5.421010862427522e-20;
#endif
    }
};

// rndSmpl: selection sampling when sampleSize*2 > dataSetSize, otherwise draws until distinct.
EGL_HD void rho_rnd_smpl(XorShift128Plus& rng, int sample_size, int data_size, int* out) {
    if (sample_size * 2 > data_size) {
        int j = 0;
        for (int i = 0; j < sample_size; ++i) {
            const double u = rng.uniform();
            if (dmul((double)(data_size - i), u) < (double)(sample_size - j)) out[j++] = i;
        }
    } else {
        for (int i = 0; i < sample_size; ++i) {
            bool dup;
            do {
                out[i] = (int)(unsigned)(long long)dmul(rng.uniform(), (double)data_size);
                dup = false;
                for (int j = 0; j < i; ++j) dup |= out[i] == out[j];
            } while (dup);
        }
    }
}

// (int) of a float the way cvttss2si does it: truncation, INT_MIN when out of range or NaN.
EGL_HD int rho_f2i(float v) { return (v > -2147483648.f && v < 2147483648.f) ? (int)v : (int)0x80000000; }

EGL_HD float rho_side(const float* P, int o, int a, int b, int c) {  // ((a x b) . c) of the points at P[o + 2k], P[o + 2k + 1]
    const float xa = P[o + 2 * a], ya = P[o + 2 * a + 1], xb = P[o + 2 * b], yb = P[o + 2 * b + 1];
    const float c0 = fsub(ya, yb), c1 = fsub(xb, xa), c2 = fsub(fmul(xa, yb), fmul(xb, ya));
    return fadd(fadd(fmul(P[o + 2 * c], c0), fmul(P[o + 2 * c + 1], c1)), c2);
}

// isSampleDegenerate on P = {x0,y0..x3,y3 (image), X0,Y0..X3,Y3 (pitch)}
EGL_HD bool rho_sample_degenerate(const float* P) {
    for (int a = 0; a < 4; ++a)
        for (int b = a + 1; b < 4; ++b)
            if (P[2 * a] == P[2 * b] || P[2 * a + 1] == P[2 * b + 1]) return true;
    if ((rho_f2i(rho_side(P, 0, 0, 1, 2)) ^ rho_f2i(rho_side(P, 8, 0, 1, 2))) < 0) return true;
    if ((rho_f2i(rho_side(P, 0, 0, 1, 3)) ^ rho_f2i(rho_side(P, 8, 0, 1, 3))) < 0) return true;
    if ((rho_f2i(rho_side(P, 0, 2, 3, 0)) ^ rho_f2i(rho_side(P, 8, 2, 3, 0))) < 0) return true;
    if ((rho_f2i(rho_side(P, 0, 2, 3, 1)) ^ rho_f2i(rho_side(P, 8, 2, 3, 1))) < 0) return true;
    return false;
}

// hFuncRefC: float32 homography through four correspondences; H[8] = 1.
EGL_HD_NOINLINE void rho_hfunc(const float* P, float* H) {
#include "rho_hfunc.inc"
}

// designSPRTTest / sacDesignSPRTTest
struct RhoSprt {
    double A, lambda_accept, lambda_reject;
};
EGL_HD_NOINLINE void rho_design_sprt(double delta, double eps, RhoSprt* s) {
    const double acc = delta / eps, rej = (1.0 - delta) / (1.0 - eps);
    const double C = dadd(dmul(1.0 - delta, log(rej)), dmul(delta, log(acc)));
    const double K = dadd(dmul(C, 25.0) / 1.0, 1.0);
    double An = K;
    for (int i = 0; i < 10; ++i) {
        const double prev = An;
        An = dadd(K, log(An));
        if (!(An - prev > 1.5e-8)) break;
    }
    s->A = An;
    s->lambda_accept = acc;
    s->lambda_reject = rej;
}

// sacCalcIterBound
EGL_HD_NOINLINE unsigned rho_iter_bound(double cfd, double inlier_rate, unsigned max_bound) {
    const double p = 1.0 - pow(inlier_rate, 4.0);
    unsigned ret;
    if (p >= 1.0)
        ret = max_bound;
    else if (p <= 0.0)
        ret = 1;
    else
        ret = (unsigned)(long long)ceil(log(1.0 - cfd) / log(p));
    return ret <= max_bound ? ret : max_bound;
}

// sacInitNonRand entry n (beta = 0.35, CHI_SQ = 1.645)
EGL_HD unsigned rho_nonrand_min_inliers(int n) {
    const double k = dmul(sqrt(dmul(0.35, 1.0 - 0.35)), 1.645);
    return (unsigned)ceil(dadd(dadd(4.0, dmul((double)n, 0.35)), dmul(sqrt((double)n), k)));
}

// sacCalcJacobianErrors: float32 sums over the inliers; JtJ lower triangle.  JtJ/Jte may be null (error only).
EGL_HD_NOINLINE float rho_jacobian_errors(const float* H, const float* sx, const float* sy, const float* dx, const float* dy, int N,
                                          uint64_t inl, float (*JtJ)[8], float* Jte) {
    if (JtJ) {
        for (int i = 0; i < 8; ++i)
            for (int j = 0; j < 8; ++j) JtJ[i][j] = 0.f;
        for (int i = 0; i < 8; ++i) Jte[i] = 0.f;
    }
    float S = 0.f;
    for (int i = 0; i < N; ++i) {
        if (!((inl >> i) & 1ull)) continue;
        const float x = sx[i], y = sy[i], X = dx[i], Y = dy[i];
        const float W = fadd(fadd(fmul(H[6], x), fmul(H[7], y)), 1.0f);
        const float iW = fabsf(W) > FLT_EPSILON ? fdiv(1.0f, W) : 0.f;
        const float rx = fmul(fadd(fadd(fmul(H[0], x), fmul(H[1], y)), H[2]), iW);
        const float ry = fmul(fadd(fadd(fmul(H[3], x), fmul(H[4], y)), H[5]), iW);
        const float eX = fsub(rx, X), eY = fsub(ry, Y);
        S = fadd(S, fadd(fmul(eX, eX), fmul(eY, eY)));
        if (!JtJ) continue;
        const float d11 = fmul(x, iW), d12 = fmul(y, iW), d13 = iW;
        const float d31x = fmul(fmul(-rx, x), iW), d32x = fmul(fmul(-rx, y), iW);
        const float d31y = fmul(fmul(-ry, x), iW), d32y = fmul(fmul(-ry, y), iW);
#define EGL_ACC(v, e) v = fadd(v, e)
        EGL_ACC(Jte[0], fmul(eX, d11)); EGL_ACC(Jte[1], fmul(eX, d12)); EGL_ACC(Jte[2], fmul(eX, d13));
        EGL_ACC(Jte[3], fmul(eY, d11)); EGL_ACC(Jte[4], fmul(eY, d12)); EGL_ACC(Jte[5], fmul(eY, d13));
        EGL_ACC(Jte[6], fadd(fmul(eX, d31x), fmul(eY, d31y)));
        EGL_ACC(Jte[7], fadd(fmul(eX, d32x), fmul(eY, d32y)));
        EGL_ACC(JtJ[0][0], fmul(d11, d11));
        EGL_ACC(JtJ[1][0], fmul(d11, d12)); EGL_ACC(JtJ[1][1], fmul(d12, d12));
        EGL_ACC(JtJ[2][0], fmul(d11, d13)); EGL_ACC(JtJ[2][1], fmul(d12, d13)); EGL_ACC(JtJ[2][2], fmul(d13, d13));
        EGL_ACC(JtJ[3][3], fmul(d11, d11));
        EGL_ACC(JtJ[4][3], fmul(d11, d12)); EGL_ACC(JtJ[4][4], fmul(d12, d12));
        EGL_ACC(JtJ[5][3], fmul(d11, d13)); EGL_ACC(JtJ[5][4], fmul(d12, d13)); EGL_ACC(JtJ[5][5], fmul(d13, d13));
        EGL_ACC(JtJ[6][0], fmul(d11, d31x)); EGL_ACC(JtJ[6][1], fmul(d12, d31x)); EGL_ACC(JtJ[6][2], fmul(d13, d31x));
        EGL_ACC(JtJ[6][3], fmul(d11, d31y)); EGL_ACC(JtJ[6][4], fmul(d12, d31y)); EGL_ACC(JtJ[6][5], fmul(d13, d31y));
        EGL_ACC(JtJ[6][6], fadd(fmul(d31x, d31x), fmul(d31y, d31y)));
        EGL_ACC(JtJ[7][0], fmul(d11, d32x)); EGL_ACC(JtJ[7][1], fmul(d12, d32x)); EGL_ACC(JtJ[7][2], fmul(d13, d32x));
        EGL_ACC(JtJ[7][3], fmul(d11, d32y)); EGL_ACC(JtJ[7][4], fmul(d12, d32y)); EGL_ACC(JtJ[7][5], fmul(d13, d32y));
        EGL_ACC(JtJ[7][6], fadd(fmul(d31x, d32x), fmul(d31y, d32y)));
        EGL_ACC(JtJ[7][7], fadd(fmul(d32x, d32x), fmul(d32y, d32y)));
#undef EGL_ACC
    }
    return S;
}

// sacChol8x8Damped: Cholesky factor of A with the diagonal scaled by (1 + lambda); false when a pivot goes negative.
EGL_HD_NOINLINE bool rho_chol_damped(const float (*A)[8], float lambda, float (*L)[8]) {
    const float lp1 = fadd(lambda, 1.0f);
    for (int i = 0; i < 8; ++i) {
        for (int j = 0; j < i; ++j) {
            float x = A[i][j];
            for (int k = 0; k < j; ++k) x = fsub(x, fmul(L[i][k], L[j][k]));
            L[i][j] = fdiv(x, L[j][j]);
        }
        float x = fmul(A[i][i], lp1);
        for (int k = 0; k < i; ++k) x = fsub(x, fmul(L[i][k], L[i][k]));
        if (x < 0) return false;
        L[i][i] = sqrtf(x);
    }
    return true;
}

// sacTRInv8x8 + sacTRISolve8x8 + sacSub8x1: dH = L^-T (L^-1 Jte), newH = H - dH
EGL_HD_NOINLINE void rho_lm_step(const float (*L)[8], const float* Jte, const float* H, float* newH, float* dH) {
#include "rho_lmstep.inc"
}

// RHO_HEST_REFC::refine: <= 100 LM iterations on H[0..7] (H[8] stays 1)
EGL_HD_NOINLINE void rho_refine(float* H, const float* sx, const float* sy, const float* dx, const float* dy, int N, uint64_t inl) {
    float JtJ[8][8], Lc[8][8], Jte[8], dH[8], newH[9];
    float lam = 100.0f;
    float S = rho_jacobian_errors(H, sx, sy, dx, dy, N, inl, JtJ, Jte);
    newH[8] = H[8];
    for (int it = 0; it < 100; ++it) {
        while (!rho_chol_damped(JtJ, lam, Lc)) lam = fmul(lam, 2.0f);
        rho_lm_step(Lc, Jte, H, newH, dH);
        const float newS = rho_jacobian_errors(newH, sx, sy, dx, dy, N, inl, nullptr, nullptr);
        const float dS = fsub(S, newS);
        float sq = fadd(fmul(dH[0], dH[0]), 0.f);
        for (int i = 1; i < 8; ++i) sq = fadd(sq, fmul(dH[i], dH[i]));
        float dL = fadd(fmul(dH[0], Jte[0]), fmul(sq, lam));
        for (int i = 1; i < 8; ++i) dL = fadd(dL, fmul(Jte[i], dH[i]));
        dL = fmul(dL, 0.5f);
        const float gain = FLT_EPSILON > fabsf(dL) ? dS : fdiv(dS, dL);
        if (gain < 0.25f) {
            lam = fmul(lam, 8.0f);
            if (lam > 8388608000.0f) break;  // 1000 / FLT_EPSILON
        } else if (gain > 0.75f) {
            lam = fmul(lam, 0.5f);
        }
        if (gain > 0) {
            for (int i = 0; i < 8; ++i) H[i] = newH[i];
            S = rho_jacobian_errors(H, sx, sy, dx, dy, N, inl, JtJ, Jte);
        }
    }
}

// rhoHest as findHomography(img, pitch, RHO) calls it.  Returns the inlier count of the best sample model (0 = no
// model: fewer than 4 inliers); H (9 floats) and *mask (bit i = point i) are valid when it is > 0.  N >= 5 here
// (findHomography handles N == 4 without a robust method).
EGL_HD_NOINLINE int rho_fit(const float* sx, const float* sy, const float* dx, const float* dy, int N, float* H, uint64_t* mask) {
    XorShift128Plus rng;
    rng.seed(~0ull);
    unsigned maxI = 2000;
    unsigned phNum = 4, phEndI = 1, phMax = (unsigned)N, phNumInl = 0;
    double phEndFpI = dmul(2000.0, 24.0) / dmul(dmul(dmul((double)(N - 1), (double)N), (double)(N - 2)), (double)(N - 3));
    double eps = 0.1, delta = 0.01;
    RhoSprt sprt;
    rho_design_sprt(delta, eps, &sprt);
    const float maxD2 = 9.0f;
    float bestH[9];
    uint64_t best_inl = 0;
    unsigned best_num = 0;
    for (int i = 0; i < 9; ++i) bestH[i] = 0.f;
    for (unsigned i = 0; i < maxI || i < 100; ++i) {
        if (i >= phEndI && phNum < phMax) {  // PROSAC: next phase
            ++phNum;
            const double next = dmul(phEndFpI, (double)phNum) / (double)(phNum - 4);
            phEndI += (unsigned)(long long)ceil(next - phEndFpI);
            phEndFpI = next;
        }
        int smpl[4];
        if (i > phEndI) {
            rho_rnd_smpl(rng, 4, (int)phNum, smpl);
        } else {
            rho_rnd_smpl(rng, 3, (int)phNum - 1, smpl);
            smpl[3] = (int)phNum - 1;
        }
        float P[16];
        for (int k = 0; k < 4; ++k) {
            P[2 * k] = sx[smpl[k]];
            P[2 * k + 1] = sy[smpl[k]];
            P[8 + 2 * k] = dx[smpl[k]];
            P[8 + 2 * k + 1] = dy[smpl[k]];
        }
        if (rho_sample_degenerate(P)) continue;
        float Hc[9];
        rho_hfunc(P, Hc);
        const float hs = fadd(fadd(fadd(fadd(fadd(fadd(fadd(Hc[0], Hc[1]), Hc[2]), Hc[3]), Hc[4]), Hc[5]), Hc[6]), Hc[7]);
        if (hs != hs) continue;
        // evaluateModelSPRT
        uint64_t inl = 0;
        unsigned num = 0, tested = 0;
        double lambda = 1.0;
        bool good = true;
        for (int k = 0; k < N && good; ++k) {
            const float x = sx[k], y = sy[k];
            float rx = fadd(fadd(fmul(Hc[0], x), fmul(Hc[1], y)), Hc[2]);
            float ry = fadd(fadd(fmul(Hc[3], x), fmul(Hc[4], y)), Hc[5]);
            const float rz = fadd(fadd(fmul(Hc[6], x), fmul(Hc[7], y)), 1.0f);
            rx = fsub(fdiv(rx, rz), dx[k]);
            ry = fsub(fdiv(ry, rz), dy[k]);
            const float d = fadd(fmul(rx, rx), fmul(ry, ry));
            const bool isin = maxD2 >= d;
            num += isin;
            inl |= (uint64_t)isin << k;
            lambda = dmul(lambda, isin ? sprt.lambda_accept : sprt.lambda_reject);
            good = sprt.A >= lambda;
            tested = (unsigned)k + 1;
        }
        // updateSPRT
        if (good) {
            if (num > best_num) {
                eps = (double)num / (double)N;
                rho_design_sprt(delta, eps, &sprt);
            }
        } else {
            const double nd = (double)num / (double)tested;
            if (nd > 0 && fabs(delta - nd) / delta > 0.1) {
                delta = nd;
                rho_design_sprt(delta, eps, &sprt);
            }
        }
        if (num <= best_num) continue;
        // saveBestModel, updateBounds, nStarOptimize
        for (int k = 0; k < 9; ++k) bestH[k] = Hc[k];
        best_inl = inl;
        best_num = num;
        maxI = rho_iter_bound(0.995, (double)best_num / (double)N, maxI);
        unsigned best_n = (unsigned)N, bn = best_num, test_n = (unsigned)N, tn = best_num;
        for (; test_n > 20 && tn; --test_n) {
            if (tn * best_n > bn * test_n) {
                if (tn < rho_nonrand_min_inliers((int)test_n)) break;
                best_n = test_n;
                bn = tn;
            }
            tn -= (unsigned)((best_inl >> (test_n - 1)) & 1ull);
        }
        if (bn * phMax > phNumInl * best_n) {
            phMax = best_n;
            phNumInl = bn;
            maxI = rho_iter_bound(0.995, (double)phNumInl / (double)phMax, maxI);
        }
    }
    if (best_num > 4) rho_refine(bestH, sx, sy, dx, dy, N, best_inl);
    if (best_num < 4) return 0;
    for (int k = 0; k < 9; ++k) H[k] = bestH[k];
    *mask = best_inl;
    return (int)best_num;
}

// ---- LMEDS ---------------------------------------------------------------------------------------------------------
// Key that orders float errors the way std::nth_element on the int view does (x86 default NaN has the sign bit set).
EGL_HD int lmeds_key(float e) {
    if (e != e) return (int)0xFFC00000;
#if defined(__CUDA_ARCH__)
    return __float_as_int(e);
#else
    union { float f; int i; } u;
    u.f = e;
    return u.i;
#endif
}

// LMeDSPointSetRegistrator::run + the common tail of findHomography, in three pieces so that cascade_kernel can rate the
// samples on the lanes of a warp (the sample SEQUENCE is a walk of cv::RNG, the rating of a sample depends on nothing
// else) while the host build and lmeds_fit() below run them one after the other.
//
// 1. the sample sequence: up to `niters` (55 at OpenCV's defaults) accepted 4-subsets; -1 = the leg fails at once
EGL_HD_NOINLINE int lmeds_draw_samples(const float* sx, const float* sy, const float* dx, const float* dy, int N, double confidence,
                                       uint8_t (*smp)[4]) {
    CvRng rng{~0ull};
    int niters = ransac_update_num_iters(confidence, 0.45, 4, 2000);
    if (niters < 3) niters = 3;
    if (niters > kLmedsMaxIters) niters = kLmedsMaxIters;
    int n_smp = 0;
    for (int iter = 0; iter < niters; ++iter) {
        int idx[4];
        bool found = false;
        for (int attempt = 0; attempt < 1000 && !found; ++attempt) {  // getSubset default maxAttempts (the RANSAC leg passes 10000)
            draw_subset(rng, N, idx);
            float qx[4], qy[4], rx[4], ry[4];
            for (int k = 0; k < 4; ++k) {
                qx[k] = sx[idx[k]];
                qy[k] = sy[idx[k]];
                rx[k] = dx[idx[k]];
                ry[k] = dy[idx[k]];
            }
            found = check_subset(qx, qy, rx, ry);
        }
        if (!found) {
            if (iter == 0) return -1;
            break;
        }
        for (int k = 0; k < 4; ++k) smp[n_smp][k] = (uint8_t)idx[k];
        ++n_smp;
    }
    return n_smp;
}

// 2. one sample: runKernel as OpenCV runs it (normal matrix + its Jacobi eigen-decomposition: the median decides by a
//    rounding) and element N/2 of the sorted float errors (std::nth_element on the int view).  false = no model.
EGL_HD_NOINLINE bool lmeds_rate_sample(const float* sx, const float* sy, const float* dx, const float* dy, int N, const uint8_t* smp4,
                                       double* Hm, double* median, double* scratch) {
    float qx[4], qy[4], rx[4], ry[4];
    for (int k = 0; k < 4; ++k) {
        qx[k] = sx[smp4[k]];
        qy[k] = sy[smp4[k]];
        rx[k] = dx[smp4[k]];
        ry[k] = dy[smp4[k]];
    }
    if (!run_kernel_ls(qx, qy, rx, ry, nullptr, 4, Hm, scratch, true)) return false;
    float Hf[8];
    for (int i = 0; i < 8; ++i) Hf[i] = (float)Hm[i];
    int key[kMaxPtsCascade];
    float err[kMaxPtsCascade];
    for (int i = 0; i < N; ++i) {
        err[i] = reproj_err_f32(Hf, sx[i], sy[i], dx[i], dy[i]);
        key[i] = lmeds_key(err[i]);
    }
    const int want = N / 2;
    float med = 0.f;
    for (int i = 0; i < N; ++i) {
        int less = 0, equal = 0;
        for (int j = 0; j < N; ++j) {
            less += key[j] < key[i];
            equal += key[j] == key[i];
        }
        if (less <= want && want < less + equal) {
            med = err[i];
            break;
        }
    }
    *median = (double)med;
    return true;
}

// 3. from the model of the smallest median: inlier band, refit on the band, final mask at the default threshold 3.
//    Returns the final inlier count (may be 0: cv2 still returns the refined H), or -1 when the band holds < 4 points.
EGL_HD_NOINLINE int lmeds_finish(const float* sx, const float* sy, const float* dx, const float* dy, int N, const double* best,
                                 double min_median, double* H, uint64_t* mask, double* scratch, uint64_t* band_mask = nullptr) {
    double sigma = dmul(dmul(dmul(2.5, 1.4826), dadd(1.0, 5.0 / (double)(N - 4))), sqrt(min_median));
    if (!(sigma > 0.001)) sigma = 0.001;   // MAX(sigma, 0.001); a NaN sigma also ends up here, as MAX's `a > b ? a : b` does
    uint64_t pm;
    const int good = inlier_mask_f32(best, sx, sy, dx, dy, N, (float)dmul(sigma, sigma), &pm);
    if (band_mask) *band_mask = pm;
    if (good < 4) return -1;
    for (int i = 0; i < 9; ++i) H[i] = best[i];
    // the band may hold points that fit no common model, so the LM system has no margin over OpenCV's eigenvalue
    // threshold at any set size: always solve the LM steps the way cv::solve(DECOMP_EIG) does
    return refit_on_inliers(H, sx, sy, dx, dy, N, pm, 9.0f, mask, scratch, true);
}

// The whole leg on one thread.  N >= 5.  Returns the final inlier count or -1 (no model); H (double, h33 = 1) and *mask
// as for the RANSAC leg.  scratch >= 192 doubles.
EGL_HD_NOINLINE int lmeds_fit(const float* sx, const float* sy, const float* dx, const float* dy, int N, double confidence,
                              double* H, uint64_t* mask, double* scratch, uint64_t* band_mask = nullptr) {
    uint8_t smp[kLmedsMaxIters][4];
    const int n_smp = lmeds_draw_samples(sx, sy, dx, dy, N, confidence, smp);
    if (n_smp < 0) return -1;
    double min_median = DBL_MAX;
    double best[9];
    bool have = false;
    for (int t = 0; t < n_smp; ++t) {
        double Hm[9], median;
        if (!lmeds_rate_sample(sx, sy, dx, dy, N, smp[t], Hm, &median, scratch)) continue;
        if (median < min_median) {  // the first sample of the smallest median wins
            min_median = median;
            for (int i = 0; i < 9; ++i) best[i] = Hm[i];
            have = true;
        }
    }
    if (!have) return -1;
    return lmeds_finish(sx, sy, dx, dy, N, best, min_median, H, mask, scratch, band_mask);
}

}  // namespace egl
