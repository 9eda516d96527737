// Scalar building blocks of the homography fit and the projection, shared by every kernel.
//
// Everything here is a pure function of its arguments, written so that the operation order is
// fixed (explicit round-to-nearest intrinsics where a fused multiply-add would change a result
// that has to match OpenCV bit for bit).  The functions are __host__ __device__: the kernels in
// fit.cu / project.cu call them on the GPU, and tests/native/host_check.cpp compiles the very
// same code with g++ (-ffp-contract=off) so the arithmetic can be checked against the oracle on
// a machine without a GPU.  That host build is a test artefact only; the product never calls it.
//
// OpenCV references are to calib3d (fundam.cpp, ptsetreg.cpp, levmarq.cpp) of opencv-python
// 4.11/4.13, the library behind the reference's cv2.findHomography call
// (eagle/models/coordinate_model.py:355).
#pragma once

#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define EGL_HD __host__ __device__ __forceinline__
#define EGL_HD_NOINLINE static __host__ __device__ __noinline__
#define EGL_UNROLL _Pragma("unroll")
#else
#define EGL_UNROLL
#define EGL_HD inline
#define EGL_HD_NOINLINE static inline
#endif

namespace egl {

// ---- non-contracted float arithmetic (OpenCV's x86 build has no FMA in computeError) ----------
#if defined(__CUDA_ARCH__)
EGL_HD float fmul(float a, float b) { return __fmul_rn(a, b); }
EGL_HD float fadd(float a, float b) { return __fadd_rn(a, b); }
EGL_HD float fsub(float a, float b) { return __fsub_rn(a, b); }
EGL_HD float fdiv(float a, float b) { return __fdiv_rn(a, b); }
EGL_HD double dmul(double a, double b) { return __dmul_rn(a, b); }
EGL_HD double dadd(double a, double b) { return __dadd_rn(a, b); }
EGL_HD double dsub(double a, double b) { return __dsub_rn(a, b); }
#else
EGL_HD float fmul(float a, float b) { return a * b; }  // host build uses -ffp-contract=off
EGL_HD float fadd(float a, float b) { return a + b; }
EGL_HD float fsub(float a, float b) { return a - b; }
EGL_HD float fdiv(float a, float b) { return a / b; }
EGL_HD double dmul(double a, double b) { return a * b; }
EGL_HD double dadd(double a, double b) { return a + b; }
EGL_HD double dsub(double a, double b) { return a - b; }
#endif

// ---- cv::RNG (multiply-with-carry); RANSACPointSetRegistrator::run seeds it with 2^64-1 -------
struct CvRng {
    uint64_t state;
    EGL_HD uint32_t next() {
        state = (uint64_t)(uint32_t)state * 4164903690ull + (state >> 32);
        return (uint32_t)state;
    }
    EGL_HD int uniform(int a, int b) { return a == b ? a : a + (int)(next() % (uint32_t)(b - a)); }
};

// getSubset(): 4 distinct indices by rejection.
EGL_HD void draw_subset(CvRng& rng, int count, int idx[4]) {
    for (int i = 0; i < 4; ++i) {
        int v;
        bool dup;
        do {
            v = rng.uniform(0, count);
            dup = false;
            for (int k = 0; k < i; ++k) dup |= (idx[k] == v);
        } while (dup);
        idx[i] = v;
    }
}

// haveCollinearPoints(ms, 4): only the last point is tested against the earlier pairs.
EGL_HD bool last_point_collinear(const float* px, const float* py) {
    const double xi = px[3], yi = py[3];
    for (int j = 0; j < 3; ++j) {
        const double dx1 = px[j] - xi, dy1 = py[j] - yi;
        for (int k = 0; k < j; ++k) {
            const double dx2 = px[k] - xi, dy2 = py[k] - yi;
            if (fabs(dsub(dmul(dx2, dy1), dmul(dy2, dx1))) <=
                (double)FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return true;
        }
    }
    return false;
}

// determinant of [[x0,y0,1],[x1,y1,1],[x2,y2,1]] in cv::determinant(Matx33d) order
EGL_HD double det_rows1(const float* px, const float* py, int a, int b, int c) {
    const double x0 = px[a], y0 = py[a], x1 = px[b], y1 = py[b], x2 = px[c], y2 = py[c];
    const double t0 = dmul(x0, dsub(y1, y2));
    const double t1 = dmul(y0, dsub(x1, x2));
    const double t2 = dsub(dmul(x1, y2), dmul(x2, y1));
    return dadd(dsub(t0, t1), t2);
}

// HomographyEstimatorCallback::checkSubset for a 4-point sample (src = image, dst = pitch).
EGL_HD bool check_subset(const float* sx, const float* sy, const float* dx, const float* dy) {
    if (last_point_collinear(sx, sy) || last_point_collinear(dx, dy)) return false;
    int negative = 0;
    negative += dmul(det_rows1(sx, sy, 0, 1, 2), det_rows1(dx, dy, 0, 1, 2)) < 0;
    negative += dmul(det_rows1(sx, sy, 1, 2, 3), det_rows1(dx, dy, 1, 2, 3)) < 0;
    negative += dmul(det_rows1(sx, sy, 0, 2, 3), det_rows1(dx, dy, 0, 2, 3)) < 0;
    negative += dmul(det_rows1(sx, sy, 0, 1, 3), det_rows1(dx, dy, 0, 1, 3)) < 0;
    return negative == 0 || negative == 4;
}

// ---- HomographyEstimatorCallback::computeError for one point, float, OpenCV's order ----------
EGL_HD float reproj_err_f32(const float* Hf, float X, float Y, float x, float y) {
    const float ww = fdiv(1.f, fadd(fadd(fmul(Hf[6], X), fmul(Hf[7], Y)), 1.f));
    const float ex = fsub(fmul(fadd(fadd(fmul(Hf[0], X), fmul(Hf[1], Y)), Hf[2]), ww), x);
    const float ey = fsub(fmul(fadd(fadd(fmul(Hf[3], X), fmul(Hf[4], Y)), Hf[5]), ww), y);
    return fadd(fmul(ex, ex), fmul(ey, ey));
}

// RANSACUpdateNumIters (ptsetreg.cpp)
EGL_HD int ransac_update_num_iters(double p, double ep, int model_points, int max_iters) {
    p = fmax(p, 0.);
    p = fmin(p, 1.);
    ep = fmax(ep, 0.);
    ep = fmin(ep, 1.);
    double num = fmax(1. - p, DBL_MIN);
    double denom = 1. - pow(1. - ep, (double)model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return (denom >= 0 || -num >= max_iters * (-denom)) ? max_iters : (int)rint(num / denom);
}

// ---- linear algebra on small dense matrices (double) ------------------------------------------

// Solve the n x n system a[n][n+1] (augmented, row-major, stride n+1) by Gaussian elimination with
// partial pivoting.  Returns false if a pivot is exactly zero / not finite.
template <int N>
EGL_HD bool gauss_solve(double* a, double* x) {
    constexpr int S = N + 1;
    for (int k = 0; k < N; ++k) {
        int piv = k;
        double best = fabs(a[k * S + k]);
        for (int i = k + 1; i < N; ++i) {
            const double v = fabs(a[i * S + k]);
            if (v > best) {
                best = v;
                piv = i;
            }
        }
        if (!(best > 0.0) || !isfinite(best)) return false;
        if (piv != k)
            for (int j = k; j < S; ++j) {
                const double t = a[k * S + j];
                a[k * S + j] = a[piv * S + j];
                a[piv * S + j] = t;
            }
        const double inv = 1.0 / a[k * S + k];
        for (int i = k + 1; i < N; ++i) {
            const double m = a[i * S + k] * inv;
            if (m != 0.0)
                for (int j = k + 1; j < S; ++j) a[i * S + j] -= m * a[k * S + j];
        }
    }
    for (int i = N - 1; i >= 0; --i) {
        double s = a[i * S + N];
        for (int j = i + 1; j < N; ++j) s -= a[i * S + j] * x[j];
        x[i] = s / a[i * S + i];
    }
    return true;
}

// Unit eigenvector of the SMALLEST eigenvalue of a symmetric positive semi-definite 9x9 matrix
// (row-major A, destroyed) by inverse iteration: one LU factorisation of A + sigma*I with a tiny
// relative shift (so that an exactly singular A -- noise-free correspondences -- still factors),
// then a few back-substitutions.  OpenCV takes the same vector from a full Jacobi
// eigen-decomposition (cv::eigen); the two agree to rounding whenever that eigenvalue is simple,
// and this costs ~1/20 of the work.  aug: 9*10 doubles of scratch.  Returns false for A == 0.
EGL_HD_NOINLINE bool smallest_eigvec9(const double* A, double* aug, double* h) {
    constexpr int n = 9, S = 10;
    double tr = 0;
    for (int i = 0; i < n; ++i) tr += A[i * n + i];
    if (!(tr > 0.0) || !isfinite(tr)) return false;
    const double sigma = tr * 1e-15;
    double y[n], z[n];
    for (int i = 0; i < n; ++i) y[i] = 1.0 / (1.37 + i);  // generic start, not orthogonal to anything special
    double prev = 1e300;
    for (int it = 0; it < 48; ++it) {
        // (re)build the augmented system: elimination is redone per iteration (n^3/3 = 243 FMAs),
        // which keeps the code to one routine; iterations needed: 2-5
        for (int i = 0; i < n; ++i) {
            for (int j = 0; j < n; ++j) aug[i * S + j] = A[i * n + j];
            aug[i * S + i] += sigma;
            aug[i * S + n] = y[i];
        }
        if (!gauss_solve<9>(aug, z)) return false;
        double nrm = 0, big = 0;
        int bi = 0;
        for (int i = 0; i < n; ++i) {
            nrm += z[i] * z[i];
            if (fabs(z[i]) > big) { big = fabs(z[i]); bi = i; }
        }
        if (!(nrm > 0.0) || !isfinite(nrm)) return false;
        const double sc = (z[bi] < 0 ? -1.0 : 1.0) / sqrt(nrm);
        double diff = 0;
        for (int i = 0; i < n; ++i) {
            z[i] *= sc;
            diff = fmax(diff, fabs(z[i] - y[i]));
            y[i] = z[i];
        }
        if (diff <= 1e-13 || (diff <= 1e-10 && diff >= prev)) break;  // converged (far below what the LM polish needs) / stagnated
        prev = diff;
    }
    for (int i = 0; i < n; ++i) h[i] = y[i];
    return true;
}

// cv::eigen for a symmetric 9x9 (core/lapack.cpp JacobiImpl_, the path an OpenCV build without Eigen takes): classical
// Jacobi with the max-off-diagonal pivot found through per-row / per-column index caches, OpenCV's own hypot, rotations
// in its order, eigenvalues sorted descending by selection.  Bit-identical to cv2.eigen on the host build
// (tests/test_host_core.py); used where the smallest eigenvalue of the DLT normal matrix is not isolated (LMEDS refits on
// an inlier band that holds no common model), so that the vector picked from the near-degenerate subspace is OpenCV's.
// A (81, destroyed), W (9), V (81: rows = eigenvectors).
EGL_HD double cv_hypot(double a, double b) {
    a = fabs(a);
    b = fabs(b);
    if (a > b) {
        b /= a;
        return dmul(a, sqrt(dadd(1.0, dmul(b, b))));
    }
    if (b > 0) {
        a /= b;
        return dmul(b, sqrt(dadd(1.0, dmul(a, a))));
    }
    return 0.0;
}
EGL_HD_NOINLINE void cv_jacobi9(double* A, double* W, double* V) {
    constexpr int n = 9;
    int indR[n], indC[n];
    for (int i = 0; i < n * n; ++i) V[i] = 0.0;
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    auto row_max = [&](int k) {
        int m = k + 1;
        double mv = fabs(A[k * n + m]);
        for (int i = k + 2; i < n; ++i) {
            const double val = fabs(A[k * n + i]);
            if (mv < val) { mv = val; m = i; }
        }
        indR[k] = m;
    };
    auto col_max = [&](int k) {
        int m = 0;
        double mv = fabs(A[k]);
        for (int i = 1; i < k; ++i) {
            const double val = fabs(A[i * n + k]);
            if (mv < val) { mv = val; m = i; }
        }
        indC[k] = m;
    };
    for (int k = 0; k < n; ++k) {
        W[k] = A[k * n + k];
        if (k < n - 1) row_max(k);
        if (k > 0) col_max(k);
    }
    for (int iters = 0; iters < n * n * 30; ++iters) {
        int k = 0;
        double mv = fabs(A[indR[0]]);
        for (int i = 1; i < n - 1; ++i) {
            const double val = fabs(A[i * n + indR[i]]);
            if (mv < val) { mv = val; k = i; }
        }
        int l = indR[k];
        for (int i = 1; i < n; ++i) {
            const double val = fabs(A[indC[i] * n + i]);
            if (mv < val) { mv = val; k = indC[i]; l = i; }
        }
        const double p = A[k * n + l];
        if (fabs(p) <= DBL_EPSILON) break;
        const double y = dmul(dsub(W[l], W[k]), 0.5);
        double t = dadd(fabs(y), cv_hypot(p, y));
        double s = cv_hypot(p, t);
        const double c = t / s;
        s = p / s;
        t = dmul(p / t, p);
        if (y < 0) { s = -s; t = -t; }
        A[k * n + l] = 0;
        W[k] = dsub(W[k], t);
        W[l] = dadd(W[l], t);
        auto rotate = [&](double& v0, double& v1) {
            const double a0 = v0, b0 = v1;
            v0 = dsub(dmul(a0, c), dmul(b0, s));
            v1 = dadd(dmul(a0, s), dmul(b0, c));
        };
        for (int i = 0; i < k; ++i) rotate(A[i * n + k], A[i * n + l]);
        for (int i = k + 1; i < l; ++i) rotate(A[k * n + i], A[i * n + l]);
        for (int i = l + 1; i < n; ++i) rotate(A[k * n + i], A[l * n + i]);
        for (int i = 0; i < n; ++i) rotate(V[k * n + i], V[l * n + i]);
        for (int j = 0; j < 2; ++j) {
            const int idx = j == 0 ? k : l;
            if (idx < n - 1) row_max(idx);
            if (idx > 0) col_max(idx);
        }
    }
    for (int k = 0; k < n - 1; ++k) {
        int m = k;
        for (int i = k + 1; i < n; ++i)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            double tmp = W[m]; W[m] = W[k]; W[k] = tmp;
            for (int i = 0; i < n; ++i) { tmp = V[m * n + i]; V[m * n + i] = V[k * n + i]; V[k * n + i] = tmp; }
        }
    }
}

// Minimum-norm solution of the singular symmetric system A d = v whose null vector is x (the
// 9-parameter homography cost is scale invariant, so J x = 0): solved through the bordered system
// [[A, s*x],[s*x^T, 0]] [d; mu] = [v; 0].  This is what cv::solve(..., DECOMP_EIG) returns for such
// an A (its back-substitution drops the zero eigenvalue).  aug: 10*11 doubles.
EGL_HD_NOINLINE bool solve_gauge_fixed9(const double* A, const double* x, const double* v, double* aug, double* d) {
    constexpr int n = 9, S = 11;
    double xn = 0, dmax = 0;
    for (int i = 0; i < n; ++i) { xn += x[i] * x[i]; dmax = fmax(dmax, fabs(A[i * n + i])); }
    if (!(xn > 0.0) || !(dmax > 0.0)) return false;
    const double s = dmax / sqrt(xn);
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) aug[i * S + j] = A[i * n + j];
        aug[i * S + n] = s * x[i];
        aug[i * S + n + 1] = v[i];
        aug[n * S + i] = s * x[i];
    }
    aug[n * S + n] = 0.0;
    aug[n * S + n + 1] = 0.0;
    double sol[10];
    if (!gauss_solve<10>(aug, sol)) return false;
    for (int i = 0; i < n; ++i) d[i] = sol[i];
    return true;
}

// ---- HomographyEstimatorCallback::runKernel: normalised DLT on n >= 4 correspondences ---------
// idx (may be null) selects the points; X,Y = image (src), x,y = pitch (dst) as float.
// Accumulates the 9x9 normal matrix with the block structure
//     LtL = sum_i (b b^T) (x) [[1,0,-x],[0,1,-y],[-x,-y,x^2+y^2]],  b = (X, Y, 1) normalised.
EGL_HD_NOINLINE bool dlt_normal_matrix(const float* sx, const float* sy, const float* dx, const float* dy,
                                       const uint8_t* idx, int n, double* LtL, double* norm /*[8]: cM,cm,sM,sm*/,
                                       bool exact = false) {
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
    for (int i = 0; i < n; ++i) {
        const int j = idx ? idx[i] : i;
        cmx += dx[j];
        cmy += dy[j];
        cMx += sx[j];
        cMy += sy[j];
    }
    cmx /= n; cmy /= n; cMx /= n; cMy /= n;
    double smx = 0, smy = 0, sMx = 0, sMy = 0;
    for (int i = 0; i < n; ++i) {
        const int j = idx ? idx[i] : i;
        smx += fabs(dx[j] - cmx);
        smy += fabs(dy[j] - cmy);
        sMx += fabs(sx[j] - cMx);
        sMy += fabs(sy[j] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return false;
    smx = n / smx; smy = n / smy; sMx = n / sMx; sMy = n / sMy;
    norm[0] = cMx; norm[1] = cMy; norm[2] = cmx; norm[3] = cmy;
    norm[4] = sMx; norm[5] = sMy; norm[6] = smx; norm[7] = smy;
    for (int i = 0; i < 81; ++i) LtL[i] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int j = idx ? idx[i] : i;
        const double x = dmul(dx[j] - cmx, smx), y = dmul(dy[j] - cmy, smy);
        const double X = dmul(sx[j] - cMx, sMx), Y = dmul(sy[j] - cMy, sMy);
        const double Lx[9] = {X, Y, 1, 0, 0, 0, dmul(-x, X), dmul(-x, Y), -x};
        const double Ly[9] = {0, 0, 0, X, Y, 1, dmul(-y, X), dmul(-y, Y), -y};
        if (exact) {  // no contraction: the library's x86 build rounds every product and sum
            for (int a = 0; a < 9; ++a)
                for (int b = a; b < 9; ++b)
                    LtL[a * 9 + b] = dadd(LtL[a * 9 + b], dadd(dmul(Lx[a], Lx[b]), dmul(Ly[a], Ly[b])));
        } else {
            for (int a = 0; a < 9; ++a)
                for (int b = a; b < 9; ++b) LtL[a * 9 + b] += Lx[a] * Lx[b] + Ly[a] * Ly[b];
        }
    }
    for (int a = 0; a < 9; ++a)
        for (int b = 0; b < a; ++b) LtL[a * 9 + b] = LtL[b * 9 + a];
    return true;
}

// H = invHnorm * H0 * Hnorm2, then scaled so that H[8] == 1 (runKernel's convertTo(1/H22)).
EGL_HD void dlt_denormalise(const double* h0, const double* norm, double* H) {
    const double cMx = norm[0], cMy = norm[1], cmx = norm[2], cmy = norm[3];
    const double sMx = norm[4], sMy = norm[5], smx = norm[6], smy = norm[7];
    // T = invHnorm * H0 ; invHnorm = [[1/smx,0,cmx],[0,1/smy,cmy],[0,0,1]]
    double T[9];
    for (int c = 0; c < 3; ++c) {
        T[0 * 3 + c] = (1.0 / smx) * h0[0 * 3 + c] + cmx * h0[2 * 3 + c];
        T[1 * 3 + c] = (1.0 / smy) * h0[1 * 3 + c] + cmy * h0[2 * 3 + c];
        T[2 * 3 + c] = h0[2 * 3 + c];
    }
    // H = T * Hnorm2 ; Hnorm2 = [[sMx,0,-cMx*sMx],[0,sMy,-cMy*sMy],[0,0,1]]
    for (int r = 0; r < 3; ++r) {
        H[r * 3 + 0] = T[r * 3 + 0] * sMx;
        H[r * 3 + 1] = T[r * 3 + 1] * sMy;
        H[r * 3 + 2] = T[r * 3 + 0] * (-cMx * sMx) + T[r * 3 + 1] * (-cMy * sMy) + T[r * 3 + 2];
    }
    const double s = 1.0 / H[8];
    for (int i = 0; i < 9; ++i) H[i] *= s;
}

// runKernel on the points selected by idx[0..n).  scratch: 81 + 90 doubles.
EGL_HD_NOINLINE bool run_kernel_ls(const float* sx, const float* sy, const float* dx, const float* dy,
                                   const uint8_t* idx, int n, double* H, double* scratch, bool exact = false) {
    double* LtL = scratch;
    double* aug = scratch + 81;
    double norm[8];
    if (!dlt_normal_matrix(sx, sy, dx, dy, idx, n, LtL, norm, exact)) return false;
    double h0[9];
    if (exact) {  // OpenCV's own eigen-decomposition (see cv_jacobi9); aug has room for the 81 vectors + 9 values
        cv_jacobi9(LtL, aug + 81, aug);
        for (int i = 0; i < 9; ++i) h0[i] = aug[8 * 9 + i];
    } else if (!smallest_eigvec9(LtL, aug, h0)) return false;
    dlt_denormalise(h0, norm, H);
    for (int i = 0; i < 9; ++i)
        if (!isfinite(H[i])) return false;
    return true;
}

// Minimal 4-point solve in double (the model cv2 evaluates per RANSAC iteration).  Same
// normalisation as runKernel; the 8x8 system with h33 = 1 in normalised coordinates is solved by
// block elimination in registers (left null vector of [X Y 1] -> 2x2 system for h31, h32 -> Cramer
// for the rest) instead of cv2's 9x9 eigen-decomposition: identical up to ~1e-13 relative, far
// below the float cast OpenCV applies before scoring.
EGL_HD double det3_f64(double Xa, double Ya, double Xb, double Yb, double Xc, double Yc) {
    return Xa * (Yb - Yc) - Ya * (Xb - Xc) + (Xb * Yc - Xc * Yb);
}

EGL_HD bool dlt4_f64(const float* sx, const float* sy, const float* dx, const float* dy, double* H) {
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0;
EGL_UNROLL
    for (int i = 0; i < 4; ++i) { cmx += dx[i]; cmy += dy[i]; cMx += sx[i]; cMy += sy[i]; }
    cmx /= 4; cmy /= 4; cMx /= 4; cMy /= 4;
    double smx = 0, smy = 0, sMx = 0, sMy = 0;
EGL_UNROLL
    for (int i = 0; i < 4; ++i) {
        smx += fabs(dx[i] - cmx); smy += fabs(dy[i] - cmy);
        sMx += fabs(sx[i] - cMx); sMy += fabs(sy[i] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON || fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return false;
    smx = 4 / smx; smy = 4 / smy; sMx = 4 / sMx; sMy = 4 / sMy;
    double X[4], Y[4], x[4], y[4];
EGL_UNROLL
    for (int i = 0; i < 4; ++i) {
        x[i] = (dx[i] - cmx) * smx; y[i] = (dy[i] - cmy) * smy;
        X[i] = (sx[i] - cMx) * sMx; Y[i] = (sy[i] - cMy) * sMy;
    }
    double n[4] = {det3_f64(X[1], Y[1], X[2], Y[2], X[3], Y[3]), -det3_f64(X[0], Y[0], X[2], Y[2], X[3], Y[3]),
                   det3_f64(X[0], Y[0], X[1], Y[1], X[3], Y[3]), -det3_f64(X[0], Y[0], X[1], Y[1], X[2], Y[2])};
EGL_UNROLL
    for (int k = 0; k < 3; ++k) {  // the row with the largest |n| goes to slot 3: rows 0..2 then have the largest 3x3 determinant
        const bool sw = fabs(n[k]) > fabs(n[3]);
        double t;
        t = n[3]; n[3] = sw ? n[k] : t; n[k] = sw ? t : n[k];
        t = X[3]; X[3] = sw ? X[k] : t; X[k] = sw ? t : X[k];
        t = Y[3]; Y[3] = sw ? Y[k] : t; Y[k] = sw ? t : Y[k];
        t = x[3]; x[3] = sw ? x[k] : t; x[k] = sw ? t : x[k];
        t = y[3]; y[3] = sw ? y[k] : t; y[k] = sw ? t : y[k];
    }
    double a11 = 0, a12 = 0, a21 = 0, a22 = 0, b1 = 0, b2 = 0;
EGL_UNROLL
    for (int k = 0; k < 4; ++k) {
        const double nx = n[k] * x[k], ny = n[k] * y[k];
        a11 -= nx * X[k]; a12 -= nx * Y[k]; b1 += nx;
        a21 -= ny * X[k]; a22 -= ny * Y[k]; b2 += ny;
    }
    const double Dt = a11 * a22 - a12 * a21;
    const double dP = det3_f64(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
    if (Dt == 0.0 || dP == 0.0) return false;
    const double rD = 1.0 / Dt, rP = 1.0 / dP;
    double h0[9];
    h0[6] = (b1 * a22 - a12 * b2) * rD;
    h0[7] = (a11 * b2 - b1 * a21) * rD;
    h0[8] = 1.0;
    double u[3], v[3];
EGL_UNROLL
    for (int k = 0; k < 3; ++k) {
        const double w = h0[6] * X[k] + h0[7] * Y[k] + 1.0;
        u[k] = x[k] * w; v[k] = y[k] * w;
    }
    const double c0 = Y[1] - Y[2], c1 = Y[2] - Y[0], c2 = Y[0] - Y[1];
    const double d0 = X[2] - X[1], d1 = X[0] - X[2], d2 = X[1] - X[0];
    const double e0 = X[1] * Y[2] - X[2] * Y[1], e1 = X[2] * Y[0] - X[0] * Y[2], e2 = X[0] * Y[1] - X[1] * Y[0];
    h0[0] = (u[0] * c0 + u[1] * c1 + u[2] * c2) * rP;
    h0[1] = (u[0] * d0 + u[1] * d1 + u[2] * d2) * rP;
    h0[2] = (u[0] * e0 + u[1] * e1 + u[2] * e2) * rP;
    h0[3] = (v[0] * c0 + v[1] * c1 + v[2] * c2) * rP;
    h0[4] = (v[0] * d0 + v[1] * d1 + v[2] * d2) * rP;
    h0[5] = (v[0] * e0 + v[1] * e1 + v[2] * e2) * rP;
    const double norm[8] = {cMx, cMy, cmx, cmy, sMx, sMy, smx, smy};
    dlt_denormalise(h0, norm, H);
    bool fin = true;
EGL_UNROLL
    for (int i = 0; i < 9; ++i) fin = fin && isfinite(H[i]);
    return fin;
}

// ---- HomographyRefineCallback + LMSolverImpl::run (9 parameters, <= 10 iterations) ------------
// Linearise at h: S = |r|^2, A = J^T J (9x9), v = J^T r, rmax = |r|_inf, over the points idx[0..n).
EGL_HD_NOINLINE void lm_linearise(const double* h, const float* sx, const float* sy, const float* dx, const float* dy,
                                  const uint8_t* idx, int n, double* A, double* v, double* S, double* rmax) {
    // A = sum (b b^T) (x) [[1,0,-xi],[0,1,-yi],[-xi,-yi,xi^2+yi^2]],  b = (Mx, My, 1) * ww
    double bb[6] = {0, 0, 0, 0, 0, 0}, bx[6] = {0, 0, 0, 0, 0, 0}, by[6] = {0, 0, 0, 0, 0, 0}, bq[6] = {0, 0, 0, 0, 0, 0};
    double vx[3] = {0, 0, 0}, vy[3] = {0, 0, 0}, vq[3] = {0, 0, 0};
    double s = 0, rm = 0;
    for (int i = 0; i < n; ++i) {
        const int j = idx ? idx[i] : i;
        const double Mx = sx[j], My = sy[j];
        double ww = h[6] * Mx + h[7] * My + h[8];
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        const double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
        const double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        const double rx = xi - dx[j], ry = yi - dy[j];
        s += rx * rx + ry * ry;
        rm = fmax(rm, fmax(fabs(rx), fabs(ry)));
        const double b[3] = {Mx * ww, My * ww, ww};
        const double q = xi * xi + yi * yi, g = xi * rx + yi * ry;
        int e = 0;
        for (int a = 0; a < 3; ++a) {
            for (int c = a; c < 3; ++c, ++e) {
                const double p = b[a] * b[c];
                bb[e] += p;
                bx[e] += p * xi;
                by[e] += p * yi;
                bq[e] += p * q;
            }
            vx[a] += b[a] * rx;
            vy[a] += b[a] * ry;
            vq[a] += b[a] * g;
        }
    }
    for (int a = 0; a < 3; ++a)
        for (int c = 0; c < 3; ++c) {
            const int lo = a < c ? a : c, hi = a < c ? c : a;
            const int e = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);  // packed upper-triangle index
            A[(0 + a) * 9 + (0 + c)] = bb[e];
            A[(3 + a) * 9 + (3 + c)] = bb[e];
            A[(0 + a) * 9 + (3 + c)] = 0.0;
            A[(3 + a) * 9 + (0 + c)] = 0.0;
            A[(0 + a) * 9 + (6 + c)] = -bx[e];
            A[(6 + a) * 9 + (0 + c)] = -bx[e];
            A[(3 + a) * 9 + (6 + c)] = -by[e];
            A[(6 + a) * 9 + (3 + c)] = -by[e];
            A[(6 + a) * 9 + (6 + c)] = bq[e];
        }
    for (int a = 0; a < 3; ++a) {
        v[a] = vx[a];
        v[3 + a] = vy[a];
        v[6 + a] = -vq[a];
    }
    *S = s;
    *rmax = rm;
}

EGL_HD double lm_cost(const double* h, const float* sx, const float* sy, const float* dx, const float* dy,
                      const uint8_t* idx, int n) {
    double s = 0;
    for (int i = 0; i < n; ++i) {
        const int j = idx ? idx[i] : i;
        const double Mx = sx[j], My = sy[j];
        double ww = h[6] * Mx + h[7] * My + h[8];
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        const double rx = (h[0] * Mx + h[1] * My + h[2]) * ww - dx[j];
        const double ry = (h[3] * Mx + h[4] * My + h[5]) * ww - dy[j];
        s += rx * rx + ry * ry;
    }
    return s;
}

// LMSolverImpl::run with maxIters = 10, eps = FLT_EPSILON, followed by the 1/h33 rescale of
// findHomography.  OpenCV solves each damped system with cv::solve(DECOMP_EIG); here the damped
// (positive definite) system goes through Gaussian elimination and the undamped, gauge-singular one
// through solve_gauge_fixed9 -- the same solutions up to rounding.  scratch: 81 + 110 doubles.
// Returns the iterations run.
// cv::solve / cv::invert with DECOMP_EIG, as LMSolverImpl::run uses them: eigen-decomposition of the symmetric
// 9x9 matrix, then back-substitution that DROPS every eigenvalue <= 2 * DBL_EPSILON * (sum of eigenvalues)
// (SVBkSb's threshold).  For a well-determined fit only the gauge direction of the 9-parameter cost falls under
// that threshold and the gauge-fixed elimination above gives the same answer; with few or nearly degenerate
// inliers J^T J (entries up to 1e16, eigenvalues down to O(1)) has further eigenvalues below it, cv2 silently
// treats the system as rank 7 or less, and only this routine follows it there.  Cyclic Jacobi in double.
//   d (may be null): pseudo-solution of A d = b;  maxdiag (may be null): max |diag(A^+)|
EGL_HD_NOINLINE void eig_threshold_solve9(const double* Ain, const double* b, double* d, double* maxdiag) {
    constexpr int n = 9;
    double A[n * n], V[n * n], w[n];
    for (int i = 0; i < n * n; ++i) { A[i] = Ain[i]; V[i] = 0.0; }
    for (int i = 0; i < n; ++i) V[i * n + i] = 1.0;
    for (int sweep = 0; sweep < 60; ++sweep) {
        double off = 0, diag = 0;
        for (int i = 0; i < n; ++i) {
            diag += A[i * n + i] * A[i * n + i];
            for (int j = i + 1; j < n; ++j) off += A[i * n + j] * A[i * n + j];
        }
        if (!(off > diag * 1e-36) || !isfinite(off)) break;
        for (int p = 0; p < n - 1; ++p)
            for (int q = p + 1; q < n; ++q) {
                const double apq = A[p * n + q];
                if (apq == 0.0) continue;
                const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
                const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                for (int k = 0; k < n; ++k) {  // columns p, q
                    const double akp = A[k * n + p], akq = A[k * n + q];
                    A[k * n + p] = c * akp - sn * akq;
                    A[k * n + q] = sn * akp + c * akq;
                }
                for (int k = 0; k < n; ++k) {  // rows p, q
                    const double apk = A[p * n + k], aqk = A[q * n + k];
                    A[p * n + k] = c * apk - sn * aqk;
                    A[q * n + k] = sn * apk + c * aqk;
                }
                for (int k = 0; k < n; ++k) {
                    const double vkp = V[k * n + p], vkq = V[k * n + q];
                    V[k * n + p] = c * vkp - sn * vkq;
                    V[k * n + q] = sn * vkp + c * vkq;
                }
            }
    }
    double sum = 0;
    for (int i = 0; i < n; ++i) { w[i] = A[i * n + i]; sum += w[i]; }
    const double thr = sum * (2.0 * DBL_EPSILON);
    if (d)
        for (int k = 0; k < n; ++k) d[k] = 0.0;
    double md = 0;
    double dg[n];
    for (int k = 0; k < n; ++k) dg[k] = 0.0;
    for (int i = 0; i < n; ++i) {
        if (!(w[i] > thr)) continue;
        const double inv = 1.0 / w[i];
        if (d) {
            double t = 0;
            for (int k = 0; k < n; ++k) t += V[k * n + i] * b[k];
            t *= inv;
            for (int k = 0; k < n; ++k) d[k] += V[k * n + i] * t;
        }
        for (int k = 0; k < n; ++k) dg[k] += V[k * n + i] * V[k * n + i] * inv;
    }
    for (int k = 0; k < n; ++k) md = fmax(md, fabs(dg[k]));
    if (maxdiag) *maxdiag = md;
}

// Inlier counts up to this take the eigen-thresholded solves in the undamped LM steps (see above; measured on
// 1500 hard synthetic fits the critical eigenvalue drops under OpenCV's threshold only with <= 8 inliers).
constexpr int kLmExactMaxPoints = 9;

EGL_HD_NOINLINE int lm_refine(double* H, const float* sx, const float* sy, const float* dx, const float* dy,
                              const uint8_t* idx, int n, double* scratch, bool always_exact = false) {
    const bool exact_small = always_exact || n <= kLmExactMaxPoints;
    double* A = scratch;         // J^T J
    double* aug = scratch + 81;  // 10 x 11 augmented system
    double x[9], xd[9], v[9], d[9], D[9];
    for (int i = 0; i < 9; ++i) x[i] = H[i];
    double S, rmax;
    lm_linearise(x, sx, sy, dx, dy, idx, n, A, v, &S, &rmax);
    for (int i = 0; i < 9; ++i) D[i] = A[i * 9 + i];
    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;
    int iter = 0;
    for (;;) {
        bool ok;
        if (lambda > 0 && always_exact) {  // cv::solve(DECOMP_EIG) on the damped system as well (see lmeds_fit)
            for (int i = 0; i < 81; ++i) aug[i] = A[i];
            for (int i = 0; i < 9; ++i) aug[i * 9 + i] += lambda * D[i];
            eig_threshold_solve9(aug, v, d, nullptr);
            ok = true;
        } else if (lambda > 0) {
            for (int i = 0; i < 9; ++i) {
                for (int j = 0; j < 9; ++j) aug[i * 10 + j] = A[i * 9 + j];
                aug[i * 10 + i] += lambda * D[i];
                aug[i * 10 + 9] = v[i];
            }
            ok = gauss_solve<9>(aug, d);
        } else if (exact_small) {
            eig_threshold_solve9(A, v, d, nullptr);
            ok = true;
        } else {
            ok = solve_gauge_fixed9(A, x, v, aug, d);
        }
        if (!ok)
            for (int i = 0; i < 9; ++i) d[i] = 0.0;
        for (int i = 0; i < 9; ++i) xd[i] = x[i] - d[i];
        const double Sd = lm_cost(xd, sx, sy, dx, dy, idx, n);
        double dS = 0;
        for (int i = 0; i < 9; ++i) {
            double t = 2.0 * v[i];
            for (int k = 0; k < 9; ++k) t -= A[i * 9 + k] * d[k];
            dS += d[i] * t;
        }
        const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            double t = 0;
            for (int i = 0; i < 9; ++i) t += d[i] * v[i];
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = fmin(fmax(nu, 2.), 10.);
            if (lambda == 0) {
                // invert(A, Ap, DECOMP_EIG): pseudo-inverse of the gauge-singular A; only the largest
                // diagonal entry is used.  Column k of A^+ = gauge-fixed solution for e_k.
                double maxval = DBL_EPSILON;
                if (exact_small) {
                    double md;
                    eig_threshold_solve9(A, nullptr, nullptr, &md);
                    maxval = fmax(maxval, md);
                } else {
                    for (int k = 0; k < 9; ++k) {
                        double e[9], col[9];
                        for (int i = 0; i < 9; ++i) e[i] = (i == k) ? 1.0 : 0.0;
                        if (solve_gauge_fixed9(A, x, e, aug, col)) maxval = fmax(maxval, fabs(col[k]));
                    }
                }
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }
        double dmax = 0;
        for (int i = 0; i < 9; ++i) dmax = fmax(dmax, fabs(d[i]));
        if (Sd < S) {
            S = Sd;
            for (int i = 0; i < 9; ++i) x[i] = xd[i];
            lm_linearise(x, sx, sy, dx, dy, idx, n, A, v, &S, &rmax);
        }
        iter++;
        const bool proceed = iter < 10 && dmax >= (double)FLT_EPSILON && rmax >= (double)FLT_EPSILON;
        if (!proceed) break;
    }
    const double sc = fabs(x[8]) > (double)FLT_EPSILON ? 1. / x[8] : 1.;  // fundam.cpp scaleFor()
    for (int i = 0; i < 9; ++i) H[i] = x[i] * sc;
    return iter;
}

// Inliers of H over the n points: float scoring in OpenCV's order.  Returns the count and sets
// bit i of *mask for inlier i (i = position in the point list).
EGL_HD int inlier_mask_f32(const double* H, const float* sx, const float* sy, const float* dx, const float* dy, int n,
                           float thr_sq, uint64_t* mask) {
    float Hf[8];
    for (int i = 0; i < 8; ++i) Hf[i] = (float)H[i];
    uint64_t m = 0;
    int c = 0;
    for (int i = 0; i < n; ++i) {
        const float e = reproj_err_f32(Hf, sx[i], sy[i], dx[i], dy[i]);
        if (e <= thr_sq) {
            m |= 1ull << i;
            ++c;
        }
    }
    *mask = m;
    return c;
}

// The tail of cv2.findHomography after RANSAC picked `H` (best model) with inlier list mask:
// runKernel on the inliers, LM polish, mask recomputed from the refined H.  scratch >= 192 doubles.
EGL_HD_NOINLINE int refit_on_inliers(double* H, const float* sx, const float* sy, const float* dx, const float* dy, int n,
                                     uint64_t ransac_mask, float thr_sq, uint64_t* final_mask, double* scratch,
                                     bool always_exact = false) {
    uint8_t idx[64];
    int m = 0;
    for (int i = 0; i < n; ++i)
        if ((ransac_mask >> i) & 1) idx[m++] = (uint8_t)i;
    double Hk[9];
    if (run_kernel_ls(sx, sy, dx, dy, idx, m, Hk, scratch, always_exact))
        for (int i = 0; i < 9; ++i) H[i] = Hk[i];
    lm_refine(H, sx, sy, dx, dy, idx, m, scratch, always_exact);
    return inlier_mask_f32(H, sx, sy, dx, dy, n, thr_sq, final_mask);
}

// ---- projection (cv::perspectiveTransform, float points, double matrix) ------------------------
EGL_HD void perspective_point(const double* H, float px, float py, float* ox, float* oy) {
    const double x = px, y = py;
    double w = dadd(dadd(dmul(x, H[6]), dmul(y, H[7])), H[8]);
    if (fabs(w) > DBL_EPSILON) {
        w = 1. / w;
        *ox = (float)dmul(dadd(dadd(dmul(x, H[0]), dmul(y, H[1])), H[2]), w);
        *oy = (float)dmul(dadd(dadd(dmul(x, H[3]), dmul(y, H[4])), H[5]), w);
    } else {
        *ox = 0.f;
        *oy = 0.f;
    }
}

// numpy float32 -> int64 .astype(int) on x86 (cvttss2si): non-finite / out of range -> INT64_MIN
EGL_HD long long trunc_like_numpy(float v) {
    if (!(v > -9.2233720368547758e18f && v < 9.2233720368547758e18f)) return (long long)0x8000000000000000ull;
    return (long long)v;
}

// find_x_at_y (coordinate_model.py:32-44) in Python-float semantics; ok=false where Python raises
// ZeroDivisionError (division by an exact zero).
EGL_HD double find_x_at_y(double x1, double y1, double x2, double y2, double y_target, bool* ok) {
    const double ddx = x2 - x1;
    if (ddx == 0.0) { *ok = false; return 0.0; }
    const double m = (y2 - y1) / ddx;
    const double c = dsub(y1, dmul(m, x1));
    if (m == 0.0) { *ok = false; return 0.0; }
    return (y_target - c) / m;
}

// ---- keypoint post-processing for one frame (coordinate_model.py:229-248) ----------------------
// flat/score: per-channel argmax and maximum.  Writes xy[57][2] for every channel, the kept
// channels in the reference's dict insertion order into order[], and returns their number.
//   kept      : score > 0.01 (keypoint_hrnet.py:592) and not score < conf (:232), compared in double
//               exactly like Python compares the float(score);
//   position  : xi = int((x / (w-1)) * img_w), yi likewise -- two double roundings, as in Python;
//   duplicates: channels sharing (xi, yi) keep the highest score; on an exact score tie the later
//               channel wins the label but takes the dict slot of the first tied channel.
EGL_HD_NOINLINE int postprocess_keypoints(const int32_t* flat, const float* score, int hm_h, int hm_w, int img_w,
                                          int img_h, double conf, int32_t* xy, uint8_t* order) {
    constexpr int C = 57;
    bool kept[C];
    const double wden = (double)(hm_w - 1 > 1 ? hm_w - 1 : 1), hden = (double)(hm_h - 1 > 1 ? hm_h - 1 : 1);
    for (int c = 0; c < C; ++c) {
        const int y = flat[c] / hm_w, x = flat[c] - y * hm_w;
        const double sc = (double)score[c];
        kept[c] = (sc > 0.01) && !(sc < conf);
        xy[2 * c] = (int32_t)dmul((double)x / wden, (double)img_w);
        xy[2 * c + 1] = (int32_t)dmul((double)y / hden, (double)img_h);
    }
    int n = 0;
    for (int c = 0; c < C; ++c) {
        if (!kept[c]) continue;
        float gmax = score[c];
        for (int o = 0; o < C; ++o)
            if (kept[o] && xy[2 * o] == xy[2 * c] && xy[2 * o + 1] == xy[2 * c + 1] && score[o] > gmax) gmax = score[o];
        if (score[c] != gmax) continue;  // a better channel owns this pixel
        bool first = true;
        int last = c;
        for (int o = 0; o < C; ++o) {
            if (!(kept[o] && xy[2 * o] == xy[2 * c] && xy[2 * o + 1] == xy[2 * c + 1] && score[o] == gmax)) continue;
            if (o < c) first = false;
            if (o > last) last = o;
        }
        if (first) order[n++] = (uint8_t)last;
    }
    return n;
}

// ---- fixed-K mode: sample generator and the FP32 minimal solve ------------------------------
// counter-based sample generator, branch free: a per-frame 64-bit key (one SplitMix64 output of seed and frame)
// gives two 32-bit keys; hypothesis h hashes (h ^ key) with a 32-bit integer mixer twice -> four 16-bit fields;
// field i is scaled to [0, N - i) and mapped to the r-th index not drawn yet ("skip" mapping over the sorted
// earlier picks), so the four indices are distinct and uniform over all ordered 4-subsets.  ~45 integer
// instructions, reproducible anywhere -- oracle/ransac_f32.c restates it.  Needs N >= 4.
EGL_HD uint64_t splitmix_mix(uint64_t s) {
    uint64_t z = s + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
EGL_HD uint32_t mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du;
    x ^= x >> 15; x *= 0x846ca68bu;
    return x ^ (x >> 16);
}
// i += (i >= a): one compare and one predicated add on the device
EGL_HD void bump_if_ge(int& i, int a) {
#if defined(__CUDA_ARCH__)
    asm("{\n.reg .pred p;\nsetp.ge.s32 p, %0, %1;\n@p add.s32 %0, %0, 1;\n}" : "+r"(i) : "r"(a));
#else
    i += (i >= a);
#endif
}
EGL_HD uint64_t seeded_frame_key(uint64_t seed, uint64_t frame) { return splitmix_mix(seed ^ (0xD1B54A32D192ED03ull * (frame + 1))); }
EGL_HD void seeded_subset_keyed(uint64_t key, uint32_t h, int N, int idx[4]) {
    const uint32_t z0 = mix32(h ^ (uint32_t)key), z1 = mix32(h ^ (uint32_t)(key >> 32));
    const int r0 = (int)(((z0 & 0xFFFFu) * (uint32_t)N) >> 16), r1 = (int)(((z0 >> 16) * (uint32_t)(N - 1)) >> 16);
    const int r2 = (int)(((z1 & 0xFFFFu) * (uint32_t)(N - 2)) >> 16), r3 = (int)(((z1 >> 16) * (uint32_t)(N - 3)) >> 16);
    const int i0 = r0;
    int i1 = r1, i2 = r2, i3 = r3;
    bump_if_ge(i1, i0);
    const int a = i0 < i1 ? i0 : i1, b = i0 < i1 ? i1 : i0;
    bump_if_ge(i2, a);
    bump_if_ge(i2, b);
    const int lo = a < i2 ? a : i2, hi = b > i2 ? b : i2, mid = a + b + i2 - lo - hi;
    bump_if_ge(i3, lo);
    bump_if_ge(i3, mid);
    bump_if_ge(i3, hi);
    idx[0] = i0; idx[1] = i1; idx[2] = i2; idx[3] = i3;
}
EGL_HD void seeded_subset(uint64_t seed, uint64_t frame, uint64_t K, uint64_t h, int N, int idx[4]) {
    (void)K;  // the first K hypotheses of a frame do not depend on K
    seeded_subset_keyed(seeded_frame_key(seed, frame), (uint32_t)h, N, idx);
}

// ---- fixed-K hypothesis arithmetic (FP32, every operation explicit so that oracle/ransac_f32.c
// can mirror it bit for bit; DESIGN.md "fixed-K arithmetic") --------------------------------------
//
// Per FRAME (once): the N correspondences are normalised
//     X' = (X - cX) * sX,  Y' = (Y - cY) * sY      cX = (sum X)/N, sX = N / sum|X - cX|   (image)
//     x' = (x - cx) * rt,  y' = (y - cy) * rt      cx = (sum x)/N, rt = 1/thr             (pitch)
// sums in WARP-TREE order (tree_sum64: t_i = v_i + v_{i+32}, then t_i += t_{i^16}, ^8, ^4, ^2, ^1; result t_0), which
// one warp evaluates with five shuffles.  Scaling the pitch side by 1/thr turns the inlier test
// |proj - dst|^2 <= thr^2 into  (nx - x'w)^2 + (ny - y'w)^2 <= w^2  with w the projective
// denominator: no division and no threshold multiply per point.
// Per HYPOTHESIS: the 8x8 DLT system of the 4 sampled points (h33 = 1)
//     X h0 + Y h1 + h2 - xX h6 - xY h7 = x ,   X h3 + Y h4 + h5 - yX h6 - yY h7 = y
// is solved by block elimination in registers: the left null vector n of the 4x3 matrix [X Y 1]
// (its 3x3 cofactors) removes h0..h5 and leaves a 2x2 system for (h6, h7); h0..h5 follow from the
// three best-conditioned rows by Cramer's rule.  Two reciprocals per hypothesis.
struct FixedKNorm {
    float cX, cY, sX, sY, cx, cy, rt;
};

// sum of v[0..n), n <= 64, in warp-tree order (scalar statement; the kernel does the same with shuffles)
EGL_HD float tree_sum64(const float* v, int n) {
    float t[32], u[32];
    for (int i = 0; i < 32; ++i) t[i] = fadd(i < n ? v[i] : 0.f, i + 32 < n ? v[i + 32] : 0.f);
    for (int m = 16; m > 0; m >>= 1) {
        for (int i = 0; i < 32; ++i) u[i] = fadd(t[i], t[i ^ m]);
        for (int i = 0; i < 32; ++i) t[i] = u[i];
    }
    return t[0];
}

EGL_HD bool fixedk_normalise(const float* sx, const float* sy, const float* dx, const float* dy, int n, float inv_thr,
                             FixedKNorm* nm) {
    const float fn = (float)n;
    nm->cX = fdiv(tree_sum64(sx, n), fn); nm->cY = fdiv(tree_sum64(sy, n), fn);
    nm->cx = fdiv(tree_sum64(dx, n), fn); nm->cy = fdiv(tree_sum64(dy, n), fn);
    float dX[64], dY[64];
    for (int i = 0; i < n; ++i) {
        dX[i] = fabsf(fsub(sx[i], nm->cX));
        dY[i] = fabsf(fsub(sy[i], nm->cY));
    }
    const float aX = tree_sum64(dX, n), aY = tree_sum64(dY, n);
    if (!(aX > 0.f) || !(aY > 0.f)) return false;
    nm->sX = fdiv(fn, aX);
    nm->sY = fdiv(fn, aY);
    nm->rt = inv_thr;
    return true;
}

EGL_HD void fixedk_normalise_point(const FixedKNorm& nm, float X, float Y, float x, float y, float* o) {
    o[0] = fmul(fsub(X, nm.cX), nm.sX);
    o[1] = fmul(fsub(Y, nm.cY), nm.sY);
    o[2] = fmul(fsub(x, nm.cx), nm.rt);
    o[3] = fmul(fsub(y, nm.cy), nm.rt);
}

// The hypothesis arithmetic is written once, over a lane type V: float (one hypothesis; host + device,
// the form tests/native/host_check.cpp and the C mirror are held against) or float2 (device only: TWO
// hypotheses per thread in Blackwell's packed FP32 instructions FFMA2 / FMUL2 / FADD2, IEEE per half,
// so each half is bit-identical to the float instantiation).  lane_ops<V> supplies the single-rounded
// operations; M is the per-lane predicate type.
template <class V> struct lane_ops;
template <> struct lane_ops<float> {
    typedef bool M;
    static EGL_HD float set(float v) { return v; }
    static EGL_HD float fma(float a, float b, float c) { return fmaf(a, b, c); }
    static EGL_HD float mul(float a, float b) { return fmul(a, b); }
    static EGL_HD float add(float a, float b) { return fadd(a, b); }
    static EGL_HD float sub(float a, float b) { return fsub(a, b); }
    static EGL_HD float neg(float a) { return -a; }
    static EGL_HD float abs(float a) { return fabsf(a); }
    static EGL_HD float rcp(float a) { return fdiv(1.f, a); }
    static EGL_HD bool gt(float a, float b) { return a > b; }
    static EGL_HD bool lt(float a, float b) { return a < b; }
    static EGL_HD float sel(bool m, float a, float b) { return m ? a : b; }
    static EGL_HD bool land(bool a, bool b) { return a && b; }
    static EGL_HD bool all_or_none(bool a, bool b, bool c, bool d) { const int n = a + b + c + d; return n == 0 || n == 4; }
};
#if defined(__CUDACC__)
struct bool2 { bool x, y; };
template <> struct lane_ops<float2> {
    typedef bool2 M;
    static __device__ __forceinline__ float2 set(float v) { return make_float2(v, v); }
    static __device__ __forceinline__ float2 fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
    static __device__ __forceinline__ float2 mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
    static __device__ __forceinline__ float2 add(float2 a, float2 b) { return __fadd2_rn(a, b); }
    static __device__ __forceinline__ float2 neg(float2 a) { return make_float2(-a.x, -a.y); }
    static __device__ __forceinline__ float2 sub(float2 a, float2 b) { return __fadd2_rn(a, neg(b)); }  // a - b == a + (-b) exactly
    static __device__ __forceinline__ float2 abs(float2 a) { return make_float2(fabsf(a.x), fabsf(a.y)); }
    static __device__ __forceinline__ float2 rcp(float2 a) { return make_float2(__fdiv_rn(1.f, a.x), __fdiv_rn(1.f, a.y)); }
    static __device__ __forceinline__ bool2 gt(float2 a, float2 b) { return bool2{a.x > b.x, a.y > b.y}; }
    static __device__ __forceinline__ bool2 lt(float2 a, float2 b) { return bool2{a.x < b.x, a.y < b.y}; }
    static __device__ __forceinline__ float2 sel(bool2 m, float2 a, float2 b) { return make_float2(m.x ? a.x : b.x, m.y ? a.y : b.y); }
    static __device__ __forceinline__ bool2 land(bool2 a, bool2 b) { return bool2{a.x && b.x, a.y && b.y}; }
    static __device__ __forceinline__ bool2 all_or_none(bool2 a, bool2 b, bool2 c, bool2 d) {
        const int nx = a.x + b.x + c.x + d.x, ny = a.y + b.y + c.y + d.y;
        return bool2{nx == 0 || nx == 4, ny == 0 || ny == 4};
    }
};
#endif

// det of rows (Xa,Ya,1),(Xb,Yb,1),(Xc,Yc,1):  Xa(Yb-Yc) - Ya(Xb-Xc) + (Xb Yc - Xc Yb)
template <class V>
EGL_HD V det3_v(V Xa, V Ya, V Xb, V Yb, V Xc, V Yc) {
    typedef lane_ops<V> L;
    const V t = L::fma(Xb, Yc, L::neg(L::mul(Xc, Yb)));
    return L::fma(Xa, L::sub(Yb, Yc), L::fma(L::neg(Ya), L::sub(Xb, Xc), t));
}
EGL_HD float det3_f32(float Xa, float Ya, float Xb, float Yb, float Xc, float Yc) { return det3_v<float>(Xa, Ya, Xb, Yb, Xc, Yc); }

// Sample test (OpenCV's checkSubset rules evaluated in float on the normalised points: last point collinear
// with an earlier pair on either side, or orientation not preserved).  S[4] = oriented areas of the image-side
// triples (0,1,2) (1,2,3) (0,2,3) (0,1,3), which the solve reuses as the left null vector of [X Y 1].
template <class V>
EGL_HD typename lane_ops<V>::M fixedk_check_v(const V* X, const V* Y, const V* x, const V* y, V* S) {
    typedef lane_ops<V> L;
    S[0] = det3_v<V>(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
    S[1] = det3_v<V>(X[1], Y[1], X[2], Y[2], X[3], Y[3]);
    S[2] = det3_v<V>(X[0], Y[0], X[2], Y[2], X[3], Y[3]);
    S[3] = det3_v<V>(X[0], Y[0], X[1], Y[1], X[3], Y[3]);
    const V D012 = det3_v<V>(x[0], y[0], x[1], y[1], x[2], y[2]);
    const V D123 = det3_v<V>(x[1], y[1], x[2], y[2], x[3], y[3]);
    const V D023 = det3_v<V>(x[0], y[0], x[2], y[2], x[3], y[3]);
    const V D013 = det3_v<V>(x[0], y[0], x[1], y[1], x[3], y[3]);
    const V tiny = L::set(1e-6f), zero = L::set(0.f);
    typename L::M ok = L::land(L::land(L::gt(L::abs(S[1]), tiny), L::gt(L::abs(S[2]), tiny)), L::gt(L::abs(S[3]), tiny));
    ok = L::land(ok, L::land(L::land(L::gt(L::abs(D123), tiny), L::gt(L::abs(D023), tiny)), L::gt(L::abs(D013), tiny)));
    return L::land(ok, L::all_or_none(L::lt(L::mul(S[0], D012), zero), L::lt(L::mul(S[1], D123), zero),
                                      L::lt(L::mul(S[2], D023), zero), L::lt(L::mul(S[3], D013), zero)));
}

// Minimal solve of an accepted sample: H[8] (h33 = 1) and the largest-|entry| guard; returns "all entries finite".
// X, Y, x, y are permuted in place by the pivoting.
template <class V>
EGL_HD typename lane_ops<V>::M fixedk_solve_v(V* X, V* Y, V* x, V* y, const V* S, V* H /*8*/, V* abs_sum = nullptr) {
    typedef lane_ops<V> L;
    // left null vector of [X Y 1]: n0 = S123, n1 = -S023, n2 = S013, n3 = -S012
    V n[4] = {S[1], L::neg(S[2]), S[3], L::neg(S[0])};
    // pivot: move the row with the largest |n| to slot 3 (the other three rows then have the largest 3x3 determinant)
EGL_UNROLL
    for (int k = 0; k < 3; ++k) {
        const typename L::M sw = L::gt(L::abs(n[k]), L::abs(n[3]));
        V t;
        t = n[3]; n[3] = L::sel(sw, n[k], t); n[k] = L::sel(sw, t, n[k]);
        t = X[3]; X[3] = L::sel(sw, X[k], t); X[k] = L::sel(sw, t, X[k]);
        t = Y[3]; Y[3] = L::sel(sw, Y[k], t); Y[k] = L::sel(sw, t, Y[k]);
        t = x[3]; x[3] = L::sel(sw, x[k], t); x[k] = L::sel(sw, t, x[k]);
        t = y[3]; y[3] = L::sel(sw, y[k], t); y[k] = L::sel(sw, t, y[k]);
    }
    // 2x2 system for (h6, h7):  sum n_i * row_i
    V a11 = L::set(0.f), a12 = a11, a21 = a11, a22 = a11, b1 = a11, b2 = a11;
EGL_UNROLL
    for (int k = 0; k < 4; ++k) {
        const V nx = L::mul(n[k], x[k]), ny = L::mul(n[k], y[k]);
        a11 = L::fma(L::neg(nx), X[k], a11); a12 = L::fma(L::neg(nx), Y[k], a12); b1 = L::add(b1, nx);
        a21 = L::fma(L::neg(ny), X[k], a21); a22 = L::fma(L::neg(ny), Y[k], a22); b2 = L::add(b2, ny);
    }
    const V Dt = L::fma(a11, a22, L::neg(L::mul(a12, a21)));
    const V rD = L::rcp(Dt);
    const V h6 = L::mul(L::fma(b1, a22, L::neg(L::mul(a12, b2))), rD);
    const V h7 = L::mul(L::fma(a11, b2, L::neg(L::mul(b1, a21))), rD);
    // h0..h5 from rows 0,1,2:  [X Y 1] (h0,h1,h2)^T = x*w,  (h3,h4,h5)^T likewise with y*w
    V u[3], v[3];
EGL_UNROLL
    for (int k = 0; k < 3; ++k) {
        const V w = L::fma(h6, X[k], L::fma(h7, Y[k], L::set(1.f)));
        u[k] = L::mul(x[k], w);
        v[k] = L::mul(y[k], w);
    }
    const V dP = det3_v<V>(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
    const V rP = L::rcp(dP);
    const V c0 = L::sub(Y[1], Y[2]), c1 = L::sub(Y[2], Y[0]), c2 = L::sub(Y[0], Y[1]);
    const V d0 = L::sub(X[2], X[1]), d1 = L::sub(X[0], X[2]), d2 = L::sub(X[1], X[0]);
    const V e0 = L::fma(X[1], Y[2], L::neg(L::mul(X[2], Y[1]))), e1 = L::fma(X[2], Y[0], L::neg(L::mul(X[0], Y[2]))),
            e2 = L::fma(X[0], Y[1], L::neg(L::mul(X[1], Y[0])));
    H[0] = L::mul(L::fma(u[0], c0, L::fma(u[1], c1, L::mul(u[2], c2))), rP);
    H[1] = L::mul(L::fma(u[0], d0, L::fma(u[1], d1, L::mul(u[2], d2))), rP);
    H[2] = L::mul(L::fma(u[0], e0, L::fma(u[1], e1, L::mul(u[2], e2))), rP);
    H[3] = L::mul(L::fma(v[0], c0, L::fma(v[1], c1, L::mul(v[2], c2))), rP);
    H[4] = L::mul(L::fma(v[0], d0, L::fma(v[1], d1, L::mul(v[2], d2))), rP);
    H[5] = L::mul(L::fma(v[0], e0, L::fma(v[1], e1, L::mul(v[2], e2))), rP);
    H[6] = h6;
    H[7] = h7;
    V acc = L::set(0.f);
EGL_UNROLL
    for (int k = 0; k < 8; ++k) acc = L::add(acc, L::abs(H[k]));
    if (abs_sum) *abs_sum = acc;          // sum of |entries|: an upper bound of every entry, for the caller's range checks
    return L::lt(acc, L::set(INFINITY));  // false for inf / NaN entries
}

// One hypothesis from 4 normalised correspondences p[k] = (X', Y', x', y').  Returns false when the
// sample is rejected or degenerate.
EGL_HD bool fixedk_hypothesis(const float (*p)[4], float* H /*8*/) {
    float X[4], Y[4], x[4], y[4], S[4];
EGL_UNROLL
    for (int k = 0; k < 4; ++k) { X[k] = p[k][0]; Y[k] = p[k][1]; x[k] = p[k][2]; y[k] = p[k][3]; }
    const bool ok = fixedk_check_v<float>(X, Y, x, y, S);
    const bool fin = fixedk_solve_v<float>(X, Y, x, y, S, H);
    return ok && fin;
}

// Inlier test of one normalised point under a normalised hypothesis (division-free).
EGL_HD bool fixedk_inlier(const float* H, float X, float Y, float x, float y) {
    const float w = fmaf(H[6], X, fmaf(H[7], Y, 1.f));
    const float ex = fmaf(-x, w, fmaf(H[0], X, fmaf(H[1], Y, H[2])));
    const float ey = fmaf(-y, w, fmaf(H[3], X, fmaf(H[4], Y, H[5])));
    const float e = fmaf(ex, ex, fmul(ey, ey));
    return fmaf(-w, w, e) <= 0.f;
}

// Back to image -> pitch units (double): H = Td^-1 * Hn * Ts, scaled so that H[8] == 1.
EGL_HD void fixedk_denormalise(const float* Hn, const FixedKNorm& nm, double thr, double* H) {
    const double h[9] = {Hn[0], Hn[1], Hn[2], Hn[3], Hn[4], Hn[5], Hn[6], Hn[7], 1.0};
    const double sX = nm.sX, sY = nm.sY, cX = nm.cX, cY = nm.cY, cx = nm.cx, cy = nm.cy;
    // A = Hn * Ts,  Ts = [[sX,0,-cX sX],[0,sY,-cY sY],[0,0,1]]
    double A[9];
    for (int r = 0; r < 3; ++r) {
        A[3 * r + 0] = h[3 * r + 0] * sX;
        A[3 * r + 1] = h[3 * r + 1] * sY;
        A[3 * r + 2] = h[3 * r + 2] - h[3 * r + 0] * cX * sX - h[3 * r + 1] * cY * sY;
    }
    // Td^-1 = [[thr,0,cx],[0,thr,cy],[0,0,1]]
    double G[9];
    for (int c = 0; c < 3; ++c) {
        G[c] = thr * A[c] + cx * A[6 + c];
        G[3 + c] = thr * A[3 + c] + cy * A[6 + c];
        G[6 + c] = A[6 + c];
    }
    const double s = 1.0 / G[8];
    for (int i = 0; i < 9; ++i) H[i] = G[i] * s;
}

// ---- line-intersection keypoint synthesis (coordinate_model.py:96-186) ------------------------
struct SynthTables {  // generated from eagle_b200/pitch.py (csrc/line_families.inc)
    const uint8_t* yfam_count; const uint8_t* yfam;  // [ny], [ny][maxm]
    const uint8_t* xfam_count; const uint8_t* xfam;  // [nx], [nx][maxm]
    const uint8_t* cross;                            // [ny][nx] channel at the crossing or 255
    int ny, nx, maxm;
};

// cv::fitLine(pts, DIST_L2, 0, 0.01, 0.01) on the detected members of one family (fitLine2D_wods:
// moments in double over float products, t = (float)atan2(2dxy, dx2-dy2)/2, direction (cos t, sin t)).
// Returns false if fewer than 2 members are detected or the direction degenerates (:111).
EGL_HD bool fit_family_line(const int32_t* xy, uint64_t detected, const uint8_t* members, int count, float* line) {
    double x = 0, y = 0, x2 = 0, y2 = 0, xy_ = 0;
    int n = 0;
    for (int m = 0; m < count; ++m) {
        const int ch = members[m];
        if (!((detected >> ch) & 1ull)) continue;
        const float px = (float)xy[2 * ch], py = (float)xy[2 * ch + 1];
        x += px; y += py;
        x2 += (double)fmul(px, px); y2 += (double)fmul(py, py); xy_ += (double)fmul(px, py);
        ++n;
    }
    if (n < 2) return false;
    const double w = (double)(float)n;
    x /= w; y /= w; x2 /= w; y2 /= w; xy_ /= w;
    const double dx2 = dsub(x2, dmul(x, x)), dy2 = dsub(y2, dmul(y, y)), dxy = dsub(xy_, dmul(x, y));
    const float t = (float)atan2(dmul(2.0, dxy), dsub(dx2, dy2)) / 2.f;
    // OpenCV calls the C library's cosf/sinf on the float angle; a correctly rounded double
    // evaluation differs from glibc's cosf by 1 ulp in ~1 % of cases (DESIGN.md, F1 tolerance).
    line[0] = (float)cos((double)t);
    line[1] = (float)sin((double)t);
    line[2] = (float)x;
    line[3] = (float)y;
    return !((double)fabsf(line[0]) + (double)fabsf(line[1]) < 1e-6);
}

// _intersect_lines (:117-138): 2x2 solve in double (LU with partial pivoting as LAPACK's gesv).
EGL_HD bool intersect_lines(const float* l1, const float* l2, double* px, double* py) {
    const double vx1 = l1[0], vy1 = l1[1], x01 = l1[2], y01 = l1[3];
    const double vx2 = l2[0], vy2 = l2[1], x02 = l2[2], y02 = l2[3];
    const double det = dsub(dmul(vx1, -vy2), dmul(vy1, -vx2));
    if (fabs(det) < 1e-8) return false;
    double a00 = vx1, a01 = -vx2, a10 = vy1, a11 = -vy2, b0 = x02 - x01, b1 = y02 - y01;
    if (fabs(a10) > fabs(a00)) {
        double t;
        t = a00; a00 = a10; a10 = t;
        t = a01; a01 = a11; a11 = t;
        t = b0; b0 = b1; b1 = t;
    }
    if (a00 == 0.0) return false;
    const double l = dmul(a10, 1.0 / a00);
    const double u11 = dsub(a11, dmul(l, a01));
    if (u11 == 0.0) return false;  // numpy raises LinAlgError -> the reference returns None
    const double s1 = dsub(b1, dmul(l, b0)) / u11;
    const double t0 = dsub(b0, dmul(a01, s1)) / a00;
    *px = dadd(x01, dmul(t0, vx1));
    *py = dadd(y01, dmul(t0, vy1));
    return true;
}

EGL_HD int32_t round_half_even_i32(double v) {  // int(round(v)) saturated to int32
    const double r = rint(v);
    if (!(r > -2147483648.0)) return (int32_t)0x80000000;
    if (!(r < 2147483647.0)) return 0x7fffffff;
    return (int32_t)r;
}

// _synthesize_keypoints_with_line_intersections (:140-186) for one frame.  xy/order/count are the
// frame's keypoint arrays; new landmarks are appended to order (and their xy written).  Returns the
// new total.  Called only when the frame has >= 2 keypoints (:326).
EGL_HD_NOINLINE int synthesize_keypoints(const SynthTables& T, int32_t* xy, uint8_t* order, int n, int max_new) {
    uint64_t present = 0;
    for (int j = 0; j < n; ++j) present |= 1ull << order[j];
    const uint64_t off_plane = (1ull << 0) | (1ull << 1) | (1ull << 24) | (1ull << 25);
    const uint64_t detected = present & ~off_plane;
    float ly[32][4], lx[32][4];
    uint32_t have_y = 0, have_x = 0;
    for (int i = 0; i < T.ny; ++i)
        if (fit_family_line(xy, detected, T.yfam + i * T.maxm, T.yfam_count[i], ly[i])) have_y |= 1u << i;
    for (int i = 0; i < T.nx; ++i)
        if (fit_family_line(xy, detected, T.xfam + i * T.maxm, T.xfam_count[i], lx[i])) have_x |= 1u << i;
    int added = 0;
    uint64_t added_mask = 0;
    for (int iy = 0; iy < T.ny && added < max_new; ++iy) {
        if (!((have_y >> iy) & 1u)) continue;
        for (int ix = 0; ix < T.nx && added < max_new; ++ix) {
            if (!((have_x >> ix) & 1u)) continue;
            const int ch = T.cross[iy * T.nx + ix];
            if (ch == 255 || ((present >> ch) & 1ull)) continue;
            double px, py;
            if (!intersect_lines(ly[iy], lx[ix], &px, &py)) continue;
            xy[2 * ch] = round_half_even_i32(px);
            xy[2 * ch + 1] = round_half_even_i32(py);
            if (!((added_mask >> ch) & 1ull)) {  // a dict: a label can be (re)written, its slot stays
                added_mask |= 1ull << ch;
                if (n + added < 64) order[n + added] = (uint8_t)ch;
                ++added;
            }
        }
    }
    return n + added;
}

}  // namespace egl
