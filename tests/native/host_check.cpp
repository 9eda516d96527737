// TEST ARTEFACT: compiles eagle_b200/csrc/geometry_core.cuh and flow_core.cuh -- the exact scalar code the CUDA
// kernels execute -- for the host with g++ -ffp-contract=off, so that its arithmetic can be compared
// with the oracle on a machine without a GPU (tests/test_host_core.py).  Never used by the product.
#include <stdint.h>
#include <string.h>
#define __constant__ static const
#include "../../include/eagle_b200.h"
#include "../../eagle_b200/csrc/geometry_core.cuh"
#include "../../eagle_b200/csrc/flow_core.cuh"
#include "../../eagle_b200/csrc/cascade_core.cuh"
#include <vector>

namespace egl {
#include "../../eagle_b200/csrc/line_families.inc"
}
using namespace egl;

extern "C" {

int hc_postprocess(const int32_t* flat, const float* score, int hm_h, int hm_w, int img_w, int img_h, double conf,
                   int32_t* xy, uint8_t* order) {
    return postprocess_keypoints(flat, score, hm_h, hm_w, img_w, img_h, conf, xy, order);
}

int hc_synthesize(int32_t* xy, uint8_t* order, int n, int max_new) {
    SynthTables T{kYFamCount, &kYFam[0][0], kXFamCount, &kXFam[0][0], &kCross[0][0], kNumYFam, kNumXFam, kMaxFam};
    return synthesize_keypoints(T, xy, order, n, max_new);
}

// sequential replay of ransac_cv2_kernel + refit_kernel for one point list
int hc_fit_cv2(const float* sx, const float* sy, const float* dx, const float* dy, int N, double thr, double confidence,
               int max_iters, double* H, uint64_t* mask, int32_t* info) {
    const float thr_sq = (float)(thr * thr);
    double scratch[192];
    info[0] = N; info[1] = 0; info[2] = -1; info[3] = 0;
    *mask = 0;
    if (N < 4) return EGL_FIT_FEW_POINTS;
    if (N == 4) {
        if (!run_kernel_ls(sx, sy, dx, dy, nullptr, 4, H, scratch)) return EGL_FIT_NO_MODEL;
        *mask = 0xF; info[1] = 4; info[2] = 0;
        return EGL_FIT_OK;
    }
    CvRng rng{~0ull};
    int niters = max_iters < 1 ? 1 : max_iters, iter = 0, best = 0;
    for (; iter < niters; ++iter) {
        int idx[4];
        bool found = false;
        float qx[4], qy[4], rx[4], ry[4];
        for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
            draw_subset(rng, N, idx);
            for (int k = 0; k < 4; ++k) { qx[k] = sx[idx[k]]; qy[k] = sy[idx[k]]; rx[k] = dx[idx[k]]; ry[k] = dy[idx[k]]; }
            found = check_subset(qx, qy, rx, ry);
        }
        if (!found) break;
        double Hm[9];
        if (!dlt4_f64(qx, qy, rx, ry, Hm)) continue;
        uint64_t m;
        const int c = inlier_mask_f32(Hm, sx, sy, dx, dy, N, thr_sq, &m);
        if (c > (best > 3 ? best : 3)) {
            best = c; *mask = m; info[2] = iter;
            memcpy(H, Hm, sizeof(Hm));
            niters = ransac_update_num_iters(confidence, (double)(N - c) / N, 4, niters);
        }
    }
    info[3] = iter;
    if (best == 0) return EGL_FIT_NO_MODEL;
    uint64_t fm;
    info[1] = refit_on_inliers(H, sx, sy, dx, dy, N, *mask, thr_sq, &fm, scratch);
    *mask = fm;
    return EGL_FIT_OK;
}

// sequential replay of ransac_fixedk_kernel (hypothesis stage only): winner denormalised to double,
// its position mask in OpenCV's float scoring, count and index
int hc_fixedk_stage(const float* sx, const float* sy, const float* dx, const float* dy, int N, int K, const uint8_t* hyp,
                    uint64_t seed, uint64_t frame, double thr, double* Hbest, uint64_t* mask, int32_t* info) {
    const float thr_sq = (float)(thr * thr);
    info[0] = N; info[1] = 0; info[2] = -1; info[3] = K;
    *mask = 0;
    FixedKNorm nm;
    if (N < 4) return EGL_FIT_FEW_POINTS;
    if (!fixedk_normalise(sx, sy, dx, dy, N, (float)(1.0 / thr), &nm)) return EGL_FIT_NO_MODEL;
    float pts[64][4];
    for (int i = 0; i < N; ++i) fixedk_normalise_point(nm, sx[i], sy[i], dx[i], dy[i], pts[i]);
    int best = 0, best_h = -1;
    float Hw[8];
    for (int h = 0; h < K; ++h) {
        int idx[4];
        if (hyp) { for (int k = 0; k < 4; ++k) idx[k] = hyp[4 * h + k]; }
        else seeded_subset(seed, frame, (uint64_t)K, (uint64_t)h, N, idx);
        bool ok = idx[0] < N && idx[1] < N && idx[2] < N && idx[3] < N;
        ok = ok && idx[0] != idx[1] && idx[0] != idx[2] && idx[0] != idx[3] && idx[1] != idx[2] && idx[1] != idx[3] && idx[2] != idx[3];
        if (!ok) continue;
        float p[4][4];
        for (int k = 0; k < 4; ++k) memcpy(p[k], pts[idx[k]], sizeof(p[k]));
        float Hn[8];
        if (!fixedk_hypothesis(p, Hn)) continue;
        int c = 0;
        for (int i = 0; i < N; ++i) c += fixedk_inlier(Hn, pts[i][0], pts[i][1], pts[i][2], pts[i][3]);
        if (c > best) { best = c; best_h = h; memcpy(Hw, Hn, sizeof(Hw)); }
    }
    info[2] = best_h;
    if (best <= 3) return EGL_FIT_NO_MODEL;
    fixedk_denormalise(Hw, nm, thr, Hbest);
    info[1] = inlier_mask_f32(Hbest, sx, sy, dx, dy, N, thr_sq, mask);
    return info[1] >= 4 ? EGL_FIT_OK : EGL_FIT_NO_MODEL;
}

int hc_refit(double* H, const float* sx, const float* sy, const float* dx, const float* dy, int N, uint64_t ransac_mask,
             double thr, uint64_t* final_mask) {
    double scratch[192];
    return refit_on_inliers(H, sx, sy, dx, dy, N, ransac_mask, (float)(thr * thr), final_mask, scratch);
}

// the RHO / LMEDS legs of the cascade (cascade_core.cuh) for one point list
int hc_fit_rho(const float* sx, const float* sy, const float* dx, const float* dy, int N, float* H, uint64_t* mask) {
    return rho_fit(sx, sy, dx, dy, N, H, mask);
}
int hc_fit_lmeds(const float* sx, const float* sy, const float* dx, const float* dy, int N, double confidence, double* H,
                 uint64_t* mask, uint64_t* band_mask) {
    double scratch[192];
    return lmeds_fit(sx, sy, dx, dy, N, confidence, H, mask, scratch, band_mask);
}
void hc_cv_jacobi9(double* A, double* W, double* V) { cv_jacobi9(A, W, V); }
int hc_run_kernel_ls(const float* sx, const float* sy, const float* dx, const float* dy, int N, double* H) {
    double scratch[192];
    uint8_t idx[64];
    for (int i = 0; i < N; ++i) idx[i] = (uint8_t)i;
    return run_kernel_ls(sx, sy, dx, dy, idx, N, H, scratch);
}
int hc_lm_refine(double* H, const float* sx, const float* sy, const float* dx, const float* dy, int N, int always_exact) {
    double scratch[192];
    uint8_t idx[64];
    for (int i = 0; i < N; ++i) idx[i] = (uint8_t)i;
    return lm_refine(H, sx, sy, dx, dy, idx, N, scratch, always_exact != 0);
}
int hc_refit_exact(double* H, const float* sx, const float* sy, const float* dx, const float* dy, int N, uint64_t ransac_mask,
                   double thr, uint64_t* final_mask) {
    double scratch[192];
    return refit_on_inliers(H, sx, sy, dx, dy, N, ransac_mask, (float)(thr * thr), final_mask, scratch, true);
}

int hc_dlt4_f64(const float* sx, const float* sy, const float* dx, const float* dy, double* H) { return dlt4_f64(sx, sy, dx, dy, H); }
int hc_check_subset(const float* sx, const float* sy, const float* dx, const float* dy) { return check_subset(sx, sy, dx, dy); }
void hc_seeded_subset(uint64_t seed, uint64_t frame, uint64_t K, uint64_t h, int N, int* idx) { seeded_subset(seed, frame, K, h, N, idx); }

void hc_project(const double* H, const float* pts, int n, float* out_f, int64_t* out_i) {
    for (int i = 0; i < n; ++i) {
        perspective_point(H, pts[2 * i], pts[2 * i + 1], &out_f[2 * i], &out_f[2 * i + 1]);
        out_i[2 * i] = trunc_like_numpy(out_f[2 * i]);
        out_i[2 * i + 1] = trunc_like_numpy(out_f[2 * i + 1]);
    }
}

// ---- keypoint propagation (flow_core.cuh) ----------------------------------------------------------
void hc_gray_hue(const uint8_t* bgr, int n, uint8_t* gray, uint8_t* hue) {
    for (int i = 0; i < n; ++i) {
        gray[i] = (uint8_t)gray_of(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
        hue[i] = (uint8_t)hue_of(bgr[3 * i], bgr[3 * i + 1], bgr[3 * i + 2]);
    }
}

long long hc_pyramid_bytes(int H, int W, int max_level) { return pyramid_layout(H, W, max_level).bytes; }

// gray image -> pyramid in the layout the kernels use (level 0 copied, further levels by pyrdown_pixel)
int hc_build_pyramid(const uint8_t* gray, int H, int W, int max_level, uint8_t* pyr) {
    const PyrLayout L = pyramid_layout(H, W, max_level);
    memcpy(pyr, gray, (size_t)H * W);
    for (int l = 1; l < L.n; ++l)
        for (int y = 0; y < L.h[l]; ++y)
            for (int x = 0; x < L.w[l]; ++x)
                pyr[L.off[l] + (long long)y * L.w[l] + x] = (uint8_t)pyrdown_pixel(pyr + L.off[l - 1], L.w[l - 1], L.h[l - 1], x, y);
    return L.n;
}

// cv2.calcOpticalFlowPyrLK(prev, next, pts, winSize=(15,15), maxLevel, criteria=(EPS|COUNT, max_count, eps)) through lk_track_point
void hc_track(const uint8_t* prev_gray, const uint8_t* next_gray, int H, int W, int max_level, const float* pts, int n, int max_count,
              double eps, float* out_pts, uint8_t* out_status) {
    const PyrLayout L = pyramid_layout(H, W, max_level);
    std::vector<uint8_t> a((size_t)L.bytes), b((size_t)L.bytes);
    hc_build_pyramid(prev_gray, H, W, max_level, a.data());
    hc_build_pyramid(next_gray, H, W, max_level, b.data());
    const int mc = max_count < 0 ? 0 : (max_count > 100 ? 100 : max_count);
    const double e = eps < 0 ? 0.0 : (eps > 10.0 ? 10.0 : eps);
    for (int i = 0; i < n; ++i)
        out_status[i] = (uint8_t)lk_track_point(a.data(), b.data(), L, pts[2 * i], pts[2 * i + 1], mc, e * e, 1e-4, out_pts + 2 * i);
}

float hc_pairwise_sum(const float* v, int n) { return pairwise_sum_f32(v, n); }

}  // extern "C"
