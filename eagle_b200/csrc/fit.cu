// K3: robust image->pitch homography per frame (replaces cv2.findHomography(RANSAC, 5.0) at
// eagle/models/coordinate_model.py:354-357 together with the correspondence gather at :335-349).
//
// Three kernels, enqueued back to back on one stream:
//   ransac_cv2_kernel   (mode EGL_FIT_CV2_COMPAT)  one warp per frame.  Walks OpenCV's RNG, so it
//                       draws the very 4-point samples cv2 would draw, solves 32 of them at a time
//                       (one per lane, double precision), scores them in float in OpenCV's
//                       operation order, then replays OpenCV's sequential accept / adaptive
//                       iteration-count rule over the 32 results.  Picks the model cv2 picks.
//   ransac_fixedk_kernel (mode EGL_FIT_FIXED_K)    one CTA per frame, one THREAD per hypothesis:
//                       points normalised once per frame, the 8x8 DLT system of each 4-point sample
//                       solved by block elimination entirely in FP32 registers, all N points scored
//                       by the same thread with a division-free test (13 FP32 ops per point);
//                       block-wide arg-max (ties -> lowest hypothesis index).  FP32-ALU bound.
//   refit_kernel        one thread per frame, FP64: OpenCV's tail of findHomography -- normalised
//                       DLT on the inliers (9x9 Jacobi), <= 10 Levenberg-Marquardt iterations over
//                       nine parameters, mask recomputed from the refined H.
#include "common.cuh"
#include "geometry_core.cuh"

namespace egl {

__constant__ float2 c_world[kLandmarks] = {
#include "pitch_table.inc"
};

struct FitArgs {
    const int32_t* kp_xy;
    const uint8_t* kp_order;
    const int32_t* kp_count;
    int F;
    int K;
    const uint8_t* hyp;
    uint64_t seed;
    float thr_sq;
    double confidence;
    double* H;
    uint64_t* used_mask;
    uint64_t* inlier_mask;
    int32_t* status;
    int32_t* info;
};

struct PointList {  // one frame's correspondences, float as cv2 receives them (:348-349)
    float sx[kMaxPts], sy[kMaxPts], dx[kMaxPts], dy[kMaxPts];
    uint8_t ch[kMaxPts];
};

// Warp-cooperative gather of the on-plane correspondences in kp_order order (:338-347).
__device__ __forceinline__ int gather_points_warp(const FitArgs& a, int f, PointList& pl, uint64_t* used) {
    const int lane = threadIdx.x & 31;
    const int n_total = min(a.kp_count[2 * f], EGL_ORDER_STRIDE);
    int base = 0;
    uint64_t um = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int j = pass * 32 + lane;
        const bool valid = j < n_total;
        const int ch = valid ? a.kp_order[(size_t)f * EGL_ORDER_STRIDE + j] : 0;
        const bool on = valid && ch < kLandmarks && !((kOffPlaneMask >> ch) & 1ull);
        const unsigned bal = __ballot_sync(kFull, on);
        const int pos = base + __popc(bal & ((1u << lane) - 1u));
        if (on) {
            pl.sx[pos] = (float)a.kp_xy[((size_t)f * kLandmarks + ch) * 2 + 0];
            pl.sy[pos] = (float)a.kp_xy[((size_t)f * kLandmarks + ch) * 2 + 1];
            pl.dx[pos] = c_world[ch].x;
            pl.dy[pos] = c_world[ch].y;
            pl.ch[pos] = (uint8_t)ch;
            um |= 1ull << ch;
        }
        base += __popc(bal);
    }
    // OR-reduce the channel mask
    unsigned lo = (unsigned)um, hi = (unsigned)(um >> 32);
    lo = __reduce_or_sync(kFull, lo);
    hi = __reduce_or_sync(kFull, hi);
    *used = ((uint64_t)hi << 32) | lo;
    __syncwarp();
    return base;
}

// Sequential gather for one thread (refit kernel).
__device__ int gather_points_thread(const FitArgs& a, int f, float* sx, float* sy, float* dx, float* dy, uint8_t* chs) {
    const int n_total = min(a.kp_count[2 * f], EGL_ORDER_STRIDE);
    int n = 0;
    for (int j = 0; j < n_total; ++j) {
        const int ch = a.kp_order[(size_t)f * EGL_ORDER_STRIDE + j];
        if (ch >= kLandmarks || ((kOffPlaneMask >> ch) & 1ull)) continue;
        sx[n] = (float)a.kp_xy[((size_t)f * kLandmarks + ch) * 2 + 0];
        sy[n] = (float)a.kp_xy[((size_t)f * kLandmarks + ch) * 2 + 1];
        dx[n] = c_world[ch].x;
        dy[n] = c_world[ch].y;
        chs[n] = (uint8_t)ch;
        ++n;
    }
    return n;
}

// Results of the hypothesis stage, parked in the output arrays until refit_kernel overwrites them:
//   H[f]           best model (double), inlier_mask[f] = its inliers as POSITION bits,
//   status[f]      EGL_FIT_OK / FEW_POINTS / NO_MODEL,  info[f] = {N, best count, best index, evaluated}
__device__ __forceinline__ void park(const FitArgs& a, int f, int status, int n, int count, int best, int iters,
                                     const double* H, uint64_t pos_mask, uint64_t used) {
    a.status[f] = status;
    a.info[4 * f + 0] = n;
    a.info[4 * f + 1] = count;
    a.info[4 * f + 2] = best;
    a.info[4 * f + 3] = iters;
    a.used_mask[f] = used;
    a.inlier_mask[f] = pos_mask;
    if (status == EGL_FIT_OK)
        for (int i = 0; i < 9; ++i) a.H[(size_t)f * 9 + i] = H[i];
}

// ------------------------------------------------------------------------------------------------
// mode EGL_FIT_CV2_COMPAT
// ------------------------------------------------------------------------------------------------
constexpr int kCv2Warps = 4;

__global__ void __launch_bounds__(kCv2Warps * 32) ransac_cv2_kernel(FitArgs a) {
    __shared__ PointList s_pl[kCv2Warps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kCv2Warps + warp;
    if (f >= a.F) return;
    PointList& pl = s_pl[warp];
    uint64_t used;
    const int N = gather_points_warp(a, f, pl, &used);
    double H[9];
    if (N < 4) {
        if (lane == 0) park(a, f, EGL_FIT_FEW_POINTS, N, 0, -1, 0, H, 0, used);
        return;
    }
    if (N == 4) {  // findHomography: npoints == 4 -> plain runKernel, mask of ones, no refinement
        if (lane == 0) {
            double scratch[192];
            const bool ok = run_kernel_ls(pl.sx, pl.sy, pl.dx, pl.dy, nullptr, 4, H, scratch);
            park(a, f, ok ? EGL_FIT_OK : EGL_FIT_NO_MODEL, 4, ok ? 4 : 0, 0, 0, H, ok ? 0xFull : 0ull, used);
        }
        return;
    }

    CvRng rng{~0ull};
    int niters = max(a.K, 1), iter = 0, best = 0, best_idx = -1;
    bool exhausted = false;
    while (iter < niters && !exhausted) {
        // every lane replays the same RNG walk; lane b keeps the b-th accepted sample
        int B = min(32, niters - iter);
        int my[4] = {0, 0, 0, 0};
        for (int b = 0; b < B; ++b) {
            int idx[4];
            bool found = false;
            for (int attempt = 0; attempt < 10000 && !found; ++attempt) {
                draw_subset(rng, N, idx);
                float qx[4], qy[4], rx[4], ry[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    qx[k] = pl.sx[idx[k]]; qy[k] = pl.sy[idx[k]];
                    rx[k] = pl.dx[idx[k]]; ry[k] = pl.dy[idx[k]];
                }
                found = check_subset(qx, qy, rx, ry);
            }
            if (!found) {  // getSubset failed: iter 0 -> no model at all, otherwise stop here
                B = b;
                exhausted = true;
                break;
            }
            if (lane == b) {
#pragma unroll
                for (int k = 0; k < 4; ++k) my[k] = idx[k];
            }
        }
        // one hypothesis per lane
        int cnt = 0;
        double Hm[9];
        if (lane < B) {
            float qx[4], qy[4], rx[4], ry[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                qx[k] = pl.sx[my[k]]; qy[k] = pl.sy[my[k]];
                rx[k] = pl.dx[my[k]]; ry[k] = pl.dy[my[k]];
            }
            if (dlt4_f64(qx, qy, rx, ry, Hm)) {
                float Hf[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) Hf[k] = (float)Hm[k];
                for (int i = 0; i < N; ++i) cnt += reproj_err_f32(Hf, pl.sx[i], pl.sy[i], pl.dx[i], pl.dy[i]) <= a.thr_sq;
            }
        }
        // OpenCV's sequential rule over the batch: accept iff strictly better (and >= 4 inliers),
        // then shrink niters; hypotheses at or beyond the new niters were never evaluated by cv2.
        int winner = -1, done = 0;
        for (int b = 0; b < B; ++b) {
            if (iter + b >= niters) break;
            const int c = __shfl_sync(kFull, cnt, b);
            ++done;
            if (c > max(best, 3)) {
                best = c;
                winner = b;
                best_idx = iter + b;
                niters = ransac_update_num_iters(a.confidence, (double)(N - c) / N, 4, niters);
            }
        }
        if (winner >= 0) {
#pragma unroll
            for (int k = 0; k < 9; ++k) H[k] = shfl_f64(Hm[k], winner);
        }
        iter += done;
        if (done < B) break;
    }
    if (lane == 0) {
        if (best == 0) {
            park(a, f, EGL_FIT_NO_MODEL, N, 0, -1, iter, H, 0, used);
        } else {
            uint64_t pm;
            inlier_mask_f32(H, pl.sx, pl.sy, pl.dx, pl.dy, N, a.thr_sq, &pm);
            park(a, f, EGL_FIT_OK, N, best, best_idx, iter, H, pm, used);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// mode EGL_FIT_FIXED_K: thread per hypothesis, FP32 registers
// ------------------------------------------------------------------------------------------------
constexpr int kFixedThreads = 128;

__global__ void __launch_bounds__(kFixedThreads) ransac_fixedk_kernel(FitArgs a, float inv_thr, double thr) {
    __shared__ PointList s_pl;
    __shared__ float4 s_pt[kMaxPts];  // normalised (X', Y', x', y'): one broadcast LDS.128 per point
    __shared__ int s_cnt[kFixedThreads / 32], s_hyp[kFixedThreads / 32];
    __shared__ float s_H[8];
    __shared__ FixedKNorm s_nm;
    __shared__ int s_N, s_ok;
    __shared__ unsigned long long s_used;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        uint64_t used;
        const int n = gather_points_warp(a, f, s_pl, &used);
        if (lane == 0) {
            s_N = n;
            s_used = used;
            s_ok = n >= 4 && fixedk_normalise(s_pl.sx, s_pl.sy, s_pl.dx, s_pl.dy, n, inv_thr, &s_nm);
        }
    }
    __syncthreads();
    const int N = s_N;
    if (N < 4 || !s_ok) {
        if (tid == 0) park(a, f, N < 4 ? EGL_FIT_FEW_POINTS : EGL_FIT_NO_MODEL, N, 0, -1, 0, nullptr, 0, s_used);
        return;
    }
    for (int i = tid; i < N; i += kFixedThreads) {
        float o[4];
        fixedk_normalise_point(s_nm, s_pl.sx[i], s_pl.sy[i], s_pl.dx[i], s_pl.dy[i], o);
        s_pt[i] = make_float4(o[0], o[1], o[2], o[3]);
    }
    __syncthreads();

    int best_cnt = 0, best_h = 0x7fffffff;
    float best_H[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int h = tid; h < a.K; h += kFixedThreads) {
        int idx[4];
        if (a.hyp) {
            const uchar4 q = reinterpret_cast<const uchar4*>(a.hyp)[(size_t)f * a.K + h];
            idx[0] = q.x; idx[1] = q.y; idx[2] = q.z; idx[3] = q.w;
        } else {
            seeded_subset(a.seed, (uint64_t)f, (uint64_t)a.K, (uint64_t)h, N, idx);
        }
        bool ok = idx[0] < N && idx[1] < N && idx[2] < N && idx[3] < N;
        ok = ok && idx[0] != idx[1] && idx[0] != idx[2] && idx[0] != idx[3] && idx[1] != idx[2] && idx[1] != idx[3] &&
             idx[2] != idx[3];
        float p[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 q = s_pt[ok ? idx[k] : k];
            p[k][0] = q.x; p[k][1] = q.y; p[k][2] = q.z; p[k][3] = q.w;
        }
        float Hn[8];
        ok = fixedk_hypothesis(p, Hn) && ok;
        int cnt = 0;
        if (ok) {
#pragma unroll 4
            for (int i = 0; i < N; ++i) {
                const float4 q = s_pt[i];
                cnt += fixedk_inlier(Hn, q.x, q.y, q.z, q.w);
            }
        }
        if (cnt > best_cnt) {  // ascending h per thread: strict > keeps the earliest
            best_cnt = cnt;
            best_h = h;
#pragma unroll
            for (int k = 0; k < 8; ++k) best_H[k] = Hn[k];
        }
    }
    // block arg-max: most inliers, ties -> lowest hypothesis index
    int c = best_cnt, hh = best_h;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const int oc = __shfl_xor_sync(kFull, c, m), oh = __shfl_xor_sync(kFull, hh, m);
        if (oc > c || (oc == c && oh < hh)) { c = oc; hh = oh; }
    }
    if (lane == 0) { s_cnt[warp] = c; s_hyp[warp] = hh; }
    __syncthreads();
    c = s_cnt[0]; hh = s_hyp[0];
#pragma unroll
    for (int w2 = 1; w2 < kFixedThreads / 32; ++w2)
        if (s_cnt[w2] > c || (s_cnt[w2] == c && s_hyp[w2] < hh)) { c = s_cnt[w2]; hh = s_hyp[w2]; }
    const bool have = c > 3;  // cv2: goodCount > max(maxGoodCount, modelPoints - 1)
    if (have && best_h == hh && best_cnt == c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) s_H[k] = best_H[k];
    }
    __syncthreads();
    if (warp == 0) {
        // winner back in image -> pitch units; its inlier list in OpenCV's float scoring (lanes are
        // points, ballots make the mask) is what the refit starts from
        double Hd[9];
        uint64_t pm = 0;
        if (have) {
            float Hn[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) Hn[k] = s_H[k];
            fixedk_denormalise(Hn, s_nm, thr, Hd);
            float Hf[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) Hf[k] = (float)Hd[k];
            for (int pass = 0; pass < 2; ++pass) {
                const int i = pass * 32 + lane;
                const bool in = i < N && reproj_err_f32(Hf, s_pl.sx[i], s_pl.sy[i], s_pl.dx[i], s_pl.dy[i]) <= a.thr_sq;
                pm |= (uint64_t)__ballot_sync(kFull, in) << (32 * pass);
            }
        }
        if (lane == 0) {
            const bool good = have && __popcll(pm) >= 4;
            park(a, f, good ? EGL_FIT_OK : EGL_FIT_NO_MODEL, N, good ? __popcll(pm) : 0, have ? hh : -1, a.K, Hd, pm, s_used);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// refit: OpenCV's tail of findHomography, one thread per frame, FP64
// ------------------------------------------------------------------------------------------------
constexpr int kRefitThreads = 64;

__global__ void __launch_bounds__(kRefitThreads) refit_kernel(FitArgs a) {
    const int f = blockIdx.x * kRefitThreads + threadIdx.x;
    if (f >= a.F) return;
    if (a.status[f] != EGL_FIT_OK) {
        a.inlier_mask[f] = 0;
        return;
    }
    float sx[kMaxPts], sy[kMaxPts], dx[kMaxPts], dy[kMaxPts];
    uint8_t ch[kMaxPts];
    const int N = gather_points_thread(a, f, sx, sy, dx, dy, ch);
    double H[9];
    for (int i = 0; i < 9; ++i) H[i] = a.H[(size_t)f * 9 + i];
    uint64_t pm = a.inlier_mask[f];
    int count = a.info[4 * f + 1];
    if (N > 4) {
        double scratch[192];
        uint64_t fm;
        count = refit_on_inliers(H, sx, sy, dx, dy, N, pm, a.thr_sq, &fm, scratch);
        pm = fm;
        for (int i = 0; i < 9; ++i) a.H[(size_t)f * 9 + i] = H[i];
    }
    uint64_t cm = 0;
    for (int i = 0; i < N; ++i)
        if ((pm >> i) & 1ull) cm |= 1ull << ch[i];
    a.inlier_mask[f] = cm;
    a.info[4 * f + 1] = count;
}

}  // namespace egl

using namespace egl;

extern "C" int egl_fit_homography(const int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count, int F, int mode,
                                  int K, const uint8_t* hyp, uint64_t seed, double thr, double confidence, double* H,
                                  uint64_t* used_mask, uint64_t* inlier_mask, int32_t* status, int32_t* info,
                                  void* stream) {
    EGL_REQUIRE(kp_xy && kp_order && kp_count && H && used_mask && inlier_mask && status && info, EGL_ERR_NULL,
                "egl_fit_homography: null pointer");
    EGL_REQUIRE(F >= 0 && K >= 1, EGL_ERR_SHAPE, "egl_fit_homography: need F >= 0 and K >= 1 (F=%d K=%d)", F, K);
    EGL_REQUIRE(mode == EGL_FIT_CV2_COMPAT || mode == EGL_FIT_FIXED_K, EGL_ERR_MODE, "egl_fit_homography: unknown mode %d", mode);
    EGL_REQUIRE(confidence > 0 && confidence < 1, EGL_ERR_SHAPE, "egl_fit_homography: confidence must be in (0,1)");
    if (F == 0) return 0;
    if (!(thr > 0)) thr = 3.0;  // findHomography: ransacReprojThreshold <= 0 -> 3
    FitArgs a{kp_xy, kp_order, kp_count, F, K, hyp, seed, (float)(thr * thr), confidence, H, used_mask, inlier_mask, status, info};
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == EGL_FIT_CV2_COMPAT) {
        ransac_cv2_kernel<<<(F + kCv2Warps - 1) / kCv2Warps, kCv2Warps * 32, 0, s>>>(a);
    } else {
        ransac_fixedk_kernel<<<F, kFixedThreads, 0, s>>>(a, (float)(1.0 / thr), thr);
    }
    int rc = cuda_status(cudaGetLastError(), "egl_fit_homography: hypothesis kernel launch");
    if (rc) return rc;
    refit_kernel<<<(F + kRefitThreads - 1) / kRefitThreads, kRefitThreads, 0, s>>>(a);
    return cuda_status(cudaGetLastError(), "egl_fit_homography: refit kernel launch");
}
