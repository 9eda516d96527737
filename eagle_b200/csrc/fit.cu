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
//   cascade_kernel      (mode EGL_FIT_CV2_COMPAT) the rest of the reference's cascade, `for method in [cv2.RANSAC,
//                       cv2.RHO, cv2.LMEDS]` (:354-357), for the frames the RANSAC leg left without a model: one warp
//                       per such frame runs OpenCV's RHO estimator (lane 0: the algorithm is a chain) and, if that
//                       fails too, LMedS with the samples rated on the 32 lanes (cascade_core.cuh); every other warp
//                       exits on its status word.
//   refit_warp_kernel   one warp per frame, FP64: OpenCV's tail of findHomography -- normalised DLT on
//                       the inliers (smallest eigenvector of the 9x9 normal matrix), <= 10
//                       Levenberg-Marquardt iterations over nine parameters, mask recomputed from the
//                       refined H.  (refit_kernel = the same algorithm, one thread per frame, running
//                       the host-checkable scalar code of geometry_core.cuh; EGL_REFIT_VARIANT=1.)
#include <stdlib.h>

#include <stddef.h>

#include "common.cuh"
#include "geometry_core.cuh"
#include "cascade_core.cuh"

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
    const uint8_t* sched;  // optional: fit frame f only where sched[f] | retry[f] (cadence of coordinate_model.py:333)
    const uint8_t* retry;
    const float* kp_sub;   // optional: sub-pixel image positions [F][57][2] used instead of the integer kp_xy
};

__device__ __forceinline__ float image_coord(const FitArgs& a, int f, int ch, int axis) {
    const size_t i = ((size_t)f * kLandmarks + ch) * 2 + axis;
    return a.kp_sub ? a.kp_sub[i] : (float)a.kp_xy[i];
}

// Frames the cadence does not fit this step: status EGL_FIT_SKIPPED, nothing else touched but the masks.
__device__ __forceinline__ bool fit_skipped(const FitArgs& a, int f) {
    return a.sched && !(a.sched[f] | (a.retry ? a.retry[f] : (uint8_t)0));
}
__device__ __forceinline__ void park_skipped(const FitArgs& a, int f) {
    a.status[f] = EGL_FIT_SKIPPED;
    a.used_mask[f] = 0;
    a.inlier_mask[f] = 0;
    a.info[4 * f + 0] = a.info[4 * f + 1] = a.info[4 * f + 3] = 0;
    a.info[4 * f + 2] = -1;
}

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
            pl.sx[pos] = image_coord(a, f, ch, 0);
            pl.sy[pos] = image_coord(a, f, ch, 1);
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
        sx[n] = image_coord(a, f, ch, 0);
        sy[n] = image_coord(a, f, ch, 1);
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
    if (fit_skipped(a, f)) {
        if (lane == 0) park_skipped(a, f);
        return;
    }
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
    int rejected = 0;  // consecutive rejected samples: OpenCV's getSubset gives up after 10000 per iteration
    bool exhausted = false;
    while (iter < niters && !exhausted) {
        // ---- fill up to 32 hypothesis slots (lane s = slot s) with the next accepted samples --------
        // The RNG stream fixes the sequence of candidate samples; whether a candidate passes checkSubset
        // does not feed back into the stream.  So every lane replays the (cheap) RNG walk for a group of
        // 32 candidates, lane j keeps candidate j and tests it -- the expensive part, now in parallel --
        // and a ballot compacts the accepted ones, in stream order, into the slots.  The stream is then
        // rewound to just after the last candidate actually consumed.
        const int B = min(32, niters - iter);
        int filled = 0;
        int my[4] = {0, 0, 0, 0};
        while (filled < B && !exhausted) {
            int cand[4] = {0, 0, 0, 0};
            uint64_t st_after = 0;
            for (int j = 0; j < 32; ++j) {
                int idx[4];
                draw_subset(rng, N, idx);
                if (lane == j) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) cand[k] = idx[k];
                    st_after = rng.state;
                }
            }
            float qx[4], qy[4], rx[4], ry[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                qx[k] = pl.sx[cand[k]]; qy[k] = pl.sy[cand[k]];
                rx[k] = pl.dx[cand[k]]; ry[k] = pl.dy[cand[k]];
            }
            unsigned bal = __ballot_sync(kFull, check_subset(qx, qy, rx, ry));
            const int legit = min(32, 10000 - rejected);       // attempts OpenCV would still make
            if (legit < 32) bal &= (1u << legit) - 1u;
            const int acc = __popc(bal);
            const int take = min(acc, B - filled);
            // slot (filled + r) <- the r-th accepted candidate of this group
            const int r = lane - filled;
            const int src = (r >= 0 && r < take) ? (int)__fns(bal, 0, r + 1) : lane;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int v = __shfl_sync(kFull, cand[k], src);
                if (r >= 0 && r < take) my[k] = v;
            }
            filled += take;
            int last = 31;                                      // last candidate examined
            if (filled == B && take > 0) last = (int)__fns(bal, 0, take);
            else if (acc == 0 && legit < 32) last = legit - 1;
            rng.state = (uint64_t)__shfl_sync(kFull, (unsigned)(st_after >> 32), last) << 32 |
                        (uint64_t)__shfl_sync(kFull, (unsigned)st_after, last);
            if (acc == 0) {
                rejected += legit;
                if (rejected >= 10000) exhausted = true;        // getSubset failed for iteration iter + filled
            } else {
                // rejected samples after the last accepted one examined so far
                rejected = last - (31 - __clz(bal & (last == 31 ? 0xffffffffu : ((2u << last) - 1u))));
            }
        }
        const int Bf = filled;  // hypotheses actually available (== B unless the sampler gave up)
        // ---- one hypothesis per lane -------------------------------------------------------------------
        int cnt = 0;
        double Hm[9];
        if (lane < Bf) {
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
        for (int b = 0; b < Bf; ++b) {
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
        if (done < Bf) break;
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
// mode EGL_FIT_FIXED_K: packed FP32 (FFMA2 / FMUL2 / FADD2) throughout
// ------------------------------------------------------------------------------------------------
constexpr int kFixedThreads = 128;
constexpr int kFixedWarps = kFixedThreads / 32;
constexpr int kFixedQueue = 96;  // per-warp queue of accepted samples: <= 63 left over + 32 new

// sample h of frame f: indices from the explicit table or the counter-based generator; false if unusable
__device__ __forceinline__ bool fixedk_sample(const FitArgs& a, int f, uint64_t key, int h, int N, int idx[4]) {
    if (!a.hyp) {
        seeded_subset_keyed(key, (uint32_t)h, N, idx);  // always four distinct indices below N
        return h < a.K;
    }
    bool ok = h < a.K;
    const uchar4 q = ok ? reinterpret_cast<const uchar4*>(a.hyp)[(size_t)f * a.K + h] : make_uchar4(0, 1, 2, 3);
    idx[0] = q.x; idx[1] = q.y; idx[2] = q.z; idx[3] = q.w;
    ok = ok && idx[0] < N && idx[1] < N && idx[2] < N && idx[3] < N && idx[0] != idx[1] && idx[0] != idx[2] && idx[0] != idx[3] &&
         idx[1] != idx[2] && idx[1] != idx[3] && idx[2] != idx[3];
    if (!ok) { idx[0] = 0; idx[1] = 1; idx[2] = 2; idx[3] = 3; }
    return ok;
}

// shared-memory read of data that is not written again after the barrier that published it
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// asm keeps the count at one compare and one predicated add per hypothesis and point
__device__ __forceinline__ void count_if_le0(int& c, float t) {
    asm("{\n.reg .pred p;\nsetp.le.f32 p, %1, 0f00000000;\n@p add.s32 %0, %0, 1;\n}" : "+r"(c) : "f"(t));
}

__device__ __forceinline__ float warp_tree_sum_f32(float lo, float hi) {  // tree_sum64 with lanes as the 32 partial sums
    float t = fadd(lo, hi);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) t = fadd(t, __shfl_xor_sync(kFull, t, m));
    return t;
}

// Per-frame input of the hypothesis kernel, written by fixedk_prepare_kernel: the gathered correspondences, their
// normalised form (X', x', Y', y') and the normalisation.  2.1 KB per frame in a stream-ordered scratch allocation.
struct FixedKPrep {
    float4 ps[kMaxPts];
    float sx[kMaxPts], sy[kMaxPts], dx[kMaxPts], dy[kMaxPts];
    uint8_t ch[kMaxPts];
    FixedKNorm nm;
    int N, ok;
    unsigned cmax_bits;   // bit pattern of the largest |normalised coordinate| of the frame
    unsigned long long used;
};

// Gather + normalise, one warp per frame.  This is a chain of dependent global loads (count -> order -> positions) and
// shuffle reductions: inside the hypothesis kernel it kept three of a CTA's four warps waiting for ~10 % of the CTA's
// life; here thousands of frames are in flight at once and the latency disappears behind them.
constexpr int kPrepWarps = 8;
__global__ void __launch_bounds__(kPrepWarps * 32) fixedk_prepare_kernel(FitArgs a, float inv_thr, FixedKPrep* prep) {
    __shared__ PointList s_pl[kPrepWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kPrepWarps + warp;
    if (f >= a.F || fit_skipped(a, f)) return;
    PointList& pl = s_pl[warp];
    FixedKPrep& o = prep[f];
    uint64_t used;
    const int n = gather_points_warp(a, f, pl, &used);
    // fixedk_normalise, lanes as the partial sums of the tree order
    const bool v0 = lane < n, v1 = lane + 32 < n;
    const float X0 = v0 ? pl.sx[lane] : 0.f, X1 = v1 ? pl.sx[lane + 32] : 0.f;
    const float Y0 = v0 ? pl.sy[lane] : 0.f, Y1 = v1 ? pl.sy[lane + 32] : 0.f;
    const float x0 = v0 ? pl.dx[lane] : 0.f, x1 = v1 ? pl.dx[lane + 32] : 0.f;
    const float y0 = v0 ? pl.dy[lane] : 0.f, y1 = v1 ? pl.dy[lane + 32] : 0.f;
    const float fn = (float)n;
    FixedKNorm nm;
    nm.cX = fdiv(warp_tree_sum_f32(X0, X1), fn); nm.cY = fdiv(warp_tree_sum_f32(Y0, Y1), fn);
    nm.cx = fdiv(warp_tree_sum_f32(x0, x1), fn); nm.cy = fdiv(warp_tree_sum_f32(y0, y1), fn);
    const float aX = warp_tree_sum_f32(v0 ? fabsf(fsub(X0, nm.cX)) : 0.f, v1 ? fabsf(fsub(X1, nm.cX)) : 0.f);
    const float aY = warp_tree_sum_f32(v0 ? fabsf(fsub(Y0, nm.cY)) : 0.f, v1 ? fabsf(fsub(Y1, nm.cY)) : 0.f);
    nm.sX = fdiv(fn, aX); nm.sY = fdiv(fn, aY); nm.rt = inv_thr;
    if (lane == 0) {
        o.N = n;
        o.used = used;
        o.ok = n >= 4 && aX > 0.f && aY > 0.f;
        o.nm = nm;
    }
    unsigned cmax = 0;
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
        const int i = pass * 32 + lane;
        if (i >= n) continue;
        float q[4];
        fixedk_normalise_point(nm, pl.sx[i], pl.sy[i], pl.dx[i], pl.dy[i], q);
        o.ps[i] = make_float4(q[0], q[2], q[1], q[3]);
        o.sx[i] = pl.sx[i]; o.sy[i] = pl.sy[i]; o.dx[i] = pl.dx[i]; o.dy[i] = pl.dy[i];
        o.ch[i] = pl.ch[i];
#pragma unroll
        for (int k = 0; k < 4; ++k) cmax = max(cmax, __float_as_uint(q[k]) & 0x7fffffffu);
    }
    cmax = __reduce_max_sync(kFull, cmax);
    if (lane == 0) o.cmax_bits = cmax;
}

// Template parameters (the shipped build instantiates one combination; -DEGL_BENCH_VARIANTS builds hold the others for
// tools/fixedk_variants.py): CTAs per SM asked of the register allocator, unroll of the scoring loop, and whether the warps
// take their batches of 32 hypotheses from a shared counter instead of a fixed stride (the result does not depend on which
// warp scores which hypothesis: most inliers, ties to the lowest hypothesis index).
// kSignCount: where no NaN can arise (every |H entry| < 1e15 and every normalised coordinate < 1e3, checked per batch) the
// inlier test  t = e - w^2 <= 0  is evaluated as the SIGN BIT of  -t = fma(w, w, -e)  (the exact negation: round-to-nearest
// is symmetric; an exact zero is +0 on both sides) and accumulated with one shift-add (LEA.HI) per hypothesis and point
// instead of a compare and a predicated add; padding rows (x' = +inf) and overflowed residuals give -inf = outlier on
// both paths.  Batches that fail the range check take the compare-and-add loop, so the counts are the specification's
// in every case (oracle/ransac_f32.c).  MEASURED AND NOT SHIPPED: 8 of the 73 instructions of a four-point loop body go
// away, issue activity falls from 66.8 to 64.6 % and the kernel takes exactly as long (1.276 against 1.275 ms for 20 k
// frames, outputs bit-identical; profiles/r2_fixedk_variants.txt) -- the loop waits on the FMA pipe
// (stall_math_pipe_throttle 2.1 per issue), not on issue slots.  Kept as a variant of the measurement build.
template <int kMinBlocks, int kUnroll, bool kDynamic, bool kSignCount = false>
__global__ void __launch_bounds__(kFixedThreads, kMinBlocks) ransac_fixedk_kernel(FitArgs a, float inv_thr, double thr, const FixedKPrep* prep) {
    static_assert(kUnroll == 2 || kUnroll == 4 || kUnroll == 8, "scoring loop unroll");
    __shared__ PointList s_pl;
    // normalised points as (X', x', Y', y') -- image/pitch pairs for the sample test -- replicated into the eight
    // 16-byte columns of a 128-byte row: lane l gathers from column l & 7, so the eight lanes of a quarter warp never
    // share a bank whatever indices they drew (a random LDS.128 gather from an unreplicated table averages 3
    // wavefronts per quarter warp)
    __shared__ float4 s_ps[kMaxPts][8];
    __shared__ float4 s_bc[kMaxPts + 8][2];  // for the scoring loop: (X',X',Y',Y') (-x',-x',-y',-y'), padded to a multiple of the unroll
    __shared__ int s_cnt[kFixedWarps], s_hyp[kFixedWarps];
    __shared__ uint32_t s_qi[kFixedWarps][kFixedQueue];  // accepted samples: four packed indices ...
    __shared__ int s_qh[kFixedWarps][kFixedQueue];       // ... and the hypothesis number
    __shared__ uint32_t s_pi[kFixedWarps * 64];          // what the warps had left over, pooled
    __shared__ int s_ph[kFixedWarps * 64];
    __shared__ FixedKNorm s_nm;
    __shared__ int s_N, s_pool, s_next;
    __shared__ unsigned long long s_used;
    const int f = blockIdx.x;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (fit_skipped(a, f)) {
        if (tid == 0) park_skipped(a, f);
        return;
    }
    const FixedKPrep& in = prep[f];
    const int N = in.N;
    const bool coords_small = in.cmax_bits < 0x447a0000u;  // every normalised coordinate of the frame below 1000 (NaN / inf order above)
    if (N < 4 || !in.ok) {
        if (tid == 0) park(a, f, N < 4 ? EGL_FIT_FEW_POINTS : EGL_FIT_NO_MODEL, N, 0, -1, 0, nullptr, 0, in.used);
        return;
    }
    if (tid == 0) {
        s_N = N;
        s_used = in.used;
        s_nm = in.nm;
        s_pool = 0;
        s_next = 0;
    }
    const int N4 = (N + kUnroll - 1) & ~(kUnroll - 1);
    for (int i = tid; i < N4; i += kFixedThreads) {
        if (i < N) {
            const float4 q = in.ps[i];  // (X', x', Y', y')
#pragma unroll
            for (int col = 0; col < 8; ++col) s_ps[i][col] = q;
            s_bc[i][0] = make_float4(q.x, q.x, q.z, q.z);
            s_bc[i][1] = make_float4(-q.y, -q.y, -q.w, -q.w);
            s_pl.sx[i] = in.sx[i]; s_pl.sy[i] = in.sy[i]; s_pl.dx[i] = in.dx[i]; s_pl.dy[i] = in.dy[i];
        } else {  // padding rows: w = 1 and an infinite residual, never an inlier (and never a NaN)
            const float q = __int_as_float(0x7f800000);
            s_bc[i][0] = make_float4(0.f, 0.f, 0.f, 0.f);
            s_bc[i][1] = make_float4(q, q, q, q);
        }
    }
    __syncthreads();
    const uint64_t key = seeded_frame_key(a.seed, (uint64_t)f);
    // 32-bit shared-window addresses, formed once: this lane's column of the gather table (row stride 128 bytes) and
    // the scoring table
    const uint32_t ps_addr = smem_u32(&s_ps[0][lane & 7]), bc_addr = smem_u32(&s_bc[0][0]);

    // Two phases so that no lane idles while others score:
    //   A  every lane draws one sample and tests it (checkSubset rules; 50-70 % are rejected at 40 % outliers).  The
    //      image-side and pitch-side triple areas are the same formula on (X,Y) and (x,y), so the packed halves are
    //      the two SIDES: one LDS.128 per sampled point lands as the pairs (X,x) (Y,y), no register shuffling.
    //      Accepted samples are ballot-compacted into a per-warp queue (8 bytes per entry);
    //   B  whenever 64 samples are queued, every lane takes two, solves both minimal systems and scores all N points
    //      against both -- the packed halves are now the two HYPOTHESES: 11 FFMA2/FMUL2 per point for the pair.
    // What the warps have left at the end is pooled so that at most one batch per frame runs partly empty.
    // The result is order independent: most inliers, ties to the lowest hypothesis index.
    uint32_t* qi = s_qi[warp];
    int* qh = s_qh[warp];
    int qn = 0;
    int best_cnt = 0, best_h = 0x7fffffff;
    typedef lane_ops<float2> L;
    auto solve_and_score = [&](const uint32_t* ei, const int* eh, int e0, int e1, bool v0, bool v1) {
        const uint32_t pa = ei[e0], pb = ei[e1];
        const int ha = eh[e0], hb = eh[e1];
        float2 X[4], Y[4], x[4], y[4], S[4], H[8];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 p = lds128(ps_addr + 128u * ((pa >> (8 * k)) & 255u)), q = lds128(ps_addr + 128u * ((pb >> (8 * k)) & 255u));
            X[k] = make_float2(p.x, q.x); x[k] = make_float2(p.y, q.y);
            Y[k] = make_float2(p.z, q.z); y[k] = make_float2(p.w, q.w);
        }
        S[0] = det3_v<float2>(X[0], Y[0], X[1], Y[1], X[2], Y[2]);
        S[1] = det3_v<float2>(X[1], Y[1], X[2], Y[2], X[3], Y[3]);
        S[2] = det3_v<float2>(X[0], Y[0], X[2], Y[2], X[3], Y[3]);
        S[3] = det3_v<float2>(X[0], Y[0], X[1], Y[1], X[3], Y[3]);
        float2 hs;
        const bool2 fin = fixedk_solve_v<float2>(X, Y, x, y, S, H, &hs);
        const float2 one = make_float2(1.f, 1.f);
        int c0 = 0, c1 = 0;
        const bool in_range = hs.x < 1e15f && hs.y < 1e15f;
        if (kSignCount && coords_small && __all_sync(kFull, in_range)) {
            unsigned o0 = 0, o1 = 0;   // outliers (padding rows included)
            for (int i = 0; i < N4; i += kUnroll) {
#pragma unroll
                for (int j = 0; j < kUnroll; ++j) {
                    const float4 u = lds128(bc_addr + 32u * (i + j)), m = lds128(bc_addr + 32u * (i + j) + 16u);
                    const float2 PX = make_float2(u.x, u.y), PY = make_float2(u.z, u.w), nx = make_float2(m.x, m.y), ny = make_float2(m.z, m.w);
                    const float2 w = L::fma(H[6], PX, L::fma(H[7], PY, one));
                    const float2 ex = L::fma(nx, w, L::fma(H[0], PX, L::fma(H[1], PY, H[2])));
                    const float2 ey = L::fma(ny, w, L::fma(H[3], PX, L::fma(H[4], PY, H[5])));
                    const float2 e = L::fma(ex, ex, L::mul(ey, ey));
                    const float2 tn = L::fma(w, w, L::neg(e));
                    o0 += __float_as_uint(tn.x) >> 31;
                    o1 += __float_as_uint(tn.y) >> 31;
                }
            }
            c0 = N4 - (int)o0;
            c1 = N4 - (int)o1;
        } else {
        for (int i = 0; i < N4; i += kUnroll) {
#pragma unroll
            for (int j = 0; j < kUnroll; ++j) {
                const float4 u = lds128(bc_addr + 32u * (i + j)), m = lds128(bc_addr + 32u * (i + j) + 16u);
                const float2 PX = make_float2(u.x, u.y), PY = make_float2(u.z, u.w), nx = make_float2(m.x, m.y), ny = make_float2(m.z, m.w);
                const float2 w = L::fma(H[6], PX, L::fma(H[7], PY, one));
                const float2 ex = L::fma(nx, w, L::fma(H[0], PX, L::fma(H[1], PY, H[2])));
                const float2 ey = L::fma(ny, w, L::fma(H[3], PX, L::fma(H[4], PY, H[5])));
                const float2 e = L::fma(ex, ex, L::mul(ey, ey));
                const float2 t = L::fma(L::neg(w), w, e);
                count_if_le0(c0, t.x);
                count_if_le0(c1, t.y);
            }
        }
        }
        if (v0 && fin.x && (c0 > best_cnt || (c0 == best_cnt && ha < best_h))) { best_cnt = c0; best_h = ha; }
        if (v1 && fin.y && (c1 > best_cnt || (c1 == best_cnt && hb < best_h))) { best_cnt = c1; best_h = hb; }
    };
    for (int base = warp * 32;; base += kFixedThreads) {
        if (kDynamic) {
            if (lane == 0) base = atomicAdd(&s_next, 32);
            base = __shfl_sync(kFull, base, 0);
        }
        if (base >= a.K) break;
        const int h = base + lane;
        int idx[4];
        bool ok = fixedk_sample(a, f, key, h, N, idx);
        float2 V[4], W[4];  // V = (X', x'), W = (Y', y')
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 p = lds128(ps_addr + 128u * idx[k]);
            V[k] = make_float2(p.x, p.y);
            W[k] = make_float2(p.z, p.w);
        }
        // (S, D): oriented areas of the triples (0,1,2) (1,2,3) (0,2,3) (0,1,3) on the image and the pitch side
        const float2 a012 = det3_v<float2>(V[0], W[0], V[1], W[1], V[2], W[2]);
        const float2 a123 = det3_v<float2>(V[1], W[1], V[2], W[2], V[3], W[3]);
        const float2 a023 = det3_v<float2>(V[0], W[0], V[2], W[2], V[3], W[3]);
        const float2 a013 = det3_v<float2>(V[0], W[0], V[1], W[1], V[3], W[3]);
        const float tiny = 1e-6f;
        ok = ok && fabsf(a123.x) > tiny && fabsf(a023.x) > tiny && fabsf(a013.x) > tiny && fabsf(a123.y) > tiny &&
             fabsf(a023.y) > tiny && fabsf(a013.y) > tiny;
        // orientation kept by all four triples or flipped by all four
        const bool n0 = fmul(a012.x, a012.y) < 0.f, n1 = fmul(a123.x, a123.y) < 0.f, n2 = fmul(a023.x, a023.y) < 0.f,
                   n3 = fmul(a013.x, a013.y) < 0.f;
        ok = ok && n0 == n1 && n1 == n2 && n2 == n3;
        const unsigned bal = __ballot_sync(kFull, ok);
        if (ok) {
            const int pos = qn + __popc(bal & ((1u << lane) - 1u));
            qi[pos] = (uint32_t)idx[0] | ((uint32_t)idx[1] << 8) | ((uint32_t)idx[2] << 16) | ((uint32_t)idx[3] << 24);
            qh[pos] = h;
        }
        qn += __popc(bal);
        __syncwarp();
        if (qn >= 64) {
            solve_and_score(qi, qh, lane, lane + 32, true, true);
            __syncwarp();
            // move the leftover entries [64, qn) to the front (at most 31 of them)
            const int left = qn - 64;
            uint32_t t0 = 0;
            int g0 = 0;
            if (lane < left) { t0 = qi[64 + lane]; g0 = qh[64 + lane]; }
            __syncwarp();
            if (lane < left) { qi[lane] = t0; qh[lane] = g0; }
            qn = left;
            __syncwarp();
        }
    }
    // pool the leftovers (< 64 per warp) and share them out again in batches of 64
    {
        int off = 0;
        if (lane == 0) off = atomicAdd(&s_pool, qn);
        off = __shfl_sync(kFull, off, 0);
        if (lane < qn) { s_pi[off + lane] = qi[lane]; s_ph[off + lane] = qh[lane]; }
        if (lane + 32 < qn) { s_pi[off + lane + 32] = qi[lane + 32]; s_ph[off + lane + 32] = qh[lane + 32]; }
    }
    __syncthreads();
    {
        const int total = s_pool, e0 = warp * 64 + lane, e1 = e0 + 32;
        if (warp * 64 < total) {  // unused slots score a copy of entry 0 and are discarded
            const bool v0 = e0 < total, v1 = e1 < total;
            solve_and_score(s_pi, s_ph, v0 ? e0 : 0, v1 ? e1 : 0, v0, v1);
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
    if (warp != 0) return;
    c = s_cnt[0]; hh = s_hyp[0];
#pragma unroll
    for (int w2 = 1; w2 < kFixedWarps; ++w2)
        if (s_cnt[w2] > c || (s_cnt[w2] == c && s_hyp[w2] < hh)) { c = s_cnt[w2]; hh = s_hyp[w2]; }
    const bool have = c > 3;  // cv2: goodCount > max(maxGoodCount, modelPoints - 1)
    // The winner's model is recomputed from its sample (same arithmetic, same bits) instead of being carried through
    // the scoring loop in registers; then back to image -> pitch units, and its inlier list in OpenCV's float scoring
    // (lanes are points, ballots make the mask) is what the refit starts from.
    double Hd[9];
    uint64_t pm = 0;
    if (have) {
        int idx[4];
        fixedk_sample(a, f, key, hh, N, idx);
        float p[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float4 q = s_ps[idx[k]][0];
            p[k][0] = q.x; p[k][1] = q.z; p[k][2] = q.y; p[k][3] = q.w;
        }
        float Hn[8];
        fixedk_hypothesis(p, Hn);
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

// ------------------------------------------------------------------------------------------------
// refit, warp-cooperative: the same algorithm as refit_on_inliers() (geometry_core.cuh), one WARP per
// frame.  Lanes are points for every sum over the correspondences (lane l owns points l and l+32;
// sums by shuffle reduction) and rows for the small dense solves (row-parallel Gaussian elimination
// with partial pivoting in shared memory).  Cuts the latency of the one-thread-per-frame kernel by
// ~20x at clip sizes where there are fewer frames than the GPU has thread slots.
// ------------------------------------------------------------------------------------------------
constexpr int kRefitWarps = 4;

struct RefitShared {
    PointList pl;
    double A[81];
    double aug[10 * 11];
    double vec[12];
    double sums[32];             // the 32 warp totals of warp_sum32 (24 block sums + 8 riders)
};

// Gauss-Jordan elimination with partial pivoting of an n x n system, rows in REGISTERS: lane r < n holds row r of
// the augmented n x (n+1) array.  Step k picks the largest remaining |entry| of column k (three warp reductions on
// the bit pattern -- non-negative doubles order like unsigned integers -- with the lowest row position on ties, like
// the scalar gauss_solve), broadcasts the pivot row by shuffles and clears column k in every other row.  Rows are
// not moved when they are "swapped": each lane tracks the position its row would have, which is all the pivot
// order and the result need.  Same pivot order and the same operations per element as the one-thread-per-frame
// code, which is what keeps the two kernels -- and cv2 -- together on poorly conditioned fits.  x[0..n) goes to
// shared memory.  Returns false (uniformly) on a zero / non-finite pivot.  (The shared-memory version this
// replaces spent two thirds of its instructions on element indices, barriers and the lane id.)
template <int n>
__device__ __forceinline__ bool warp_row_solve(double (&row)[n + 1], double* x, int lane) {
    const bool is_row = lane < n;
    int pos = lane;
    double diag = 1.0;   // this row's pivot element, once it has been the pivot row
    // The row is shifted left by one element per step, so the current column is always row[0]: the loop body has
    // static register indices without being unrolled n times (the unrolled form was instruction-fetch bound).
#pragma unroll 1
    for (int k = 0; k < n; ++k) {
        const bool eligible = is_row && pos >= k;
        const unsigned long long bits = eligible ? (unsigned long long)__double_as_longlong(fabs(row[0])) : 0ull;
        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
        const unsigned mhi = __reduce_max_sync(kFull, hi);
        bool cand = eligible && hi == mhi;
        const unsigned mlo = __reduce_max_sync(kFull, cand ? lo : 0u);
        cand = cand && lo == mlo;
        const unsigned ppos = __reduce_min_sync(kFull, cand ? (unsigned)pos : 99u);
        const bool is_piv = cand && (unsigned)pos == ppos;
        const double best = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
        if (!(best > 0.0) || !isfinite(best)) return false;
        const int blane = __ffs(__ballot_sync(kFull, is_piv)) - 1;
        if (pos == k) pos = (int)ppos;      // the row that sat at position k takes the pivot row's place ...
        else if (is_piv) pos = k;           // ... and the pivot row moves to position k
        const double inv = 1.0 / shfl_f64(row[0], blane);
        const double f = row[0] * inv;
        if (is_piv) diag = row[0];
        const bool upd = is_row && !is_piv;
#pragma unroll
        for (int j = 1; j <= n; ++j) {
            const double pj = shfl_f64(row[j], blane);
            row[j - 1] = upd ? row[j] - f * pj : row[j];   // eliminate and shift; columns already used carry zeros along
        }
        row[n] = 0.0;
    }
    // after n shifts the right-hand side sits in row[0]
    if (is_row) x[pos] = row[0] / diag;
    __syncwarp();
    return true;
}

// sums over the inlier points of (b b^T) (x) C-terms: 24 block sums -> the 9x9 matrix (see
// dlt_normal_matrix / lm_linearise for the block structure); v (9) and two scalars ride along.
struct PointSums {
    double bb[6], bx[6], by[6], bq[6];
};

__device__ __forceinline__ void accumulate_point(PointSums& s, const double* b, double cx, double cy, double q) {
    int e = 0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = a; c < 3; ++c, ++e) {
            const double p = b[a] * b[c];
            s.bb[e] += p; s.bx[e] += p * cx; s.by[e] += p * cy; s.bq[e] += p * q;
        }
}

// 32 warp sums at once: lane l ends up with the total over all lanes of v[l].  Same additions in the same order as 32
// butterflies `v += shfl_xor(v, m)`, m = 16 ... 1 (so the totals are bit-identical to warp_sum_f64's), but at offset m each
// lane only keeps the half of the values whose index bit matches its own lane bit and trades the other half: 31 exchanges
// instead of 160.  The per-frame refit spent half of its instructions in those butterflies.
__device__ __forceinline__ double warp_sum32(double (&v)[32], int lane) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const bool up = (lane & m) != 0;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            const double lo = v[i], hi = v[i + m];
            const double keep = up ? hi : lo, send = up ? lo : hi;
            v[i] = keep + shfl_xor_f64(send, m);
        }
    }
    return v[0];
}

// slots of the 32 totals: 0-5 bb, 6-11 bx, 12-17 by, 18-23 bq, 24-31 riders of the caller
__device__ __forceinline__ void pack_point_sums(const PointSums& s, double (&v)[32]) {
#pragma unroll
    for (int e = 0; e < 6; ++e) { v[e] = s.bb[e]; v[6 + e] = s.bx[e]; v[12 + e] = s.by[e]; v[18 + e] = s.bq[e]; }
}

// totals (shared memory, written by every lane for its own slot) -> the 9x9 matrix
__device__ __forceinline__ void store_matrix_from_sums(const double* sums, double* A) {
    const int lane = threadIdx.x & 31;
    for (int t = lane; t < 81; t += 32) {
        const int r = t / 9, c = t % 9;
        const int br = r / 3, bc = c / 3, a = r % 3, cc = c % 3;
        const int lo = a < cc ? a : cc, hi = a < cc ? cc : a;
        const int e = lo == 0 ? hi : (lo == 1 ? 2 + hi : 5);
        double v;
        if (br == bc) v = (br == 2) ? sums[18 + e] : sums[e];
        else if (br + bc == 1) v = 0.0;
        else v = -((br == 0 || bc == 0) ? sums[6 + e] : sums[12 + e]);
        A[t] = v;
    }
    __syncwarp();
}

__global__ void __launch_bounds__(kRefitWarps * 32, 4) refit_warp_kernel(FitArgs a) {
    __shared__ RefitShared s_all[kRefitWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kRefitWarps + warp;
    if (f >= a.F) return;
    if (a.status[f] != EGL_FIT_OK) {
        if (lane == 0) a.inlier_mask[f] = 0;
        return;
    }
    RefitShared& sh = s_all[warp];
    uint64_t used;
    const int N = gather_points_warp(a, f, sh.pl, &used);
    uint64_t pm = a.inlier_mask[f];  // position bits of the RANSAC winner's inliers
    double H[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) H[i] = a.H[(size_t)f * 9 + i];
    __syncwarp();
    int count = a.info[4 * f + 1];
    if (N > 4 && __popcll(pm) <= kLmExactMaxPoints) {
        // Few inliers: nothing for 32 lanes to share, and the undamped LM steps need OpenCV's eigenvalue
        // threshold (eig_threshold_solve9) to stay with cv2 on these numerically singular fits.  Lane 0 runs the
        // scalar code of geometry_core.cuh -- the same the one-thread-per-frame kernel and the host check run.
        if (lane == 0) {
            static_assert(offsetof(RefitShared, vec) - offsetof(RefitShared, A) >= 191 * sizeof(double) &&
                          sizeof(((RefitShared*)0)->vec) >= sizeof(double), "A, aug, vec are used as one 192-double scratch");
            double* scratch = sh.A;  // A[81] + aug[110] + vec[0]: unused on this path until the results are parked below
            uint64_t fm = 0;
            const int c = refit_on_inliers(H, sh.pl.sx, sh.pl.sy, sh.pl.dx, sh.pl.dy, N, pm, a.thr_sq, &fm, scratch);
#pragma unroll
            for (int i = 0; i < 9; ++i) a.H[(size_t)f * 9 + i] = H[i];
            sh.vec[0] = __longlong_as_double((long long)fm);
            sh.vec[1] = (double)c;
        }
        __syncwarp();
        pm = (uint64_t)__double_as_longlong(sh.vec[0]);
        count = (int)sh.vec[1];
    } else if (N > 4) {
        const int m = __popcll(pm);
        // this lane's (up to two) inlier points
        bool act[2];
        double Mx[2], My[2], mx[2], my[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = lane + 32 * h;
            act[h] = i < N && ((pm >> i) & 1ull);
            Mx[h] = act[h] ? (double)sh.pl.sx[i] : 0.0; My[h] = act[h] ? (double)sh.pl.sy[i] : 0.0;
            mx[h] = act[h] ? (double)sh.pl.dx[i] : 0.0; my[h] = act[h] ? (double)sh.pl.dy[i] : 0.0;
        }
        // ---- runKernel on the inliers: normalisation, 9x9 normal matrix, smallest eigenvector -------
        const double cMx = warp_sum_f64(Mx[0] + Mx[1]) / m, cMy = warp_sum_f64(My[0] + My[1]) / m;
        const double cmx = warp_sum_f64(mx[0] + mx[1]) / m, cmy = warp_sum_f64(my[0] + my[1]) / m;
        double aMx = 0, aMy = 0, amx = 0, amy = 0;
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (act[h]) { aMx += fabs(Mx[h] - cMx); aMy += fabs(My[h] - cMy); amx += fabs(mx[h] - cmx); amy += fabs(my[h] - cmy); }
        aMx = warp_sum_f64(aMx); aMy = warp_sum_f64(aMy); amx = warp_sum_f64(amx); amy = warp_sum_f64(amy);
        bool have_ls = !(fabs(amx) < DBL_EPSILON || fabs(amy) < DBL_EPSILON || fabs(aMx) < DBL_EPSILON || fabs(aMy) < DBL_EPSILON);
        if (have_ls) {
            const double sMx = m / aMx, sMy = m / aMy, smx = m / amx, smy = m / amy;
            PointSums ps = {};
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (act[h]) {
                    const double x = (mx[h] - cmx) * smx, y = (my[h] - cmy) * smy;
                    const double b[3] = {(Mx[h] - cMx) * sMx, (My[h] - cMy) * sMy, 1.0};
                    accumulate_point(ps, b, x, y, x * x + y * y);
                }
            {
                double t32[32];
                pack_point_sums(ps, t32);
#pragma unroll
                for (int i = 24; i < 32; ++i) t32[i] = 0.0;
                sh.sums[lane] = warp_sum32(t32, lane);
                __syncwarp();
                store_matrix_from_sums(sh.sums, sh.A);
            }
            // inverse iteration (same as smallest_eigvec9)
            double tr = 0;
            for (int i = 0; i < 9; ++i) tr += sh.A[i * 9 + i];
            have_ls = tr > 0.0 && isfinite(tr);
            const double sigma = tr * 1e-15;
            double y_l = lane < 9 ? 1.0 / (1.37 + lane) : 0.0;  // lane i < 9 holds y[i]
            double prev = 1e300;
            for (int it = 0; it < 48 && have_ls; ++it) {
                double row[10];
#pragma unroll
                for (int c = 0; c < 9; ++c) row[c] = lane < 9 ? sh.A[lane * 9 + c] + (lane == c ? sigma : 0.0) : 0.0;
                row[9] = y_l;
                if (!warp_row_solve<9>(row, sh.vec, lane)) { have_ls = false; break; }
                double z = lane < 9 ? sh.vec[lane] : 0.0;
                const double nrm = warp_sum_f64(z * z);
                // sign: component of largest magnitude positive (lowest index on ties)
                double big = fabs(z);
                int bi = lane < 9 ? lane : 99;
                double bz = z;
#pragma unroll
                for (int mm = 16; mm > 0; mm >>= 1) {
                    const double ob = shfl_xor_f64(big, mm), oz = shfl_xor_f64(bz, mm);
                    const int oi = __shfl_xor_sync(kFull, bi, mm);
                    if (ob > big || (ob == big && oi < bi)) { big = ob; bi = oi; bz = oz; }
                }
                if (!(nrm > 0.0) || !isfinite(nrm)) { have_ls = false; break; }
                const double sc = (bz < 0 ? -1.0 : 1.0) / sqrt(nrm);
                z *= sc;
                const double diff = warp_max_f64(lane < 9 ? fabs(z - y_l) : 0.0);
                y_l = z;
                __syncwarp();
                if (diff <= 1e-13 || (diff <= 1e-10 && diff >= prev)) break;
                prev = diff;
            }
            if (have_ls) {
                double h0[9];
#pragma unroll
                for (int i = 0; i < 9; ++i) h0[i] = shfl_f64(y_l, i);
                const double norm[8] = {cMx, cMy, cmx, cmy, sMx, sMy, smx, smy};
                double Hk[9];
                dlt_denormalise(h0, norm, Hk);
                bool fin = true;
#pragma unroll
                for (int i = 0; i < 9; ++i) fin &= isfinite(Hk[i]);
                if (fin)
#pragma unroll
                    for (int i = 0; i < 9; ++i) H[i] = Hk[i];
            }
        }
        // ---- LM polish (LMSolverImpl::run, 9 parameters, <= 10 iterations) ---------------------------
        double x[9], v[9], D[9], d[9], xd[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) x[i] = H[i];
        double S = 0, rmax = 0;
        auto linearise = [&](const double* hh) {
            PointSums ps = {};
            double vx[3] = {0, 0, 0}, vy[3] = {0, 0, 0}, vq[3] = {0, 0, 0}, s = 0, rm = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (act[h]) {
                    double ww = hh[6] * Mx[h] + hh[7] * My[h] + hh[8];
                    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                    const double xi = (hh[0] * Mx[h] + hh[1] * My[h] + hh[2]) * ww;
                    const double yi = (hh[3] * Mx[h] + hh[4] * My[h] + hh[5]) * ww;
                    const double rx = xi - mx[h], ry = yi - my[h];
                    s += rx * rx + ry * ry;
                    rm = fmax(rm, fmax(fabs(rx), fabs(ry)));
                    const double b[3] = {Mx[h] * ww, My[h] * ww, ww};
                    accumulate_point(ps, b, xi, yi, xi * xi + yi * yi);
                    const double g = xi * rx + yi * ry;
#pragma unroll
                    for (int q = 0; q < 3; ++q) { vx[q] += b[q] * rx; vy[q] += b[q] * ry; vq[q] += b[q] * g; }
                }
            double t32[32];
            pack_point_sums(ps, t32);
#pragma unroll
            for (int q = 0; q < 3; ++q) { t32[24 + q] = vx[q]; t32[27 + q] = vy[q]; }
            t32[30] = vq[0]; t32[31] = vq[1];
            __syncwarp();                       // the previous readers of sh.sums are done
            sh.sums[lane] = warp_sum32(t32, lane);
            const double vq2 = warp_sum_f64(vq[2]);
            S = warp_sum_f64(s);
            rmax = warp_max_f64(rm);
            __syncwarp();
            store_matrix_from_sums(sh.sums, sh.A);
#pragma unroll
            for (int q = 0; q < 3; ++q) { v[q] = sh.sums[24 + q]; v[3 + q] = sh.sums[27 + q]; }
            v[6] = -sh.sums[30]; v[7] = -sh.sums[31]; v[8] = -vq2;
        };
        auto cost = [&](const double* hh) {
            double s = 0;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (act[h]) {
                    double ww = hh[6] * Mx[h] + hh[7] * My[h] + hh[8];
                    ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
                    const double rx = (hh[0] * Mx[h] + hh[1] * My[h] + hh[2]) * ww - mx[h];
                    const double ry = (hh[3] * Mx[h] + hh[4] * My[h] + hh[5]) * ww - my[h];
                    s += rx * rx + ry * ry;
                }
            return warp_sum_f64(s);
        };
        // element `lane` of a 9-vector every lane holds in registers: a select tree on the lane bits (static register
        // indices; staging the vector through shared memory cost a store -> barrier -> load round trip per use)
        auto pick9 = [&](const double* q) -> double {
            const bool b0 = lane & 1, b1 = lane & 2, b2 = lane & 4, b3 = lane & 8;
            const double t0 = b0 ? q[1] : q[0], t1 = b0 ? q[3] : q[2], t2 = b0 ? q[5] : q[4], t3 = b0 ? q[7] : q[6];
            const double u0 = b1 ? t1 : t0, u1 = b1 ? t3 : t2;
            const double w = b2 ? u1 : u0;
            return b3 ? q[8] : w;
        };
        // gauge-fixed solve A d = rhs (A singular along x): bordered system [[A, s x],[s x^T, 0]] [d; mu] = [rhs; 0]
        // (solve_gauge_fixed9 of geometry_core.cuh, parallelised), result in sh.vec[0..9)
        auto solve_gauge = [&](const double* rhs) -> bool {
            double xn = 0, dmx = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) { xn += x[i] * x[i]; dmx = fmax(dmx, fabs(sh.A[i * 9 + i])); }
            if (!(xn > 0.0) || !(dmx > 0.0)) return false;
            const double sc = dmx / sqrt(xn);
            double row[11];
#pragma unroll
            for (int c = 0; c < 9; ++c) row[c] = lane < 9 ? sh.A[lane * 9 + c] : (lane == 9 ? sc * x[c] : 0.0);
            row[9] = lane < 9 ? sc * pick9(x) : 0.0;
            row[10] = lane < 9 ? pick9(rhs) : 0.0;
            return warp_row_solve<10>(row, sh.vec, lane);
        };
        linearise(x);
#pragma unroll
        for (int i = 0; i < 9; ++i) D[i] = sh.A[i * 9 + i];
        const double Rlo = 0.25, Rhi = 0.75;
        double lambda = 1, lc = 0.75;
        int iter = 0;
        for (;;) {
            bool ok;
            if (lambda > 0) {
                double row[10];
#pragma unroll
                for (int c = 0; c < 9; ++c) row[c] = lane < 9 ? sh.A[lane * 9 + c] + (lane == c ? lambda * D[c] : 0.0) : 0.0;
                row[9] = lane < 9 ? pick9(v) : 0.0;
                ok = warp_row_solve<9>(row, sh.vec, lane);
            } else {
                ok = solve_gauge(v);
            }
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 9; ++i) { d[i] = ok ? sh.vec[i] : 0.0; xd[i] = x[i] - d[i]; }
            const double Sd = cost(xd);
            double dS = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                double t = 2.0 * v[i];
#pragma unroll
                for (int k = 0; k < 9; ++k) t -= sh.A[i * 9 + k] * d[k];
                dS += d[i] * t;
            }
            const double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);
            if (R > Rhi) {
                lambda *= 0.5;
                if (lambda < lc) lambda = 0;
            } else if (R < Rlo) {
                double t = 0;
#pragma unroll
                for (int i = 0; i < 9; ++i) t += d[i] * v[i];
                double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
                nu = fmin(fmax(nu, 2.), 10.);
                if (lambda == 0) {
                    double maxval = DBL_EPSILON;
                    for (int k = 0; k < 9; ++k) {
                        double e[9];
#pragma unroll
                        for (int i = 0; i < 9; ++i) e[i] = (i == k) ? 1.0 : 0.0;
                        __syncwarp();
                        if (solve_gauge(e)) maxval = fmax(maxval, fabs(sh.vec[k]));
                        __syncwarp();
                    }
                    lambda = lc = 1. / maxval;
                    nu *= 0.5;
                }
                lambda *= nu;
            }
            double dmax = 0;
#pragma unroll
            for (int i = 0; i < 9; ++i) dmax = fmax(dmax, fabs(d[i]));
            __syncwarp();
            if (Sd < S) {
#pragma unroll
                for (int i = 0; i < 9; ++i) x[i] = xd[i];
                linearise(x);
            }
            iter++;
            const bool proceed = iter < 10 && dmax >= (double)FLT_EPSILON && rmax >= (double)FLT_EPSILON;
            if (!proceed) break;
        }
        const double sc = fabs(x[8]) > (double)FLT_EPSILON ? 1. / x[8] : 1.;
#pragma unroll
        for (int i = 0; i < 9; ++i) H[i] = x[i] * sc;
        // ---- final mask from the refined H: lanes are points, ballots make the mask ------------------
        float Hf[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) Hf[k] = (float)H[k];
        pm = 0;
        for (int pass = 0; pass < 2; ++pass) {
            const int i = pass * 32 + lane;
            const bool in = i < N && reproj_err_f32(Hf, sh.pl.sx[i], sh.pl.sy[i], sh.pl.dx[i], sh.pl.dy[i]) <= a.thr_sq;
            pm |= (uint64_t)__ballot_sync(kFull, in) << (32 * pass);
        }
        count = __popcll(pm);
#pragma unroll
        for (int i = 0; i < 9; ++i)
            if (lane == i) a.H[(size_t)f * 9 + i] = H[i];
    }
    // position bits -> channel bits
    uint64_t cm = 0;
    for (int pass = 0; pass < 2; ++pass) {
        const int i = pass * 32 + lane;
        if (i < N && ((pm >> i) & 1ull)) cm |= 1ull << sh.pl.ch[i];
    }
    unsigned lo = __reduce_or_sync(kFull, (unsigned)cm), hi = __reduce_or_sync(kFull, (unsigned)(cm >> 32));
    if (lane == 0) {
        a.inlier_mask[f] = ((uint64_t)hi << 32) | lo;
        a.info[4 * f + 1] = count;
    }
}

// ------------------------------------------------------------------------------------------------
// The RHO and LMEDS legs of coordinate_model.py:354-357 for frames whose RANSAC leg returned no model.
// Rare frames (RANSAC gives up only when no sample passes checkSubset or no model reaches 4 inliers), both
// estimators are sequential by construction (every sample depends on the SPRT / iteration bounds the
// previous one left); the scalar code is shared with the host build.
//   status -> EGL_FIT_OK, H / inlier_mask / info[1] as for the RANSAC leg, info[2] = EGL_FIT_LEG_RHO / _LMEDS
// ------------------------------------------------------------------------------------------------
constexpr int kCascadeWarps = 4;

__global__ void __launch_bounds__(kCascadeWarps * 32) cascade_kernel(FitArgs a) {
    // One WARP per frame (hard frames that are neighbours in the clip must not share a warp: they would run one after
    // the other).  RHO is a chain -- every sample depends on the SPRT state and the iteration bound the previous one
    // left -- and runs on lane 0.  LMedS rates a fixed sample sequence: lane 0 walks cv::RNG for the sequence, all lanes
    // rate samples (one 9x9 Jacobi eigen-decomposition each), a warp arg-min picks the first smallest median.
    __shared__ PointList s_pl[kCascadeWarps];
    __shared__ uint8_t s_smp[kCascadeWarps][kLmedsMaxIters][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kCascadeWarps + warp;
    if (f >= a.F || a.status[f] != EGL_FIT_NO_MODEL) return;
    PointList& pl = s_pl[warp];
    uint64_t used;
    const int N = gather_points_warp(a, f, pl, &used);
    if (N <= 4) return;  // findHomography solves exactly four points without a robust method: every leg fails alike
    double H[9];
    uint64_t pm = 0;
    int leg = EGL_FIT_LEG_RHO, count = 0;
    if (lane == 0) {
        float Hf[9];
        count = rho_fit(pl.sx, pl.sy, pl.dx, pl.dy, N, Hf, &pm);
        for (int i = 0; i < 9; ++i) H[i] = (double)Hf[i];
    }
    count = __shfl_sync(kFull, count, 0);
    if (count <= 0) {
        leg = EGL_FIT_LEG_LMEDS;
        int n_smp = 0;
        if (lane == 0) n_smp = lmeds_draw_samples(pl.sx, pl.sy, pl.dx, pl.dy, N, a.confidence, s_smp[warp]);
        n_smp = __shfl_sync(kFull, n_smp, 0);
        __syncwarp();
        if (n_smp < 0) return;
        double scratch[192];
        double bm = DBL_MAX, bH[9];
        int bt = 0x7fffffff;
        for (int i = 0; i < 9; ++i) bH[i] = 0.0;
        for (int t = lane; t < n_smp; t += 32) {
            double Hm[9], med;
            if (lmeds_rate_sample(pl.sx, pl.sy, pl.dx, pl.dy, N, s_smp[warp][t], Hm, &med, scratch) && med < bm) {
                bm = med;
                bt = t;
                for (int i = 0; i < 9; ++i) bH[i] = Hm[i];
            }
        }
        // `median < minMedian` in sequence order = smallest median, ties to the earliest sample
        double wm = bm;
        int wt = bt;
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) {
            const double om = shfl_xor_f64(wm, m);
            const int ot = __shfl_xor_sync(kFull, wt, m);
            if (om < wm || (om == wm && ot < wt)) { wm = om; wt = ot; }
        }
        if (wt == 0x7fffffff) return;  // no sample gave a model
        const int src = wt & 31;
#pragma unroll
        for (int i = 0; i < 9; ++i) bH[i] = shfl_f64(bH[i], src);
        if (lane == 0) count = lmeds_finish(pl.sx, pl.sy, pl.dx, pl.dy, N, bH, wm, H, &pm, scratch);
        count = __shfl_sync(kFull, count, 0);
        if (count < 0) return;
    }
    if (lane != 0) return;
    uint64_t cm = 0;
    for (int i = 0; i < N; ++i)
        if ((pm >> i) & 1ull) cm |= 1ull << pl.ch[i];
    for (int i = 0; i < 9; ++i) a.H[(size_t)f * 9 + i] = H[i];
    a.inlier_mask[f] = cm;
    a.status[f] = EGL_FIT_OK;
    a.info[4 * f + 1] = count;
    a.info[4 * f + 2] = leg;
}

}  // namespace egl

using namespace egl;

// Stream-ordered scratch of the fixed-K mode comes from a pool of this library's own that never returns memory to the
// driver (release threshold = max): with the default pool every call paid a fresh 100 MB allocation (4 ms at 50 k frames).
static cudaMemPool_t scratch_pool() {
    static cudaMemPool_t pools[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    if (!pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPool_t p = nullptr;
        if (cudaMemPoolCreate(&p, &props) != cudaSuccess) return nullptr;
        unsigned long long keep = ~0ull;
        cudaMemPoolSetAttribute(p, cudaMemPoolAttrReleaseThreshold, &keep);
        pools[dev] = p;
    }
    return pools[dev];
}

static int fit_impl(const char* who, const int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count, int F, int mode, int K,
                    const uint8_t* hyp, uint64_t seed, double thr, double confidence, double* H, uint64_t* used_mask,
                    uint64_t* inlier_mask, int32_t* status, int32_t* info, const uint8_t* sched, const uint8_t* retry, void* stream,
                    const float* kp_sub = nullptr) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(kp_xy && kp_order && kp_count && H && used_mask && inlier_mask && status && info, EGL_ERR_NULL, "%s: null pointer", who);
    EGL_REQUIRE(F >= 0 && K >= 1, EGL_ERR_SHAPE, "%s: need F >= 0 and K >= 1 (F=%d K=%d)", who, F, K);
    EGL_REQUIRE(mode == EGL_FIT_CV2_COMPAT || mode == EGL_FIT_FIXED_K, EGL_ERR_MODE, "%s: unknown mode %d", who, mode);
    EGL_REQUIRE(confidence > 0 && confidence < 1, EGL_ERR_SHAPE, "%s: confidence must be in (0,1)", who);
    if (!(thr > 0)) thr = 3.0;  // findHomography: ransacReprojThreshold <= 0 -> 3
    FitArgs a{kp_xy, kp_order, kp_count, F, K, hyp, seed, (float)(thr * thr), confidence, H, used_mask, inlier_mask, status, info,
              sched, retry, kp_sub};
    cudaStream_t s = (cudaStream_t)stream;
    if (mode == EGL_FIT_CV2_COMPAT) {
        ransac_cv2_kernel<<<(F + kCv2Warps - 1) / kCv2Warps, kCv2Warps * 32, 0, s>>>(a);
    } else {
        // scratch for the prepared frames: stream-ordered, so the free below only takes effect after the kernels
        FixedKPrep* prep = nullptr;
        cudaMemPool_t pool = scratch_pool();
        int rc0 = cuda_status(pool ? cudaMallocFromPoolAsync((void**)&prep, sizeof(FixedKPrep) * (size_t)F, pool, s) : cudaErrorMemoryAllocation,
                              "egl_fit_homography: scratch allocation");
        if (rc0) return rc0;
        fixedk_prepare_kernel<<<(F + kPrepWarps - 1) / kPrepWarps, kPrepWarps * 32, 0, s>>>(a, (float)(1.0 / thr), prep);
        int fk_variant = 0;
#ifdef EGL_BENCH_VARIANTS
        static const char* fk_env = getenv("EGL_FIXEDK_VARIANT");
        fk_variant = fk_env ? atoi(fk_env) : 0;
#endif
        const float it = (float)(1.0 / thr);
        switch (fk_variant) {
#ifdef EGL_BENCH_VARIANTS
            case 1: ransac_fixedk_kernel<6, 4, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 2: ransac_fixedk_kernel<6, 8, false><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 3: ransac_fixedk_kernel<6, 8, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 4: ransac_fixedk_kernel<7, 4, false><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 5: ransac_fixedk_kernel<7, 2, false><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 6: ransac_fixedk_kernel<7, 2, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 7: ransac_fixedk_kernel<8, 2, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 8: ransac_fixedk_kernel<6, 2, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 9: ransac_fixedk_kernel<5, 8, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 10: ransac_fixedk_kernel<6, 4, false, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 11: ransac_fixedk_kernel<6, 8, false, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
            case 12: ransac_fixedk_kernel<6, 4, true, true><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
#endif
            default: ransac_fixedk_kernel<6, 4, false, false><<<F, kFixedThreads, 0, s>>>(a, it, thr, prep); break;
        }
        cudaFreeAsync(prep, s);
    }
    int rc = cuda_status(cudaGetLastError(), "egl_fit_homography: hypothesis kernel launch");
    if (rc) return rc;
    // Same algorithm, two parallelisations: one warp per frame has the lower latency (0.28 ms for a
    // 2250-frame clip, a single wave), one thread per frame -- the host-checkable scalar code of
    // geometry_core.cuh -- the higher throughput once there are more frames than resident warps
    // (measured cross-over ~10-12 k frames: 50 k frames 2.2 ms vs 3.9 ms).  In builds with
    // -DEGL_BENCH_VARIANTS, EGL_REFIT_VARIANT = 1 / 2 forces the thread / warp kernel.
    int refit_variant = 0;
#ifdef EGL_BENCH_VARIANTS
    static const char* refit_env = getenv("EGL_REFIT_VARIANT");
    refit_variant = refit_env ? atoi(refit_env) : 0;
#endif
    if (refit_variant == 1 || (refit_variant == 0 && F > 12288))
        refit_kernel<<<(F + kRefitThreads - 1) / kRefitThreads, kRefitThreads, 0, s>>>(a);
    else
        refit_warp_kernel<<<(F + kRefitWarps - 1) / kRefitWarps, kRefitWarps * 32, 0, s>>>(a);
    rc = cuda_status(cudaGetLastError(), "egl_fit_homography: refit kernel launch");
    if (rc || mode != EGL_FIT_CV2_COMPAT) return rc;
    cascade_kernel<<<(F + kCascadeWarps - 1) / kCascadeWarps, kCascadeWarps * 32, 0, s>>>(a);
    return cuda_status(cudaGetLastError(), "egl_fit_homography: cascade kernel launch");
}

extern "C" int egl_fit_homography(const int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count, int F, int mode,
                                  int K, const uint8_t* hyp, uint64_t seed, double thr, double confidence, double* H,
                                  uint64_t* used_mask, uint64_t* inlier_mask, int32_t* status, int32_t* info,
                                  void* stream) {
    return fit_impl("egl_fit_homography", kp_xy, kp_order, kp_count, F, mode, K, hyp, seed, thr, confidence, H, used_mask, inlier_mask,
                    status, info, nullptr, nullptr, stream);
}

extern "C" int egl_fit_homography_masked(const int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count, int F, int mode,
                                         int K, const uint8_t* hyp, uint64_t seed, double thr, double confidence, double* H,
                                         uint64_t* used_mask, uint64_t* inlier_mask, int32_t* status, int32_t* info,
                                         const uint8_t* sched, const uint8_t* retry, void* stream) {
    if (F > 0) EGL_REQUIRE(sched, EGL_ERR_NULL, "egl_fit_homography_masked: sched is null");
    return fit_impl("egl_fit_homography_masked", kp_xy, kp_order, kp_count, F, mode, K, hyp, seed, thr, confidence, H, used_mask,
                    inlier_mask, status, info, sched, retry, stream);
}

extern "C" int egl_fit_homography_subpixel(const float* kp_sub, const int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count,
                                           int F, int mode, int K, const uint8_t* hyp, uint64_t seed, double thr, double confidence,
                                           double* H, uint64_t* used_mask, uint64_t* inlier_mask, int32_t* status, int32_t* info,
                                           void* stream) {
    if (F > 0) EGL_REQUIRE(kp_sub, EGL_ERR_NULL, "egl_fit_homography_subpixel: kp_sub is null");
    return fit_impl("egl_fit_homography_subpixel", kp_xy, kp_order, kp_count, F, mode, K, hyp, seed, thr, confidence, H, used_mask,
                    inlier_mask, status, info, nullptr, nullptr, stream, kp_sub);
}
