// F1: line-intersection keypoint synthesis between decode and fit (replaces
// CoordinateModel._synthesize_keypoints_with_line_intersections, eagle/models/coordinate_model.py
// :140-186 with the helpers at :76-138).  At most 38 tiny line fits and a few dozen 2x2 solves per
// frame -- latency-bound bookkeeping that stays on the device so that the decode -> fit chain needs no
// host round trip.  One warp per frame (synthesize_warp_kernel); synthesize_kernel is the same rules
// run by one thread per frame through the host-checkable scalar code.
#include <stdlib.h>

#include "common.cuh"
#include "geometry_core.cuh"

namespace egl {

#include "line_families.inc"

#ifdef EGL_BENCH_VARIANTS
__global__ void __launch_bounds__(64) synthesize_kernel(int32_t* kp_xy, uint8_t* kp_order, int32_t* kp_count, int F,
                                                        int max_new) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int n = kp_count[2 * f];
    if (n < 2 || n > EGL_ORDER_STRIDE) return;  // reference: only when len(keypoints) >= 2 (:326)
    int32_t xy[2 * kLandmarks];
    uint8_t order[EGL_ORDER_STRIDE];
    for (int c = 0; c < 2 * kLandmarks; ++c) xy[c] = kp_xy[(size_t)f * 2 * kLandmarks + c];
    for (int j = 0; j < EGL_ORDER_STRIDE; ++j) order[j] = kp_order[(size_t)f * EGL_ORDER_STRIDE + j];
    SynthTables T{kYFamCount, &kYFam[0][0], kXFamCount, &kXFam[0][0], &kCross[0][0], kNumYFam, kNumXFam, kMaxFam};
    const int total = synthesize_keypoints(T, xy, order, n, max_new);
    if (total == n) return;
    for (int j = n; j < total && j < EGL_ORDER_STRIDE; ++j) {
        const int ch = order[j];
        kp_order[(size_t)f * EGL_ORDER_STRIDE + j] = (uint8_t)ch;
        kp_xy[((size_t)f * kLandmarks + ch) * 2] = xy[2 * ch];
        kp_xy[((size_t)f * kLandmarks + ch) * 2 + 1] = xy[2 * ch + 1];
    }
    kp_count[2 * f] = total < EGL_ORDER_STRIDE ? total : EGL_ORDER_STRIDE;
}
#endif  // EGL_BENCH_VARIANTS

// Warp-per-frame version (the product kernel): lanes are line families for the fits (19 world-y
// families, then 19 world-x families) and x-families for the crossings of each y-family; a ballot
// prefix keeps the reference's append order (y-family major) and its cap of max_new additions.
constexpr int kSynWarps = 4;

__global__ void __launch_bounds__(kSynWarps * 32) synthesize_warp_kernel(int32_t* kp_xy, uint8_t* kp_order, int32_t* kp_count, int F,
                                                                        int max_new) {
    __shared__ float s_ly[kSynWarps][kNumYFam][4], s_lx[kSynWarps][kNumXFam][4];
    static_assert(kNumYFam <= 32 && kNumXFam <= 32, "one lane per line family");
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kSynWarps + warp;
    if (f >= F) return;
    const int n = kp_count[2 * f];
    if (n < 2 || n > EGL_ORDER_STRIDE) return;  // reference: only when len(keypoints) >= 2 (:326)
    const int32_t* xy = kp_xy + (size_t)f * 2 * kLandmarks;
    uint64_t present = 0;
    for (int j = lane; j < n; j += 32) present |= 1ull << kp_order[(size_t)f * EGL_ORDER_STRIDE + j];
    present = ((uint64_t)__reduce_or_sync(kFull, (unsigned)(present >> 32)) << 32) | __reduce_or_sync(kFull, (unsigned)present);
    const uint64_t detected = present & ~kOffPlaneMask;
    float line[4];
    bool ok = lane < kNumYFam && fit_family_line(xy, detected, &kYFam[lane][0], kYFamCount[lane], line);
    if (ok)
#pragma unroll
        for (int k = 0; k < 4; ++k) s_ly[warp][lane][k] = line[k];
    const unsigned have_y = __ballot_sync(kFull, ok);
    ok = lane < kNumXFam && fit_family_line(xy, detected, &kXFam[lane][0], kXFamCount[lane], line);
    if (ok)
#pragma unroll
        for (int k = 0; k < 4; ++k) s_lx[warp][lane][k] = line[k];
    const unsigned have_x = __ballot_sync(kFull, ok);
    __syncwarp();
    int added = 0;
    for (int iy = 0; iy < kNumYFam && added < max_new; ++iy) {
        if (!((have_y >> iy) & 1u)) continue;
        bool valid = false;
        int ch = 255;
        double px = 0, py = 0;
        if (lane < kNumXFam && ((have_x >> lane) & 1u)) {
            ch = kCross[iy][lane];
            if (ch != 255 && !((present >> ch) & 1ull)) valid = intersect_lines(s_ly[warp][iy], s_lx[warp][lane], &px, &py);
        }
        const unsigned bal = __ballot_sync(kFull, valid);
        const int pos = added + __popc(bal & ((1u << lane) - 1u));
        if (valid && pos < max_new && n + pos < EGL_ORDER_STRIDE) {
            kp_xy[((size_t)f * kLandmarks + ch) * 2] = round_half_even_i32(px);
            kp_xy[((size_t)f * kLandmarks + ch) * 2 + 1] = round_half_even_i32(py);
            kp_order[(size_t)f * EGL_ORDER_STRIDE + n + pos] = (uint8_t)ch;
        }
        added = min(added + __popc(bal), max_new);
    }
    if (lane == 0 && added > 0) kp_count[2 * f] = min(n + added, EGL_ORDER_STRIDE);
}

}  // namespace egl

using namespace egl;

extern "C" int egl_synthesize_keypoints(int32_t* kp_xy, uint8_t* kp_order, int32_t* kp_count, int F, int max_new,
                                        void* stream) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(kp_xy && kp_order && kp_count, EGL_ERR_NULL, "egl_synthesize_keypoints: null pointer");
    EGL_REQUIRE(F >= 0 && max_new >= 0, EGL_ERR_SHAPE, "egl_synthesize_keypoints: bad arguments");
    if (F == 0 || max_new == 0) return 0;
#ifdef EGL_BENCH_VARIANTS
    static const char* env = getenv("EGL_SYNTH_VARIANT");  // measurement switch: 1 = one thread per frame (host-checkable code)
    if (env && atoi(env) == 1) {
        synthesize_kernel<<<(F + 63) / 64, 64, 0, (cudaStream_t)stream>>>(kp_xy, kp_order, kp_count, F, max_new);
        return cuda_status(cudaGetLastError(), "egl_synthesize_keypoints: kernel launch");
    }
#endif
    synthesize_warp_kernel<<<(F + kSynWarps - 1) / kSynWarps, kSynWarps * 32, 0, (cudaStream_t)stream>>>(kp_xy, kp_order, kp_count,
                                                                                                             F, max_new);
    return cuda_status(cudaGetLastError(), "egl_synthesize_keypoints: kernel launch");
}
