// F1: line-intersection keypoint synthesis between decode and fit (replaces
// CoordinateModel._synthesize_keypoints_with_line_intersections, eagle/models/coordinate_model.py
// :140-186 with the helpers at :76-138).  One thread per frame: at most 38 tiny line fits and a few
// dozen 2x2 solves -- latency-bound bookkeeping that stays on the device so that the decode -> fit
// chain needs no host round trip.
#include "common.cuh"
#include "geometry_core.cuh"

namespace egl {

#include "line_families.inc"

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

}  // namespace egl

using namespace egl;

extern "C" int egl_synthesize_keypoints(int32_t* kp_xy, uint8_t* kp_order, int32_t* kp_count, int F, int max_new,
                                        void* stream) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(kp_xy && kp_order && kp_count, EGL_ERR_NULL, "egl_synthesize_keypoints: null pointer");
    EGL_REQUIRE(F >= 0 && max_new >= 0, EGL_ERR_SHAPE, "egl_synthesize_keypoints: bad arguments");
    if (F == 0 || max_new == 0) return 0;
    synthesize_kernel<<<(F + 63) / 64, 64, 0, (cudaStream_t)stream>>>(kp_xy, kp_order, kp_count, F, max_new);
    return cuda_status(cudaGetLastError(), "egl_synthesize_keypoints: kernel launch");
}
