// K4: projection of every frame's foot points into pitch coordinates + visible-pitch boundaries
// (replaces the per-object cv2.perspectiveTransform loop, eagle/models/coordinate_model.py:369-392,
// and the corner projection / find_x_at_y block, :396-414 with :32-44).
//
// One launch covers all points of all frames: thread (f, p) projects point p of frame f with the
// homography row h_index[f] selects (the reference's H_use: current, else previous, else none);
// threads with p >= P - but < P + 1 - handle the frame's four corners and the boundary arithmetic.
// Latency-bound (~750 B per frame); it exists as its own kernel so that the host can re-project
// with a different h_index (reference cadence: H is refreshed only every homography_interval frames).
#include "common.cuh"
#include "geometry_core.cuh"

namespace egl {

constexpr int kPitchXMax = 105;  // PITCH_WIDTH,  coordinate_model.py:18
constexpr int kPitchYMax = 68;   // PITCH_HEIGHT, coordinate_model.py:19

__global__ void __launch_bounds__(128) project_kernel(const double* __restrict__ H, const int32_t* __restrict__ h_index,
                                                      const float* __restrict__ pts, const int32_t* __restrict__ npts, int F,
                                                      int P, int img_w, int img_h, float* out_f, int64_t* out_i, uint8_t* inb,
                                                      double* bounds) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int per = P + 1;
    const int f = (int)(t / per), p = (int)(t % per);
    if (f >= F) return;
    const int hi = h_index ? h_index[f] : f;
    double Hm[9];
    if (hi >= 0)
#pragma unroll
        for (int i = 0; i < 9; ++i) Hm[i] = H[(size_t)hi * 9 + i];
    if (p < P) {
        const size_t o = (size_t)f * P + p;
        float ox = 0.f, oy = 0.f;
        long long ix = 0, iy = 0;
        uint8_t ok = 0;
        if (hi >= 0 && p < npts[f]) {
            perspective_point(Hm, pts[2 * o], pts[2 * o + 1], &ox, &oy);
            ix = trunc_like_numpy(ox);
            iy = trunc_like_numpy(oy);
            ok = !(ix < 0 || ix > kPitchXMax || iy < 0 || iy > kPitchYMax);  // :385
        }
        out_f[2 * o] = ox;
        out_f[2 * o + 1] = oy;
        out_i[2 * o] = ix;
        out_i[2 * o + 1] = iy;
        inb[o] = ok;
        return;
    }
    // boundaries (:396-414): corners (0,0),(W,0),(0,H),(W,H) -> int -> find_x_at_y x4
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    double b[4] = {nan, nan, nan, nan};
    if (hi >= 0) {
        const float cx[4] = {0.f, (float)img_w, 0.f, (float)img_w};
        const float cy[4] = {0.f, 0.f, (float)img_h, (float)img_h};
        double qx[4], qy[4];  // top_left, top_right, bottom_left, bottom_right as Python ints
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            float ox, oy;
            perspective_point(Hm, cx[k], cy[k], &ox, &oy);
            qx[k] = (double)trunc_like_numpy(ox);
            qy[k] = (double)trunc_like_numpy(oy);
        }
        bool ok = true;
        // top_left  = (find_x_at_y(top_left, bottom_left, 68), 68); top_right likewise
        const double tlx = find_x_at_y(qx[0], qy[0], qx[2], qy[2], (double)kPitchYMax, &ok);
        const double trx = ok ? find_x_at_y(qx[1], qy[1], qx[3], qy[3], (double)kPitchYMax, &ok) : 0.0;
        // bottom_left = (find_x_at_y(bottom_left, NEW top_left, 0), 0); bottom_right likewise (:410-411)
        const double blx = ok ? find_x_at_y(qx[2], qy[2], tlx, (double)kPitchYMax, 0.0, &ok) : 0.0;
        const double brx = ok ? find_x_at_y(qx[3], qy[3], trx, (double)kPitchYMax, 0.0, &ok) : 0.0;
        if (ok) { b[0] = blx; b[1] = tlx; b[2] = trx; b[3] = brx; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) bounds[(size_t)f * 4 + k] = b[k];
}

// Which homography does frame i use?  Reproduces the reference's cadence state machine
// (coordinate_model.py:333,350-367,375-378) over fits that were computed for EVERY frame:
// a fit is attempted on frame i iff i % interval == 0 or the previous attempt failed
// (compute_homography flag); H_use(i) = the latest attempted-and-successful fit at or before i.
// Within a segment [b, b+interval) the attempted frames are b..s where s is the first success, so
//   event(i) = i  if status[i] == OK and no frame in [b(i), i) succeeded, else -1
//   h_index  = running maximum of event.
// One CTA; each thread scans a contiguous slice, slice maxima are combined by a block scan.
constexpr int kSelThreads = 1024;

// `base` = clip index of status[0] (a multiple of interval) -- h_index holds CLIP rows, so a clip can be processed chunk
// by chunk; the row valid before the chunk comes from carry_in or, when carry_ptr is given, from device memory (the
// previous chunk's carry_out: no host round trip between chunks).
__global__ void __launch_bounds__(kSelThreads) select_kernel(const int32_t* __restrict__ status, int F, int interval,
                                                             int carry_in, int32_t* h_index, uint8_t* attempted, int base = 0,
                                                             const int32_t* carry_ptr = nullptr, int32_t* carry_out = nullptr) {
    __shared__ int s_max[kSelThreads];
    const int tid = threadIdx.x;
    const int per = (F + kSelThreads - 1) / kSelThreads;
    const int lo = min(tid * per, F), hi = min(lo + per, F);
    int run = -1;
    for (int i = lo; i < hi; ++i) {
        const int b = (i / interval) * interval;
        bool earlier = false;
        for (int j = b; j < i && !earlier; ++j) earlier = status[j] == EGL_FIT_OK;
        attempted[i] = !earlier;
        const int ev = (!earlier && status[i] == EGL_FIT_OK) ? base + i : -1;
        run = max(run, ev);
        h_index[i] = run;  // local running max, fixed up below
    }
    s_max[tid] = run;
    __syncthreads();
    // exclusive prefix max over slices (Hillis-Steele on 1024 entries)
    int v = s_max[tid];
    for (int d = 1; d < kSelThreads; d <<= 1) {
        const int o = tid >= d ? s_max[tid - d] : -1;
        __syncthreads();
        v = max(v, o);
        s_max[tid] = v;
        __syncthreads();
    }
    if (carry_ptr) carry_in = *carry_ptr;
    const int before = max(tid > 0 ? s_max[tid - 1] : -1, carry_in);
    for (int i = lo; i < hi; ++i) h_index[i] = max(h_index[i], before);
    if (carry_out) {  // every thread has read *carry_ptr before anyone overwrites it (carry_out may alias carry_ptr)
        __syncthreads();
        if (tid == kSelThreads - 1) *carry_out = max(s_max[tid], carry_in);
    }
}

}  // namespace egl

using namespace egl;

extern "C" int egl_select_homography(const int32_t* status, int F, int interval, int carry_in, int32_t* h_index,
                                     uint8_t* attempted, void* stream) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(status && h_index && attempted, EGL_ERR_NULL, "egl_select_homography: null pointer");
    EGL_REQUIRE(F >= 0 && interval >= 1, EGL_ERR_SHAPE, "egl_select_homography: bad arguments");
    if (F == 0) return 0;
    select_kernel<<<1, kSelThreads, 0, (cudaStream_t)stream>>>(status, F, interval, carry_in, h_index, attempted);
    return cuda_status(cudaGetLastError(), "egl_select_homography: kernel launch");
}

extern "C" int egl_select_homography_chunk(const int32_t* status, int F, int interval, int first_frame, const int32_t* carry_in,
                                           int32_t* carry_out, int32_t* h_index, uint8_t* attempted, void* stream) {
    if (F == 0) return 0;
    EGL_REQUIRE(status && h_index && attempted, EGL_ERR_NULL, "egl_select_homography_chunk: null pointer");
    EGL_REQUIRE(F >= 0 && interval >= 1 && first_frame >= 0, EGL_ERR_SHAPE, "egl_select_homography_chunk: bad arguments");
    EGL_REQUIRE(first_frame % interval == 0, EGL_ERR_SHAPE,
                "egl_select_homography_chunk: a chunk must start on a multiple of the homography interval (first_frame=%d interval=%d)",
                first_frame, interval);
    select_kernel<<<1, kSelThreads, 0, (cudaStream_t)stream>>>(status, F, interval, -1, h_index, attempted, first_frame, carry_in,
                                                               carry_out);
    return cuda_status(cudaGetLastError(), "egl_select_homography_chunk: kernel launch");
}

extern "C" int egl_project_points(const double* H, const int32_t* h_index, const float* pts, const int32_t* npts, int F, int P,
                                  int img_w, int img_h, float* out_f, int64_t* out_i, uint8_t* inb, double* bounds,
                                  void* stream) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(H && pts && npts && out_f && out_i && inb && bounds, EGL_ERR_NULL, "egl_project_points: null pointer");
    EGL_REQUIRE(F >= 0 && P >= 0 && img_w > 0 && img_h > 0, EGL_ERR_SHAPE, "egl_project_points: bad shape");
    if (F == 0) return 0;
    const long long total = (long long)F * (P + 1);
    project_kernel<<<(unsigned)((total + 127) / 128), 128, 0, (cudaStream_t)stream>>>(H, h_index, pts, npts, F, P, img_w, img_h,
                                                                                     out_f, out_i, inb, bounds);
    return cuda_status(cudaGetLastError(), "egl_project_points: kernel launch");
}
