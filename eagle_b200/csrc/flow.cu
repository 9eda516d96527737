// F4  keypoint propagation between network frames: gray pyramid, pyramidal Lucas-Kanade tracker,
// the reference's movement / hue filters, dict merge, brightness calibration and the inlier commit.
//
// Replaces, bit for bit, what coordinate_model.py:281,313-330,359-362,419-478,520-555 do through OpenCV:
//   cv2.cvtColor(BGR2GRAY)            15-bit fixed point
//   cv2.calcOpticalFlowPyrLK          pyrDown pyramid (REFLECT_101), Scharr derivatives, 14-bit bilinear
//                                     window weights, float sums in the lane order of OpenCV's SSE build
//   cv2.cvtColor(BGR2HSV)             hue (12-bit fixed point) and value
//   np.linalg.norm / np.mean / np.std float32 with numpy's pairwise summation
// The CPU statement of the same arithmetic, pinned against the live libraries, is oracle/optflow.py.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "flow_core.cuh"  // the scalar statement of the same arithmetic (also compiled for the host by the CPU tests)

namespace egl {

constexpr int kWin = kLkWin;      // 15: lk_params winSize (coordinate_model.py:65)
constexpr int kHalf = kLkHalf;    // (winSize - 1) / 2
constexpr int kMaxLevels = kLkMaxLevels;
constexpr int kWBits = kLkWBits;

// ------------------------------------------------------------------------------------------------
// gray + pyramid
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gray_kernel(const uint8_t* __restrict__ frames, int H, int W, long long row_stride,
                                                   long long frame_stride, uint8_t* __restrict__ pyr, long long pyr_stride, int f0) {
    const int f = blockIdx.z + f0, y = blockIdx.y;
    const int x0 = (blockIdx.x * 256 + threadIdx.x) * 4;
    if (x0 >= W) return;
    const uint8_t* src = frames + (long long)f * frame_stride + (long long)y * row_stride + 3ll * x0;
    uint8_t* dst = pyr + (long long)f * pyr_stride + (long long)y * W + x0;
    if (x0 + 4 <= W && (((uintptr_t)src | (uintptr_t)dst) & 3) == 0) {
        const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
        const uint32_t a = __ldcs(s4), b = __ldcs(s4 + 1), c = __ldcs(s4 + 2);  // B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
        const int g0 = gray_of(a & 255, (a >> 8) & 255, (a >> 16) & 255);
        const int g1 = gray_of(a >> 24, b & 255, (b >> 8) & 255);
        const int g2 = gray_of((b >> 16) & 255, b >> 24, c & 255);
        const int g3 = gray_of((c >> 8) & 255, (c >> 16) & 255, c >> 24);
        *reinterpret_cast<uint32_t*>(dst) = (uint32_t)g0 | ((uint32_t)g1 << 8) | ((uint32_t)g2 << 16) | ((uint32_t)g3 << 24);
    } else {
        for (int k = 0; k < 4 && x0 + k < W; ++k) dst[k] = (uint8_t)gray_of(src[3 * k], src[3 * k + 1], src[3 * k + 2]);
    }
}

// cv2.pyrDown: separable [1 4 6 4 1], exact integer sum, (s + 128) >> 8, REFLECT_101
__global__ void __launch_bounds__(256) pyrdown_kernel(uint8_t* __restrict__ pyr, long long pyr_stride, long long src_off, int sw, int sh,
                                                      long long dst_off, int dw, int dh, int f0) {
    const int f = blockIdx.z + f0;
    const int ox = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    if (ox >= dw || oy >= dh) return;
    const uint8_t* src = pyr + (long long)f * pyr_stride + src_off;
    int xs[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * ox + k - 2, sw);
    int acc = 0;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint8_t* row = src + (long long)reflect101(2 * oy + r - 2, sh) * sw;
        const int s = row[xs[0]] + 4 * row[xs[1]] + 6 * row[xs[2]] + 4 * row[xs[3]] + row[xs[4]];
        acc += (r == 0 || r == 4) ? s : (r == 2 ? 6 * s : 4 * s);
    }
    pyr[(long long)f * pyr_stride + dst_off + (long long)oy * dw + ox] = (uint8_t)((acc + 128) >> 8);
}

// Streaming variants for frames whose rows are 16-byte friendly (W % 16 == 0, aligned strides): the
// general kernels above spend their time on per-byte loads and index arithmetic (1.3 TB/s on 1080p).
//
// gray: one thread turns 16 pixels (three 16-byte loads) into one 16-byte store.
// The integer pipe, not HBM, bounds these kernels once the loads are wide, so the 15-bit weights are split
// into bytes (w = 256*hi + lo) and a pixel costs two byte dot products: B,G,R sit in three bytes of one word.
__device__ __forceinline__ uint32_t gray_px(uint32_t p) {  // p = B | G << 8 | R << 16 | (ignored) << 24
    const uint32_t hi = __dp4a(p, 0x00264b0eu, 64u);       // 14 B + 75 G + 38 R + (1 << 14) / 256
    return __dp4a(p, 0x00462397u, hi << 8) >> 15;          // + 151 B + 35 G + 70 R:  3735, 19235, 9798 in total
}
__device__ __forceinline__ uint32_t gray4(uint32_t a, uint32_t b, uint32_t c) {  // 12 bytes B0 G0 R0 B1 | G1 R1 B2 G2 | R2 B3 G3 R3
    const uint32_t g0 = gray_px(a), g1 = gray_px(__funnelshift_r(a, b, 24));
    const uint32_t g2 = gray_px(__funnelshift_r(b, c, 16)), g3 = gray_px(c >> 8);
    return g0 | (g1 << 8) | (g2 << 16) | (g3 << 24);
}

__global__ void __launch_bounds__(256) gray16_kernel(const uint8_t* __restrict__ frames, int H, int W16, long long row_stride,
                                                     long long frame_stride, uint8_t* __restrict__ pyr, long long pyr_stride, int f0) {
    const int f = blockIdx.y + f0;
    const int items = H * W16;
    const uint8_t* fsrc = frames + (long long)f * frame_stride;
    uint8_t* fdst = pyr + (long long)f * pyr_stride;
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const int item = (blockIdx.x * 2 + u) * 256 + threadIdx.x;
        if (item >= items) return;
        const int y = item / W16, xb = item - y * W16;
        const uint4* s4 = reinterpret_cast<const uint4*>(fsrc + (long long)y * row_stride + 48ll * xb);
        const uint4 a = __ldcs(s4), b = __ldcs(s4 + 1), c = __ldcs(s4 + 2);
        uint4 o;
        o.x = gray4(a.x, a.y, a.z);
        o.y = gray4(a.w, b.x, b.y);
        o.z = gray4(b.z, b.w, c.x);
        o.w = gray4(c.y, c.z, c.w);
        *reinterpret_cast<uint4*>(fdst + ((long long)y * W16 + xb) * 16) = o;
    }
}

// pyrDown, 8 output pixels per thread, source width a multiple of 16.  The 5x5 kernel is evaluated as
// byte dot products (IDP4A): per input row a 4-byte window [1 4 6 4] * (vertical weight) plus the fifth
// tap, two instructions per output and row, no unpacking.  Per input row the thread loads one 16-byte
// word and its two 4-byte neighbours; at the image borders the neighbours are rebuilt by REFLECT_101
// from the word itself, rows are reflected by index, so every thread takes the same path.
__global__ void __launch_bounds__(256) pyrdown8_kernel(uint8_t* __restrict__ pyr, long long pyr_stride, long long src_off, int sw, int sh,
                                                       long long dst_off, int dw, int dh, int f0) {
    const int f = blockIdx.z + f0;
    const int xb = blockIdx.x * 32 + (threadIdx.x & 31), oy = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int ox = xb * 8;
    if (ox >= dw || oy >= dh) return;
    const uint8_t* src = pyr + (long long)f * pyr_stride + src_off;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 128;  // rounding term of (s + 128) >> 8
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint8_t* row = src + (long long)reflect101(2 * oy + r - 2, sh) * sw + 2 * ox;
        const uint4 m = *reinterpret_cast<const uint4*>(row);
        uint32_t W[6];
        W[1] = m.x; W[2] = m.y; W[3] = m.z; W[4] = m.w;
        // columns 2*ox-4 .. 2*ox-1; left border: col -1 -> 1, col -2 -> 2
        W[0] = ox > 0 ? *reinterpret_cast<const uint32_t*>(row - 4) : ((m.x & 0x00ff0000u) | ((m.x & 0x0000ff00u) << 16));
        // columns 2*ox+16 .. ; right border: col sw -> sw-2
        W[5] = 2 * ox + 16 < sw ? *reinterpret_cast<const uint32_t*>(row + 16) : ((m.w >> 16) & 0xffu);
        const uint32_t wr = (r == 0 || r == 4) ? 1u : (r == 2 ? 6u : 4u);
        const uint32_t taps = wr * 0x04060401u;  // bytes [1, 4, 6, 4] * wr (<= 36)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            // output 2j: columns 4j-2 .. 4j+1 (relative to 2*ox) + column 4j+2
            acc[2 * j] = __dp4a(__funnelshift_r(W[j], W[j + 1], 16), taps, acc[2 * j]);
            acc[2 * j] = __dp4a(W[j + 1], wr << 16, acc[2 * j]);
            // output 2j+1: columns 4j .. 4j+3 + column 4j+4
            acc[2 * j + 1] = __dp4a(W[j + 1], taps, acc[2 * j + 1]);
            acc[2 * j + 1] = __dp4a(W[j + 2], wr, acc[2 * j + 1]);
        }
    }
    uint2 o;
    o.x = (acc[0] >> 8) | ((acc[1] >> 8) << 8) | ((acc[2] >> 8) << 16) | ((acc[3] >> 8) << 24);
    o.y = (acc[4] >> 8) | ((acc[5] >> 8) << 8) | ((acc[6] >> 8) << 16) | ((acc[7] >> 8) << 24);
    *reinterpret_cast<uint2*>(pyr + (long long)f * pyr_stride + dst_off + (long long)oy * dw + ox) = o;
}

// gray + pyramid level 1 in one pass: the level-0 gray of a 32 x 256 tile (plus the 5x5 kernel's halo) is
// built in shared memory, written out once, and level 1 is computed from the shared copy, so level 0 is
// never read back from HBM.  Same arithmetic as gray16_kernel / pyrdown8_kernel.  W % 16 == 0.
constexpr int kFuseW = 256, kFuseH = 32;
constexpr int kFuseSmW = kFuseW + 32;  // columns x0-16 .. x0+271 (16-pixel groups)
constexpr int kFuseSmH = kFuseH + 3;   // rows y0-2 .. y0+32

__global__ void __launch_bounds__(256) gray_l1_kernel(const uint8_t* __restrict__ frames, int H, int W, long long row_stride,
                                                      long long frame_stride, uint8_t* __restrict__ pyr, long long pyr_stride,
                                                      long long off1, int w1, int h1, int f0) {
    __shared__ __align__(16) uint8_t sg[kFuseSmH][kFuseSmW];
    const int f = blockIdx.z + f0;
    const int x0 = blockIdx.x * kFuseW, y0 = blockIdx.y * kFuseH;
    const int tid = threadIdx.x;
    const uint8_t* fsrc = frames + (long long)f * frame_stride;
    uint8_t* fdst = pyr + (long long)f * pyr_stride;
    constexpr int kGroups = kFuseSmW / 16;  // 18
    for (int item = tid; item < kFuseSmH * kGroups; item += 256) {
        const int r = item / kGroups, gi = item - r * kGroups;
        const int gx = x0 - 16 + gi * 16;
        if (gx < 0 || gx >= W) continue;
        const int ya = y0 - 2 + r;
        const int yy = reflect101(ya, H);
        const uint4* s4 = reinterpret_cast<const uint4*>(fsrc + (long long)yy * row_stride + 3ll * gx);
        const uint4 a = __ldg(s4), b = __ldg(s4 + 1), c = __ldg(s4 + 2);
        uint4 o;
        o.x = gray4(a.x, a.y, a.z);
        o.y = gray4(a.w, b.x, b.y);
        o.z = gray4(b.z, b.w, c.x);
        o.w = gray4(c.y, c.z, c.w);
        *reinterpret_cast<uint4*>(&sg[r][gi * 16]) = o;
        if (r >= 2 && r < 2 + kFuseH && gi >= 1 && gi <= kFuseW / 16 && ya < H)
            __stcs(reinterpret_cast<uint4*>(fdst + (long long)ya * W + gx), o);
    }
    __syncthreads();
    // REFLECT_101 columns the kernel touches outside the image: -2, -1 and W
    if (x0 == 0 && tid < kFuseSmH) {
        sg[tid][15] = sg[tid][17];
        sg[tid][14] = sg[tid][18];
    }
    if (x0 + kFuseW >= W && tid >= 64 && tid < 64 + kFuseSmH) {
        const int e = W - x0 + 16;
        sg[tid - 64][e] = sg[tid - 64][e - 2];
    }
    __syncthreads();
    const int ry = tid >> 4, xb = tid & 15;
    const int ox = x0 / 2 + xb * 8, oy = y0 / 2 + ry;
    if (ox >= w1 || oy >= h1) return;
    uint32_t acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 128;
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const uint8_t* row = &sg[2 * ry + r][16 + 16 * xb];
        const uint4 m = *reinterpret_cast<const uint4*>(row);
        uint32_t Wd[6];
        Wd[0] = *reinterpret_cast<const uint32_t*>(row - 4);
        Wd[1] = m.x; Wd[2] = m.y; Wd[3] = m.z; Wd[4] = m.w;
        Wd[5] = *reinterpret_cast<const uint32_t*>(row + 16);
        const uint32_t wr = (r == 0 || r == 4) ? 1u : (r == 2 ? 6u : 4u);
        const uint32_t taps = wr * 0x04060401u;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[2 * j] = __dp4a(__funnelshift_r(Wd[j], Wd[j + 1], 16), taps, acc[2 * j]);
            acc[2 * j] = __dp4a(Wd[j + 1], wr << 16, acc[2 * j]);
            acc[2 * j + 1] = __dp4a(Wd[j + 1], taps, acc[2 * j + 1]);
            acc[2 * j + 1] = __dp4a(Wd[j + 2], wr, acc[2 * j + 1]);
        }
    }
    uint2 o;
    o.x = (acc[0] >> 8) | ((acc[1] >> 8) << 8) | ((acc[2] >> 8) << 16) | ((acc[3] >> 8) << 24);
    o.y = (acc[4] >> 8) | ((acc[5] >> 8) << 8) | ((acc[6] >> 8) << 16) | ((acc[7] >> 8) << 24);
    *reinterpret_cast<uint2*>(fdst + off1 + (long long)oy * w1 + ox) = o;
}

// ------------------------------------------------------------------------------------------------
// Lucas-Kanade tracker: one warp per point, all pyramid levels
// ------------------------------------------------------------------------------------------------
constexpr int kTrackWarps = 4;  // blockIdx.x = group of 4 points, blockIdx.y = frame pair: every point of every pair runs concurrently
constexpr int kPatch = kWin + 3;    // 18: gray support of the 16x16 derivative support
constexpr int kSup = kWin + 1;      // 16

// Per-warp staging.  The 15x15 window arrays use a row pitch of 16 so that the eight pixels a lane owns
// (row = lane / 2, columns 0-7 or 8-14) are one 16-byte (int16) or two 16-byte (int32) accesses.
struct alignas(16) TrackShared {
    uint8_t g[kPatch * kPatch + 12];
    int16_t d[2][kSup * kSup];
    int16_t Iw[kWin * kSup], dIx[kWin * kSup], dIy[kWin * kSup];
    int32_t px[kWin * kSup], py[kWin * kSup];
};

struct TrackArgs {
    const uint8_t* pyr;
    long long pyr_stride;
    PyrLayout L;
    const int32_t* kp_xy;
    const uint8_t* kp_order;
    const int32_t* kp_count;
    int n, prev0, next0, fstep;
    int max_count;
    double eps2, min_eig;
    float* out_pts;
    uint8_t* out_status;
};

__device__ __forceinline__ void unpack9(const uint8_t* row8, int* v) {  // 9 bytes from an 8-byte aligned shared address
    const uint32_t* w = reinterpret_cast<const uint32_t*>(row8);
    const uint32_t a = w[0], b = w[1], c = w[2];
    v[0] = a & 255; v[1] = (a >> 8) & 255; v[2] = (a >> 16) & 255; v[3] = a >> 24;
    v[4] = b & 255; v[5] = (b >> 8) & 255; v[6] = (b >> 16) & 255; v[7] = b >> 24;
    v[8] = c & 255;
}

__global__ void __launch_bounds__(kTrackWarps * 32) track_kernel(TrackArgs a) {
    __shared__ TrackShared s_all[kTrackWarps];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.y;
    TrackShared& s = s_all[warp];
    const int npts = min(a.kp_count[2 * p], kMaxPts);
    const uint8_t* Ibase = a.pyr + (long long)(a.prev0 + p * a.fstep) * a.pyr_stride;
    const uint8_t* Jbase = a.pyr + (long long)(a.next0 + p * a.fstep) * a.pyr_stride;
    const float kScale = 1.f / (1 << 20);  // FLT_SCALE
    const int top = a.L.n - 1;
    // window ownership: lanes 0..29 -> row lane/2, columns 0-7 (even lane) or 8-14 (odd lane)
    const int wr = lane >> 1, wc0 = (lane & 1) * 8;
    const int wn = lane < 30 ? ((lane & 1) ? 7 : 8) : 0;
    const int wbase = wr * kSup + wc0;
    const unsigned kExact = 1u << 24, kClamp = 1u << 25;  // 32 lanes x 2^25 cannot overflow

    const int j = blockIdx.x * kTrackWarps + warp;
    if (j >= npts) return;
    const int ch = a.kp_order[(size_t)p * EGL_ORDER_STRIDE + j];
    const float ptx = (float)a.kp_xy[((size_t)p * kLandmarks + ch) * 2], pty = (float)a.kp_xy[((size_t)p * kLandmarks + ch) * 2 + 1];
    float sx = 0.f, sy = 0.f;  // nextPts[ptidx]
    int status = 1;
    for (int level = top; level >= 0; --level) {
        const int cols = a.L.w[level], rows = a.L.h[level];
        const uint8_t* I = Ibase + a.L.off[level];
        const uint8_t* J = Jbase + a.L.off[level];
        const float scale = (float)(1.0 / (1 << level));
        float px = __fmul_rn(ptx, scale), py = __fmul_rn(pty, scale);
        float nx, ny;
        if (level == top) { nx = px; ny = py; } else { nx = __fmul_rn(sx, 2.f); ny = __fmul_rn(sy, 2.f); }
        sx = nx; sy = ny;
        px = __fsub_rn(px, (float)kHalf); py = __fsub_rn(py, (float)kHalf);
        const int ipx = (int)floorf(px), ipy = (int)floorf(py);
        if (ipx < -kWin || ipx >= cols || ipy < -kWin || ipy >= rows) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy), w00, w01, w10, w11);
        __syncwarp();
        for (int i = lane; i < kPatch * kPatch; i += 32) {
            const int r = i / kPatch, c = i - r * kPatch;
            s.g[i] = I[(long long)reflect101(ipy - 1 + r, rows) * cols + reflect101(ipx - 1 + c, cols)];
        }
        __syncwarp();
        // Scharr derivatives on the 16x16 support; zero outside the image (BORDER_CONSTANT)
        for (int i = lane; i < kSup * kSup; i += 32) {
            const int r = i >> 4, c = i & 15;
            const int ay = ipy + r, ax = ipx + c;
            int dx = 0, dy = 0;
            if (ay >= 0 && ay < rows && ax >= 0 && ax < cols) {
                const uint8_t* g = s.g + r * kPatch + c;  // top-left of the 3x3 neighbourhood
                const int g00 = g[0], g01 = g[1], g02 = g[2];
                const int g10 = g[kPatch], g12 = g[kPatch + 2];
                const int g20 = g[2 * kPatch], g21 = g[2 * kPatch + 1], g22 = g[2 * kPatch + 2];
                dx = ((g02 + g22) * 3 + g12 * 10) - ((g00 + g20) * 3 + g10 * 10);
                dy = ((g22 - g02) + (g20 - g00)) * 3 + (g21 - g01) * 10;
            }
            s.d[0][i] = (int16_t)dx;
            s.d[1][i] = (int16_t)dy;
        }
        __syncwarp();
        unsigned l11 = 0, l22 = 0, t12 = 0;
        int l12 = 0;
        for (int i = lane; i < kWin * kSup; i += 32) {
            const int r = i >> 4, c = i & 15;
            if (c == kWin) continue;
            const uint8_t* g = s.g + (r + 1) * kPatch + c + 1;
            s.Iw[i] = (int16_t)((g[0] * w00 + g[1] * w01 + g[kPatch] * w10 + g[kPatch + 1] * w11 + (1 << (kWBits - 6))) >> (kWBits - 5));
            const int16_t* d0 = s.d[0] + i;
            const int16_t* d1 = s.d[1] + i;
            const int ixv = (d0[0] * w00 + d0[1] * w01 + d0[kSup] * w10 + d0[kSup + 1] * w11 + (1 << (kWBits - 1))) >> kWBits;
            const int iyv = (d1[0] * w00 + d1[1] * w01 + d1[kSup] * w10 + d1[kSup + 1] * w11 + (1 << (kWBits - 1))) >> kWBits;
            s.dIx[i] = (int16_t)ixv;
            s.dIy[i] = (int16_t)iyv;
            l11 += (unsigned)(ixv * ixv); l22 += (unsigned)(iyv * iyv); l12 += ixv * iyv; t12 += (unsigned)abs(ixv * iyv);
        }
        __syncwarp();
        // Float sums of integers are exact, whatever the order, while every partial sum stays below 2^24:
        // then OpenCV's lane-ordered accumulation equals the integer sum and one warp reduction gives it.
        const unsigned S11 = __reduce_add_sync(kFull, min(l11, kClamp)), S22 = __reduce_add_sync(kFull, min(l22, kClamp));
        const unsigned T12 = __reduce_add_sync(kFull, min(t12, kClamp));
        const int S12 = __reduce_add_sync(kFull, l12);
        float A11, A12, A22;
        if (S11 < kExact && S22 < kExact && T12 < kExact) {
            A11 = __fmul_rn((float)S11, kScale); A12 = __fmul_rn((float)S12, kScale); A22 = __fmul_rn((float)S22, kScale);
        } else {
            // 15 accumulation chains in OpenCV's order: 3 sums x {4 SIMD lanes (columns 0-7), scalar tail (columns 8-14)}
            float acc = 0.f;
            if (lane < 15) {
                const int q = lane / 5, r = lane - q * 5;
                const int16_t* u = q == 2 ? s.dIy : s.dIx;
                const int16_t* v = q == 0 ? s.dIx : s.dIy;
                for (int y = 0; y < kWin; ++y) {
                    const int o = y * kSup;
                    if (r < 4) {
                        acc = __fadd_rn(acc, (float)((int)u[o + r] * (int)v[o + r]));
                        acc = __fadd_rn(acc, (float)((int)u[o + 4 + r] * (int)v[o + 4 + r]));
                    } else {
#pragma unroll
                        for (int x = 8; x < kWin; ++x) acc = __fadd_rn(acc, (float)((int)u[o + x] * (int)v[o + x]));
                    }
                }
            }
            float A[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                const float l0 = __shfl_sync(kFull, acc, q * 5), l1 = __shfl_sync(kFull, acc, q * 5 + 1);
                const float l2 = __shfl_sync(kFull, acc, q * 5 + 2), l3 = __shfl_sync(kFull, acc, q * 5 + 3);
                const float t = __shfl_sync(kFull, acc, q * 5 + 4);
                A[q] = __fmul_rn(__fadd_rn(t, reduce4(l0, l1, l2, l3)), kScale);
            }
            A11 = A[0]; A12 = A[1]; A22 = A[2];
        }
        float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
        const float dd = __fsub_rn(A11, A22);
        const float min_eig = __fdiv_rn(
            __fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(__fadd_rn(__fmul_rn(dd, dd), __fmul_rn(__fmul_rn(4.f, A12), A12)))),
            (float)(2 * kWin * kWin));
        if ((double)min_eig < a.min_eig || D < 1.1920928955078125e-7f) {
            if (level == 0) status = 0;
            continue;
        }
        D = __fdiv_rn(1.f, D);
        nx = __fsub_rn(nx, (float)kHalf); ny = __fsub_rn(ny, (float)kHalf);
        // this lane's eight window pixels of I, Ix, Iy stay in registers for all iterations of the level
        int Iv[8], Xv[8], Yv[8];
        {
            const uint4 qi = wn ? *reinterpret_cast<const uint4*>(s.Iw + wbase) : make_uint4(0, 0, 0, 0);
            const uint4 qx = wn ? *reinterpret_cast<const uint4*>(s.dIx + wbase) : make_uint4(0, 0, 0, 0);
            const uint4 qy = wn ? *reinterpret_cast<const uint4*>(s.dIy + wbase) : make_uint4(0, 0, 0, 0);
            const uint32_t wi[4] = {qi.x, qi.y, qi.z, qi.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w}, wy[4] = {qy.x, qy.y, qy.z, qy.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                Iv[2 * k] = (int16_t)(wi[k] & 0xffff); Iv[2 * k + 1] = (int16_t)(wi[k] >> 16);
                Xv[2 * k] = (int16_t)(wx[k] & 0xffff); Xv[2 * k + 1] = (int16_t)(wx[k] >> 16);
                Yv[2 * k] = (int16_t)(wy[k] & 0xffff); Yv[2 * k + 1] = (int16_t)(wy[k] >> 16);
            }
            if (wn == 7) { Xv[7] = 0; Yv[7] = 0; }  // column 15 does not exist: contributes nothing
        }
        float pdx = 0.f, pdy = 0.f;
        for (int it = 0; it < a.max_count; ++it) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -kWin || inx >= cols || iny < -kWin || iny >= rows) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny), w00, w01, w10, w11);
            __syncwarp();
            if (inx >= 0 && inx + kSup <= cols && iny >= 0 && iny + kSup <= rows) {
                // 16x16 patch inside the image: each lane fetches half a row
                const uint8_t* src = J + (long long)(iny + wr) * cols + inx + wc0;
                uint32_t lo = 0, hi = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    lo |= (uint32_t)src[k] << (8 * k);
                    hi |= (uint32_t)src[4 + k] << (8 * k);
                }
                *reinterpret_cast<uint2*>(s.g + wbase) = make_uint2(lo, hi);
            } else {
                for (int i = lane; i < kSup * kSup; i += 32) {
                    const int r = i >> 4, c = i & 15;
                    s.g[i] = J[(long long)reflect101(iny + r, rows) * cols + reflect101(inx + c, cols)];
                }
            }
            __syncwarp();
            int sbx = 0, sby = 0;
            unsigned tbx = 0, tby = 0;
            int vx[8], vy[8];
            if (wn) {
                int t[9], b[9];
                unpack9(s.g + wbase, t);
                unpack9(s.g + wbase + kSup, b);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int diff = ((t[k] * w00 + t[k + 1] * w01 + b[k] * w10 + b[k + 1] * w11 + (1 << (kWBits - 6))) >> (kWBits - 5)) - Iv[k];
                    vx[k] = diff * Xv[k];
                    vy[k] = diff * Yv[k];
                    sbx += vx[k]; sby += vy[k]; tbx += (unsigned)abs(vx[k]); tby += (unsigned)abs(vy[k]);
                }
            }
            const unsigned Tx = __reduce_add_sync(kFull, min(tbx, kClamp)), Ty = __reduce_add_sync(kFull, min(tby, kClamp));
            float b1, b2;
            if (Tx < kExact && Ty < kExact) {  // exact regime (see above): the usual case
                b1 = __fmul_rn((float)__reduce_add_sync(kFull, sbx), kScale);
                b2 = __fmul_rn((float)__reduce_add_sync(kFull, sby), kScale);
            } else {
                if (wn) {
#pragma unroll
                    for (int k = 0; k < 8; k += 4) {
                        *reinterpret_cast<int4*>(s.px + wbase + k) = make_int4(vx[k], vx[k + 1], vx[k + 2], vx[k + 3]);
                        *reinterpret_cast<int4*>(s.py + wbase + k) = make_int4(vy[k], vy[k + 1], vy[k + 2], vy[k + 3]);
                    }
                }
                __syncwarp();
                // mismatch vector in OpenCV's order: lanes 0-3 = qb0, 4-7 = qb1 (pair sums k, k+4 as int, then float), 8/9 = scalar tails
                float bacc = 0.f;
                if (lane < 8) {
                    const int32_t* P = (lane & 1) ? s.py : s.px;
                    const int k = ((lane >> 2) << 1) + ((lane >> 1) & 1);  // qb0: pixels 0,1; qb1: pixels 2,3
                    for (int y = 0; y < kWin; ++y) bacc = __fadd_rn(bacc, (float)(P[y * kSup + k] + P[y * kSup + k + 4]));
                } else if (lane < 10) {
                    const int32_t* P = (lane & 1) ? s.py : s.px;
                    for (int y = 0; y < kWin; ++y)
#pragma unroll
                        for (int x = 8; x < kWin; ++x) bacc = __fadd_rn(bacc, (float)P[y * kSup + x]);
                }
                // q = qb0 + qb1 = (X0, Y0, X1, Y1); ib1 = t1 + (X0 + X1), ib2 = t2 + (Y0 + Y1)
                const float q0 = __fadd_rn(__shfl_sync(kFull, bacc, 0), __shfl_sync(kFull, bacc, 4));
                const float q1 = __fadd_rn(__shfl_sync(kFull, bacc, 1), __shfl_sync(kFull, bacc, 5));
                const float q2 = __fadd_rn(__shfl_sync(kFull, bacc, 2), __shfl_sync(kFull, bacc, 6));
                const float q3 = __fadd_rn(__shfl_sync(kFull, bacc, 3), __shfl_sync(kFull, bacc, 7));
                b1 = __fmul_rn(__fadd_rn(__shfl_sync(kFull, bacc, 8), __fadd_rn(q0, q2)), kScale);
                b2 = __fmul_rn(__fadd_rn(__shfl_sync(kFull, bacc, 9), __fadd_rn(q1, q3)), kScale);
            }
            const float dx = __fmul_rn(__fsub_rn(__fmul_rn(A12, b2), __fmul_rn(A22, b1)), D);
            const float dy = __fmul_rn(__fsub_rn(__fmul_rn(A12, b1), __fmul_rn(A11, b2)), D);
            nx = __fadd_rn(nx, dx); ny = __fadd_rn(ny, dy);
            sx = __fadd_rn(nx, (float)kHalf); sy = __fadd_rn(ny, (float)kHalf);
            if (__dadd_rn(__dmul_rn((double)dx, (double)dx), __dmul_rn((double)dy, (double)dy)) <= a.eps2) break;
            if (it > 0 && fabs((double)__fadd_rn(dx, pdx)) < 0.01 && fabs((double)__fadd_rn(dy, pdy)) < 0.01) {
                sx = __fsub_rn(sx, __fmul_rn(dx, 0.5f));
                sy = __fsub_rn(sy, __fmul_rn(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {
            // the error pass re-checks the final window position
            const int fx = (int)floorf(__fsub_rn(sx, (float)kHalf)), fy = (int)floorf(__fsub_rn(sy, (float)kHalf));
            if (fx < -kWin || fx >= cols || fy < -kWin || fy >= rows) status = 0;
        }
    }
    if (lane == 0) {  // results are stored by channel, so that a consumer may look at a subset of the tracked set
        a.out_pts[((size_t)p * kMaxPts + ch) * 2] = sx;
        a.out_pts[((size_t)p * kMaxPts + ch) * 2 + 1] = sy;
        a.out_status[(size_t)p * kMaxPts + ch] = (uint8_t)status;
    }
}

// The same tracker as one thread per point: lk_track_point of flow_core.cuh, the scalar statement that the
// CPU suite compiles for the host and holds against live cv2.  EGL_TRACK_VARIANT=1 selects it (cross-check
// of the warp kernel above; ~20x slower).
#ifdef EGL_BENCH_VARIANTS
__global__ void __launch_bounds__(64) track_thread_kernel(TrackArgs a) {
    const int p = blockIdx.x, j = threadIdx.x;
    if (j >= min(a.kp_count[2 * p], kMaxPts)) return;
    const int ch = a.kp_order[(size_t)p * EGL_ORDER_STRIDE + j];
    const uint8_t* I = a.pyr + (long long)(a.prev0 + p * a.fstep) * a.pyr_stride;
    const uint8_t* J = a.pyr + (long long)(a.next0 + p * a.fstep) * a.pyr_stride;
    float out[2];
    const int status = lk_track_point(I, J, a.L, (float)a.kp_xy[((size_t)p * kLandmarks + ch) * 2],
                                      (float)a.kp_xy[((size_t)p * kLandmarks + ch) * 2 + 1], a.max_count, a.eps2, a.min_eig, out);
    a.out_pts[((size_t)p * kMaxPts + ch) * 2] = out[0];
    a.out_pts[((size_t)p * kMaxPts + ch) * 2 + 1] = out[1];
    a.out_status[(size_t)p * kMaxPts + ch] = (uint8_t)status;
}
#endif  // EGL_BENCH_VARIANTS

// ------------------------------------------------------------------------------------------------
// the reference's filters on the tracked points (coordinate_model.py:438-478): one warp per frame pair
// ------------------------------------------------------------------------------------------------
constexpr int kFilterWarps = 4;

struct FilterArgs {
    const uint8_t* frames;
    int H, W;
    long long row_stride, frame_stride;
    int hue0, fstep;
    const int32_t* prev_xy;
    const uint8_t* prev_order;
    const int32_t* prev_count;
    const float* new_pts;
    const uint8_t* status;
    int n;
    int32_t* out_xy;
    uint8_t* out_order;
    int32_t* out_count;
    uint8_t* out_src;
};

// mean hue (cv2 BGR2HSV, 0..179) of the clipped 3x3 block around (x, y), as np.mean gives it (float64)
__device__ double mean_hue_3x3(const uint8_t* frame, int H, int W, long long row_stride, long long xi, long long yi) {
    const int x = (int)max(0ll, min(xi, (long long)W - 1)), y = (int)max(0ll, min(yi, (long long)H - 1));
    const int x0 = max(0, x - 1), x1 = min(W, x + 2), y0 = max(0, y - 1), y1 = min(H, y + 2);
    int sum = 0;
    for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) {
            const uint8_t* px = frame + (long long)yy * row_stride + 3ll * xx;
            const int h = hue_of(px[0], px[1], px[2]);
            sum += h;
        }
    return (double)sum / (double)((y1 - y0) * (x1 - x0));
}

__global__ void __launch_bounds__(kFilterWarps * 32) filter_kernel(FilterArgs a) {
    __shared__ float s_move[kFilterWarps][kMaxPts];
    __shared__ float s_dev[kFilterWarps][kMaxPts];
    __shared__ uint8_t s_keep[kFilterWarps][kMaxPts];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int p = blockIdx.x * kFilterWarps + warp;
    if (p >= a.n) return;
    const int n0 = min(a.prev_count[2 * p], kMaxPts);
    const uint8_t* order = a.prev_order + (size_t)p * EGL_ORDER_STRIDE;
    const int32_t* pxy = a.prev_xy + (size_t)p * kLandmarks * 2;
    const float* npt = a.new_pts + (size_t)p * kMaxPts * 2;
    // every entry of this frame starts untagged; the ones emitted below become numpy-int tuples
    for (int i = lane; i < EGL_ORDER_STRIDE; i += 32) a.out_src[(size_t)p * EGL_ORDER_STRIDE + i] = 0;
    // status == 1 compaction (:438-439)
    int nk = 0;
    for (int base = 0; base < n0; base += 32) {
        const int j = base + lane;
        const bool st = j < n0 && a.status[(size_t)p * kMaxPts + order[j]] == 1;
        const unsigned m = __ballot_sync(kFull, st);
        if (st) s_keep[warp][nk + __popc(m & ((1u << lane) - 1))] = (uint8_t)j;
        nk += __popc(m);
    }
    __syncwarp();
    if (nk == 0) {
        if (lane == 0) { a.out_count[2 * p] = 0; a.out_count[2 * p + 1] = 0; }
        return;
    }
    for (int jj = lane; jj < nk; jj += 32) {
        const int j = s_keep[warp][jj];
        const int ch = order[j];
        const float dx = __fsub_rn(npt[2 * ch], (float)pxy[2 * ch]), dy = __fsub_rn(npt[2 * ch + 1], (float)pxy[2 * ch + 1]);
        s_move[warp][jj] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
    }
    __syncwarp();
    float mean = 0.f, stdv = 0.f;
    if (lane == 0) {
        const float fn = (float)nk;
        mean = __fdiv_rn(pairwise_sum_f32(s_move[warp], nk), fn);
        for (int i = 0; i < nk; ++i) {
            const float d = __fsub_rn(s_move[warp][i], mean);
            s_dev[warp][i] = __fmul_rn(d, d);
        }
        stdv = __fadd_rn(__fsqrt_rn(__fdiv_rn(pairwise_sum_f32(s_dev[warp], nk), fn)), 1e-6f);
    }
    mean = __shfl_sync(kFull, mean, 0);
    stdv = __shfl_sync(kFull, stdv, 0);
    const uint8_t* frame = a.frames + (long long)(a.hue0 + p * a.fstep) * a.frame_stride;
    int nout = 0;
    for (int base = 0; base < nk; base += 32) {
        const int jj = base + lane;
        bool pass = false;
        long long nxi = 0, nyi = 0;
        if (jj < nk) {
            const int j = s_keep[warp][jj];
            const int ch = order[j];
            const float z = __fdiv_rn(__fsub_rn(s_move[warp][jj], mean), stdv);
            if (!(z > 2.f)) {
                nxi = (long long)npt[2 * ch]; nyi = (long long)npt[2 * ch + 1];  // astype(int): truncation
                const double hc = mean_hue_3x3(frame, a.H, a.W, a.row_stride, nxi, nyi);
                const double hp = mean_hue_3x3(frame, a.H, a.W, a.row_stride, (long long)pxy[2 * ch], (long long)pxy[2 * ch + 1]);
                pass = !(fabs(hc - hp) > 25.0);
            }
        }
        const unsigned m = __ballot_sync(kFull, pass);
        if (pass) {
            // :446 -- the label is taken at the position in the status-filtered arrays
            const int label = order[jj];
            const int pos = nout + __popc(m & ((1u << lane) - 1));
            a.out_order[(size_t)p * EGL_ORDER_STRIDE + pos] = (uint8_t)label;
            a.out_xy[((size_t)p * kLandmarks + label) * 2] = (int32_t)nxi;
            a.out_xy[((size_t)p * kLandmarks + label) * 2 + 1] = (int32_t)nyi;
            a.out_src[(size_t)p * EGL_ORDER_STRIDE + label] = EGL_KP_NUMPY_INT;
        }
        nout += __popc(m);
    }
    if (lane == 0) { a.out_count[2 * p] = nout; a.out_count[2 * p + 1] = 0; }
}

// ------------------------------------------------------------------------------------------------
// dict merge {**a, **b}, brightness calibration, inlier commit
// ------------------------------------------------------------------------------------------------
__global__ void merge_kernel(int32_t* a_xy, uint8_t* a_order, int32_t* a_count, uint8_t* a_src, const int32_t* b_xy,
                             const uint8_t* b_order, const int32_t* b_count, const uint8_t* b_src, const uint8_t* apply, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n || (apply && !apply[f])) return;
    int na = min(a_count[2 * f], kMaxPts);
    const int nb = min(b_count[2 * f], kMaxPts);
    uint64_t have = 0;
    for (int k = 0; k < na; ++k) have |= 1ull << a_order[(size_t)f * EGL_ORDER_STRIDE + k];
    for (int k = 0; k < nb; ++k) {
        const int ch = b_order[(size_t)f * EGL_ORDER_STRIDE + k];
        a_xy[((size_t)f * kLandmarks + ch) * 2] = b_xy[((size_t)f * kLandmarks + ch) * 2];
        a_xy[((size_t)f * kLandmarks + ch) * 2 + 1] = b_xy[((size_t)f * kLandmarks + ch) * 2 + 1];
        a_src[(size_t)f * EGL_ORDER_STRIDE + ch] = b_src ? b_src[(size_t)f * EGL_ORDER_STRIDE + ch] : (uint8_t)EGL_KP_PY_INT;
        if (!((have >> ch) & 1)) {
            have |= 1ull << ch;
            a_order[(size_t)f * EGL_ORDER_STRIDE + na++] = (uint8_t)ch;
        }
    }
    a_count[2 * f] = na;
}

constexpr int kCalWarps = 4;

__global__ void __launch_bounds__(kCalWarps * 32) calibrate_kernel(const uint8_t* frames, int H, int W, long long row_stride,
                                                                   long long frame_stride, int frame0, int fstep, int32_t* kp_xy,
                                                                   const uint8_t* kp_order, const int32_t* kp_count, uint8_t* kp_src,
                                                                   int32_t* err, int n) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int f = blockIdx.x * kCalWarps + warp;
    if (f >= n) return;
    const uint8_t* frame = frames + (long long)(frame0 + f * fstep) * frame_stride;
    const int cnt = min(kp_count[2 * f], kMaxPts);
    bool bad = false;
    for (int k = lane; k < cnt; k += 32) {
        const int ch = kp_order[(size_t)f * EGL_ORDER_STRIDE + k];
        int32_t* xy = kp_xy + ((size_t)f * kLandmarks + ch) * 2;
        const int x = xy[0], y = xy[1];
        if (!(x >= 0 && x < W && y >= 0 && y < H)) continue;  // :535-537
        const uint8_t* c = frame + (long long)y * row_stride + 3ll * x;
        if (max(max(c[0], c[1]), c[2]) >= 150) continue;       // :538-541
        const int x0 = max(0, x - 3), x1 = min(W, x + 3), y0 = max(0, y - 3), y1 = min(H, y + 3);
        if (y1 - y0 <= 3 || x1 - x0 <= 3) {  // grid_hsv[3, 3] would be out of range (:548): IndexError in the reference
            bad = true;
            continue;
        }
        int best = -1, bx = 0, by = 0;
        for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx) {
                const uint8_t* q = frame + (long long)yy * row_stride + 3ll * xx;
                const int v = max(max(q[0], q[1]), q[2]);
                if (v > best) { best = v; bx = xx - x0; by = yy - y0; }  // np.argmax: first maximum in row-major order
            }
        xy[0] = max(0, min(x + bx - 3, W - 1));  // :551-552
        xy[1] = max(0, min(y + by - 3, H - 1));
        kp_src[(size_t)f * EGL_ORDER_STRIDE + ch] = EGL_KP_NUMPY_INT;
    }
    if (__any_sync(kFull, bad) && lane == 0) err[f] = 1;
}

__global__ void commit_kernel(uint8_t* kp_order, int32_t* kp_count, uint8_t* kp_src, const int32_t* status, const uint64_t* inlier_mask,
                              const uint8_t* sched, uint8_t* retry, uint8_t* fit_ok, int n) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= n) return;
    const bool attempted = (sched ? sched[f] : 1) | (retry ? retry[f] : 0);
    uint8_t ok = 0;
    if (attempted) {
        if (status[f] == EGL_FIT_OK) {
            // keypoints := inliers, as float lists (coordinate_model.py:359-362)
            const uint64_t inl = inlier_mask[f];
            const int cnt = min(kp_count[2 * f], kMaxPts);
            int m = 0;
            for (int k = 0; k < cnt; ++k) {
                const int ch = kp_order[(size_t)f * EGL_ORDER_STRIDE + k];
                if ((inl >> ch) & 1) {
                    kp_order[(size_t)f * EGL_ORDER_STRIDE + m++] = (uint8_t)ch;
                    kp_src[(size_t)f * EGL_ORDER_STRIDE + ch] = EGL_KP_FLOAT;
                }
            }
            kp_count[2 * f] = m;
            if (retry) retry[f] = 0;
            ok = 1;
        } else if (retry) {
            retry[f] = 1;  // :350-352, :366-367
        }
    }
    if (fit_ok) fit_ok[f] = ok;
}

}  // namespace egl

using namespace egl;

extern "C" int64_t egl_pyramid_bytes(int H, int W, int max_level) {
    if (H <= 0 || W <= 0 || max_level < 0 || max_level >= kMaxLevels) return -1;
    return pyramid_layout(H, W, max_level).bytes;
}

extern "C" int egl_gray_pyramid(const uint8_t* frames, int F, int H, int W, size_t row_stride, size_t frame_stride, int max_level,
                                uint8_t* pyr, void* stream) {
    return egl_gray_pyramid_strided(frames, F, H, W, row_stride, frame_stride, max_level, pyr, 0, stream);
}

extern "C" int egl_gray_pyramid_strided(const uint8_t* frames, int F, int H, int W, size_t row_stride, size_t frame_stride,
                                        int max_level, uint8_t* pyr, size_t pyr_stride, void* stream) {
    if (F == 0) return 0;
    EGL_REQUIRE(frames && pyr, EGL_ERR_NULL, "egl_gray_pyramid: null pointer");
    EGL_REQUIRE(F > 0 && H > 0 && W > 0 && row_stride >= (size_t)3 * W && max_level >= 0 && max_level < kMaxLevels, EGL_ERR_SHAPE,
                "egl_gray_pyramid: bad shape (F=%d H=%d W=%d max_level=%d)", F, H, W, max_level);
    PyrLayout L = pyramid_layout(H, W, max_level);
    EGL_REQUIRE(pyr_stride == 0 || pyr_stride >= (size_t)L.bytes, EGL_ERR_SHAPE, "egl_gray_pyramid_strided: pyr_stride smaller than one pyramid");
    if (pyr_stride) L.bytes = (long long)pyr_stride;  // the kernels only use it as the distance between two frames' pyramids
    cudaStream_t s = (cudaStream_t)stream;
    bool general_only = false, unfused = false;
#ifdef EGL_BENCH_VARIANTS
    static const char* env = getenv("EGL_PYRAMID_VARIANT");  // measurement switch: 1 = general per-byte kernels only, 2 = unfused fast kernels
    general_only = env && atoi(env) == 1;
    unfused = env && atoi(env) == 2;
#endif
    const bool fast_gray = !general_only && W % 16 == 0 && row_stride % 16 == 0 && frame_stride % 16 == 0 &&
                           ((uintptr_t)frames & 15) == 0 && ((uintptr_t)pyr & 15) == 0 && L.bytes % 16 == 0;
    // One pass over all frames per level: splitting the clip into L2-sized groups (so that pyrDown would find
    // the gray level still cached) was measured slower at every group size -- 4.8 ms whole, 5.6 ms in groups of
    // 32, 8.6 ms in groups of 8 for 2250 frames at 1080p -- the launch tails cost more than the re-read.
    for (int f0 = 0; f0 < F; f0 += 32768) {
        const int nf = min(32768, F - f0);
        int first_level = 1;
        if (fast_gray && !unfused && L.n > 1 && H >= 3) {
            gray_l1_kernel<<<dim3((W + kFuseW - 1) / kFuseW, (H + kFuseH - 1) / kFuseH, nf), 256, 0, s>>>(
                frames, H, W, (long long)row_stride, (long long)frame_stride, pyr, L.bytes, L.off[1], L.w[1], L.h[1], f0);
            first_level = 2;
        } else if (fast_gray) {
            const int items = H * (W / 16);
            gray16_kernel<<<dim3((items + 511) / 512, nf), 256, 0, s>>>(frames, H, W / 16, (long long)row_stride, (long long)frame_stride,
                                                                        pyr, L.bytes, f0);
        } else {
            gray_kernel<<<dim3((W + 1023) / 1024, H, nf), 256, 0, s>>>(frames, H, W, (long long)row_stride, (long long)frame_stride, pyr,
                                                                       L.bytes, f0);
        }
        for (int l = first_level; l < L.n; ++l) {
            if (!general_only && L.w[l - 1] % 16 == 0 && ((uintptr_t)pyr & 15) == 0)
                pyrdown8_kernel<<<dim3((L.w[l] + 255) / 256, (L.h[l] + 7) / 8, nf), 256, 0, s>>>(pyr, L.bytes, L.off[l - 1], L.w[l - 1],
                                                                                             L.h[l - 1], L.off[l], L.w[l], L.h[l], f0);
            else
                pyrdown_kernel<<<dim3((L.w[l] + 31) / 32, (L.h[l] + 7) / 8, nf), 256, 0, s>>>(pyr, L.bytes, L.off[l - 1], L.w[l - 1],
                                                                                            L.h[l - 1], L.off[l], L.w[l], L.h[l], f0);
        }
    }
    return cuda_status(cudaGetLastError(), "egl_gray_pyramid: kernel launch");
}

extern "C" int egl_track_keypoints(const uint8_t* pyr, int H, int W, int max_level, const int32_t* kp_xy, const uint8_t* kp_order,
                                   const int32_t* kp_count, int n, int prev0, int next0, int frame_step, int max_count, double eps,
                                   float* new_pts, uint8_t* status, void* stream) {
    if (n == 0) return 0;
    EGL_REQUIRE(pyr && kp_xy && kp_order && kp_count && new_pts && status, EGL_ERR_NULL, "egl_track_keypoints: null pointer");
    EGL_REQUIRE(n > 0 && H > kWin && W > kWin && max_level >= 0 && max_level < kMaxLevels && prev0 >= 0 && next0 >= 0, EGL_ERR_SHAPE,
                "egl_track_keypoints: bad arguments (n=%d H=%d W=%d max_level=%d)", n, H, W, max_level);
    TrackArgs a;
    a.pyr = pyr;
    a.L = pyramid_layout(H, W, max_level);
    a.pyr_stride = a.L.bytes;
    a.kp_xy = kp_xy; a.kp_order = kp_order; a.kp_count = kp_count;
    a.n = n; a.prev0 = prev0; a.next0 = next0; a.fstep = frame_step;
    a.max_count = min(max(max_count, 0), 100);
    const double e = fmin(fmax(eps, 0.0), 10.0);
    a.eps2 = e * e;
    a.min_eig = 1e-4;
    a.out_pts = new_pts; a.out_status = status;
#ifdef EGL_BENCH_VARIANTS
    static const char* env = getenv("EGL_TRACK_VARIANT");  // 1 = one thread per point (the host-checkable scalar code)
    if (env && atoi(env) == 1) {
        track_thread_kernel<<<n, 64, 0, (cudaStream_t)stream>>>(a);
        return cuda_status(cudaGetLastError(), "egl_track_keypoints: kernel launch");
    }
#endif
    track_kernel<<<dim3(kMaxPts / kTrackWarps, n), kTrackWarps * 32, 0, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "egl_track_keypoints: kernel launch");
}

extern "C" int egl_filter_flow(const uint8_t* frames, int H, int W, size_t row_stride, size_t frame_stride, int hue0, int frame_step,
                               const int32_t* prev_xy, const uint8_t* prev_order, const int32_t* prev_count, const float* new_pts,
                               const uint8_t* status, int n, int32_t* out_xy, uint8_t* out_order, int32_t* out_count, uint8_t* out_src,
                               void* stream) {
    if (n == 0) return 0;
    EGL_REQUIRE(frames && prev_xy && prev_order && prev_count && new_pts && status && out_xy && out_order && out_count && out_src,
                EGL_ERR_NULL, "egl_filter_flow: null pointer");
    EGL_REQUIRE(n > 0 && H > 0 && W > 0 && row_stride >= (size_t)3 * W, EGL_ERR_SHAPE, "egl_filter_flow: bad arguments");
    FilterArgs a{frames, H, W, (long long)row_stride, (long long)frame_stride, hue0, frame_step, prev_xy, prev_order, prev_count,
                 new_pts, status, n, out_xy, out_order, out_count, out_src};
    filter_kernel<<<(n + kFilterWarps - 1) / kFilterWarps, kFilterWarps * 32, 0, (cudaStream_t)stream>>>(a);
    return cuda_status(cudaGetLastError(), "egl_filter_flow: kernel launch");
}

extern "C" int egl_merge_keypoints(int32_t* a_xy, uint8_t* a_order, int32_t* a_count, uint8_t* a_src, const int32_t* b_xy,
                                   const uint8_t* b_order, const int32_t* b_count, const uint8_t* b_src, const uint8_t* apply, int n,
                                   void* stream) {
    if (n == 0) return 0;
    EGL_REQUIRE(a_xy && a_order && a_count && a_src && b_xy && b_order && b_count, EGL_ERR_NULL, "egl_merge_keypoints: null pointer");
    EGL_REQUIRE(n > 0, EGL_ERR_SHAPE, "egl_merge_keypoints: n < 0");
    merge_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(a_xy, a_order, a_count, a_src, b_xy, b_order, b_count, b_src, apply, n);
    return cuda_status(cudaGetLastError(), "egl_merge_keypoints: kernel launch");
}

extern "C" int egl_calibrate_keypoints(const uint8_t* frames, int H, int W, size_t row_stride, size_t frame_stride, int frame0,
                                       int frame_step, int32_t* kp_xy, const uint8_t* kp_order, const int32_t* kp_count,
                                       uint8_t* kp_src, int32_t* err, int n, void* stream) {
    if (n == 0) return 0;
    EGL_REQUIRE(frames && kp_xy && kp_order && kp_count && kp_src && err, EGL_ERR_NULL, "egl_calibrate_keypoints: null pointer");
    EGL_REQUIRE(n > 0 && H > 0 && W > 0 && row_stride >= (size_t)3 * W, EGL_ERR_SHAPE, "egl_calibrate_keypoints: bad arguments");
    calibrate_kernel<<<(n + kCalWarps - 1) / kCalWarps, kCalWarps * 32, 0, (cudaStream_t)stream>>>(
        frames, H, W, (long long)row_stride, (long long)frame_stride, frame0, frame_step, kp_xy, kp_order, kp_count, kp_src, err, n);
    return cuda_status(cudaGetLastError(), "egl_calibrate_keypoints: kernel launch");
}

extern "C" int egl_commit_fit(uint8_t* kp_order, int32_t* kp_count, uint8_t* kp_src, const int32_t* status, const uint64_t* inlier_mask,
                              const uint8_t* sched, uint8_t* retry, uint8_t* fit_ok, int n, void* stream) {
    if (n == 0) return 0;
    EGL_REQUIRE(kp_order && kp_count && kp_src && status && inlier_mask, EGL_ERR_NULL, "egl_commit_fit: null pointer");
    EGL_REQUIRE(n > 0, EGL_ERR_SHAPE, "egl_commit_fit: n < 0");
    commit_kernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(kp_order, kp_count, kp_src, status, inlier_mask, sched, retry, fit_ok, n);
    return cuda_status(cudaGetLastError(), "egl_commit_fit: kernel launch");
}
