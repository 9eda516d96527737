// Scalar statement of the keypoint-propagation arithmetic (colour conversions, pyrDown, Scharr, the
// pyramidal Lucas-Kanade tracker, numpy's pairwise float sum), shared by the kernels of flow.cu and by the
// host-check build.
//
// Like geometry_core.cuh, everything here is a pure __host__ __device__ function with a fixed operation
// order: flow.cu uses the small pieces inside its warp-cooperative kernels and runs lk_track_point as the
// one-thread-per-point variant of the tracker (EGL_TRACK_VARIANT=1), and tests/native/host_check.cpp
// compiles the very same code with g++ (-ffp-contract=off) so that the CPU suite can hold it against live
// cv2.calcOpticalFlowPyrLK without a GPU.  The host build is a test artefact; the product never calls it.
//
// OpenCV references: imgproc color_yuv/color_hsv (cvtColor), pyramids.cpp (pyrDown), video/lkpyramid.cpp
// (calcScharrDeriv, LKTrackerInvoker) of opencv-python 4.11/4.13 -- the library behind
// eagle/models/coordinate_model.py:281,435,461.
#pragma once

#include "geometry_core.cuh"

namespace egl {

constexpr int kLkWin = 15;        // lk_params winSize (coordinate_model.py:65)
constexpr int kLkHalf = 7;        // (winSize - 1) / 2
constexpr int kLkSup = kLkWin + 1;
constexpr int kLkMaxLevels = 4;   // maxLevel <= 3
constexpr int kLkWBits = 14;

struct PyrLayout {
    int n;
    int w[kLkMaxLevels], h[kLkMaxLevels];
    long long off[kLkMaxLevels];
    long long bytes;
};

// buildOpticalFlowPyramid: halve until a level would not be larger than the window; levels 16-byte aligned
static inline PyrLayout pyramid_layout(int H, int W, int max_level) {
    PyrLayout L{};
    int w = W, h = H;
    long long off = 0;
    for (int l = 0; l <= max_level && l < kLkMaxLevels; ++l) {
        if (l > 0) {
            w = (w + 1) / 2;
            h = (h + 1) / 2;
            if (w <= kLkWin || h <= kLkWin) break;
        }
        L.w[l] = w;
        L.h[l] = h;
        L.off[l] = off;
        off += ((long long)w * h + 15) / 16 * 16;
        L.n++;
    }
    L.bytes = off;
    return L;
}

EGL_HD int reflect101(int i, int n) {  // BORDER_REFLECT_101
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = i < 0 ? -i : 2 * (n - 1) - i;
    return i;
}

// cvtColor(BGR2GRAY), uint8: 15-bit fixed point, round half up
EGL_HD int gray_of(int b, int g, int r) { return (b * 3735 + g * 19235 + r * 9798 + (1 << 14)) >> 15; }

EGL_HD int round_half_even_to_int(double v) {
#if defined(__CUDA_ARCH__)
    return __double2int_rn(v);
#else
    return (int)lrint(v);
#endif
}
EGL_HD int round_half_even_to_int(float v) {
#if defined(__CUDA_ARCH__)
    return __float2int_rn(v);
#else
    return (int)lrintf(v);
#endif
}
EGL_HD float fsqrt(float v) {
#if defined(__CUDA_ARCH__)
    return __fsqrt_rn(v);
#else
    return sqrtf(v);
#endif
}

// H of cvtColor(BGR2HSV), uint8, range 0..179: 12-bit fixed point with hdiv_table180[diff] = round((180 << 12) / (6 diff))
EGL_HD int hue_of(int b, int g, int r) {
    const int v = b > g ? (b > r ? b : r) : (g > r ? g : r);
    const int vmin = b < g ? (b < r ? b : r) : (g < r ? g : r);
    const int diff = v - vmin;
    int h = v == r ? g - b : (v == g ? b - r + 2 * diff : r - g + 4 * diff);
    const int hdiv = diff ? round_half_even_to_int((double)(180 << 12) / (6.0 * diff)) : 0;
    h = (h * hdiv + (1 << 11)) >> 12;
    return h < 0 ? h + 180 : h;
}

// one pixel of cv2.pyrDown (uint8): separable [1 4 6 4 1], exact integer sum, (s + 128) >> 8
EGL_HD int pyrdown_pixel(const uint8_t* src, int sw, int sh, int ox, int oy) {
    int xs[5];
    for (int k = 0; k < 5; ++k) xs[k] = reflect101(2 * ox + k - 2, sw);
    int acc = 0;
    for (int r = 0; r < 5; ++r) {
        const uint8_t* row = src + (long long)reflect101(2 * oy + r - 2, sh) * sw;
        const int s = row[xs[0]] + 4 * row[xs[1]] + 6 * row[xs[2]] + 4 * row[xs[3]] + row[xs[4]];
        acc += (r == 0 || r == 4) ? s : (r == 2 ? 6 * s : 4 * s);
    }
    return (acc + 128) >> 8;
}

// numpy's pairwise float32 sum for n <= 128 (sequential for n < 8, else 8 interleaved accumulators)
EGL_HD_NOINLINE float pairwise_sum_f32(const float* v, int n) {
    if (n < 8) {
        float s = n > 0 ? v[0] : 0.f;
        for (int i = 1; i < n; ++i) s = fadd(s, v[i]);
        return s;
    }
    float r[8];
    for (int k = 0; k < 8; ++k) r[k] = v[k];
    int i = 8;
    for (; i + 8 <= n; i += 8)
        for (int k = 0; k < 8; ++k) r[k] = fadd(r[k], v[i + k]);
    float s = fadd(fadd(fadd(r[0], r[1]), fadd(r[2], r[3])), fadd(fadd(r[4], r[5]), fadd(r[6], r[7])));
    for (; i < n; ++i) s = fadd(s, v[i]);
    return s;
}

// LKTrackerInvoker: integer bilinear weights, iw11 takes the remainder
EGL_HD void lk_weights(float a, float b, int& w00, int& w01, int& w10, int& w11) {
    const float s = (float)(1 << kLkWBits);
    const float na = fsub(1.f, a), nb = fsub(1.f, b);
    w00 = round_half_even_to_int(fmul(fmul(na, nb), s));
    w01 = round_half_even_to_int(fmul(fmul(a, nb), s));
    w10 = round_half_even_to_int(fmul(fmul(na, b), s));
    w11 = (1 << kLkWBits) - w00 - w01 - w10;
}

// (l0 + l2) + (l1 + l3): the movehl/shuffle reduction of a 4-lane float register
EGL_HD float reduce4(float l0, float l1, float l2, float l3) { return fadd(fadd(l0, l2), fadd(l1, l3)); }

EGL_HD int gray_at(const uint8_t* img, int cols, int rows, int y, int x) {  // level image with its REFLECT_101 border
    return img[(long long)reflect101(y, rows) * cols + reflect101(x, cols)];
}

// calcScharrDeriv at (y, x): 3/10/3 smoothing across, central difference along; zero outside the image
EGL_HD void scharr_at(const uint8_t* img, int cols, int rows, int y, int x, int& dx, int& dy) {
    dx = dy = 0;
    if (y < 0 || y >= rows || x < 0 || x >= cols) return;
    const int g00 = gray_at(img, cols, rows, y - 1, x - 1), g01 = gray_at(img, cols, rows, y - 1, x), g02 = gray_at(img, cols, rows, y - 1, x + 1);
    const int g10 = gray_at(img, cols, rows, y, x - 1), g12 = gray_at(img, cols, rows, y, x + 1);
    const int g20 = gray_at(img, cols, rows, y + 1, x - 1), g21 = gray_at(img, cols, rows, y + 1, x), g22 = gray_at(img, cols, rows, y + 1, x + 1);
    dx = ((g02 + g22) * 3 + g12 * 10) - ((g00 + g20) * 3 + g10 * 10);
    dy = ((g22 - g02) + (g20 - g00)) * 3 + (g21 - g01) * 10;
}

// One point through all pyramid levels, exactly as cv2.calcOpticalFlowPyrLK(winSize=(15,15)) does it on its
// 128-bit SIMD build: float sums of the structure tensor / mismatch vector in four lanes over columns 0-7
// (pixels k and k+4 of a row added as integers first for the mismatch) plus a scalar tail over columns 8-14.
//   prev / next: pyramids laid out by pyramid_layout; out[2] = nextPts, returns status (1 = found)
EGL_HD_NOINLINE int lk_track_point(const uint8_t* prev, const uint8_t* next, const PyrLayout& L, float ptx, float pty, int max_count,
                                   double eps2, double min_eig_thr, float* out) {
    const float kScale = 1.f / (1 << 20);  // FLT_SCALE
    const int W = kLkWin;
    const int top = L.n - 1;
    float sx = 0.f, sy = 0.f;  // nextPts[ptidx]
    int status = 1;
    short Iw[kLkWin * kLkWin], Ix[kLkWin * kLkWin], Iy[kLkWin * kLkWin];
    short d0[kLkSup * kLkSup], d1[kLkSup * kLkSup];
    int px[kLkWin * kLkWin], py[kLkWin * kLkWin];
    for (int level = top; level >= 0; --level) {
        const int cols = L.w[level], rows = L.h[level];
        const uint8_t* I = prev + L.off[level];
        const uint8_t* J = next + L.off[level];
        const float scale = (float)(1.0 / (1 << level));
        float ppx = fmul(ptx, scale), ppy = fmul(pty, scale);
        float nx, ny;
        if (level == top) { nx = ppx; ny = ppy; } else { nx = fmul(sx, 2.f); ny = fmul(sy, 2.f); }
        sx = nx; sy = ny;
        ppx = fsub(ppx, (float)kLkHalf); ppy = fsub(ppy, (float)kLkHalf);
        const int ipx = (int)floorf(ppx), ipy = (int)floorf(ppy);
        if (ipx < -W || ipx >= cols || ipy < -W || ipy >= rows) {
            if (level == 0) status = 0;
            continue;
        }
        int w00, w01, w10, w11;
        lk_weights(fsub(ppx, (float)ipx), fsub(ppy, (float)ipy), w00, w01, w10, w11);
        for (int r = 0; r < kLkSup; ++r)
            for (int c = 0; c < kLkSup; ++c) {
                int dx, dy;
                scharr_at(I, cols, rows, ipy + r, ipx + c, dx, dy);
                d0[r * kLkSup + c] = (short)dx;
                d1[r * kLkSup + c] = (short)dy;
            }
        float lane[3][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, tail[3] = {0.f, 0.f, 0.f};
        for (int y = 0; y < W; ++y)
            for (int x = 0; x < W; ++x) {
                const int g00 = gray_at(I, cols, rows, ipy + y, ipx + x), g01 = gray_at(I, cols, rows, ipy + y, ipx + x + 1);
                const int g10 = gray_at(I, cols, rows, ipy + y + 1, ipx + x), g11 = gray_at(I, cols, rows, ipy + y + 1, ipx + x + 1);
                Iw[y * W + x] = (short)((g00 * w00 + g01 * w01 + g10 * w10 + g11 * w11 + (1 << (kLkWBits - 6))) >> (kLkWBits - 5));
                const short* a0 = d0 + y * kLkSup + x;
                const short* a1 = d1 + y * kLkSup + x;
                const int ixv = (a0[0] * w00 + a0[1] * w01 + a0[kLkSup] * w10 + a0[kLkSup + 1] * w11 + (1 << (kLkWBits - 1))) >> kLkWBits;
                const int iyv = (a1[0] * w00 + a1[1] * w01 + a1[kLkSup] * w10 + a1[kLkSup + 1] * w11 + (1 << (kLkWBits - 1))) >> kLkWBits;
                Ix[y * W + x] = (short)ixv;
                Iy[y * W + x] = (short)iyv;
                const float p[3] = {(float)(ixv * ixv), (float)(ixv * iyv), (float)(iyv * iyv)};
                for (int q = 0; q < 3; ++q) {
                    if (x < 8) lane[q][x & 3] = fadd(lane[q][x & 3], p[q]);
                    else tail[q] = fadd(tail[q], p[q]);
                }
            }
        const float A11 = fmul(fadd(tail[0], reduce4(lane[0][0], lane[0][1], lane[0][2], lane[0][3])), kScale);
        const float A12 = fmul(fadd(tail[1], reduce4(lane[1][0], lane[1][1], lane[1][2], lane[1][3])), kScale);
        const float A22 = fmul(fadd(tail[2], reduce4(lane[2][0], lane[2][1], lane[2][2], lane[2][3])), kScale);
        float D = fsub(fmul(A11, A22), fmul(A12, A12));
        const float dd = fsub(A11, A22);
        const float min_eig = fdiv(fsub(fadd(A22, A11), fsqrt(fadd(fmul(dd, dd), fmul(fmul(4.f, A12), A12)))), (float)(2 * W * W));
        if ((double)min_eig < min_eig_thr || D < 1.1920928955078125e-7f) {
            if (level == 0) status = 0;
            continue;
        }
        D = fdiv(1.f, D);
        nx = fsub(nx, (float)kLkHalf); ny = fsub(ny, (float)kLkHalf);
        float pdx = 0.f, pdy = 0.f;
        for (int it = 0; it < max_count; ++it) {
            const int inx = (int)floorf(nx), iny = (int)floorf(ny);
            if (inx < -W || inx >= cols || iny < -W || iny >= rows) {
                if (level == 0) status = 0;
                break;
            }
            lk_weights(fsub(nx, (float)inx), fsub(ny, (float)iny), w00, w01, w10, w11);
            for (int y = 0; y < W; ++y)
                for (int x = 0; x < W; ++x) {
                    const int g00 = gray_at(J, cols, rows, iny + y, inx + x), g01 = gray_at(J, cols, rows, iny + y, inx + x + 1);
                    const int g10 = gray_at(J, cols, rows, iny + y + 1, inx + x), g11 = gray_at(J, cols, rows, iny + y + 1, inx + x + 1);
                    const int diff = ((g00 * w00 + g01 * w01 + g10 * w10 + g11 * w11 + (1 << (kLkWBits - 6))) >> (kLkWBits - 5)) - Iw[y * W + x];
                    px[y * W + x] = diff * Ix[y * W + x];
                    py[y * W + x] = diff * Iy[y * W + x];
                }
            // qb0 = (bx(0,4), by(0,4), bx(1,5), by(1,5)), qb1 = (bx(2,6), by(2,6), bx(3,7), by(3,7)); pair sums as int32 first
            float qb0[4] = {0.f, 0.f, 0.f, 0.f}, qb1[4] = {0.f, 0.f, 0.f, 0.f}, t1 = 0.f, t2 = 0.f;
            for (int y = 0; y < W; ++y) {
                const int* rx = px + y * W;
                const int* ry = py + y * W;
                qb0[0] = fadd(qb0[0], (float)(rx[0] + rx[4])); qb0[1] = fadd(qb0[1], (float)(ry[0] + ry[4]));
                qb0[2] = fadd(qb0[2], (float)(rx[1] + rx[5])); qb0[3] = fadd(qb0[3], (float)(ry[1] + ry[5]));
                qb1[0] = fadd(qb1[0], (float)(rx[2] + rx[6])); qb1[1] = fadd(qb1[1], (float)(ry[2] + ry[6]));
                qb1[2] = fadd(qb1[2], (float)(rx[3] + rx[7])); qb1[3] = fadd(qb1[3], (float)(ry[3] + ry[7]));
                for (int x = 8; x < W; ++x) { t1 = fadd(t1, (float)rx[x]); t2 = fadd(t2, (float)ry[x]); }
            }
            const float q0 = fadd(qb0[0], qb1[0]), q1 = fadd(qb0[1], qb1[1]), q2 = fadd(qb0[2], qb1[2]), q3 = fadd(qb0[3], qb1[3]);
            const float b1 = fmul(fadd(t1, fadd(q0, q2)), kScale), b2 = fmul(fadd(t2, fadd(q1, q3)), kScale);
            const float dx = fmul(fsub(fmul(A12, b2), fmul(A22, b1)), D);
            const float dy = fmul(fsub(fmul(A12, b1), fmul(A11, b2)), D);
            nx = fadd(nx, dx); ny = fadd(ny, dy);
            sx = fadd(nx, (float)kLkHalf); sy = fadd(ny, (float)kLkHalf);
            if (dadd(dmul((double)dx, (double)dx), dmul((double)dy, (double)dy)) <= eps2) break;
            if (it > 0 && fabs((double)fadd(dx, pdx)) < 0.01 && fabs((double)fadd(dy, pdy)) < 0.01) {
                sx = fsub(sx, fmul(dx, 0.5f));
                sy = fsub(sy, fmul(dy, 0.5f));
                break;
            }
            pdx = dx; pdy = dy;
        }
        if (status && level == 0) {  // the error pass re-checks the final window position
            const int fx = (int)floorf(fsub(sx, (float)kLkHalf)), fy = (int)floorf(fsub(sy, (float)kLkHalf));
            if (fx < -W || fx >= cols || fy < -W || fy >= rows) status = 0;
        }
    }
    out[0] = sx;
    out[1] = sy;
    return status;
}

}  // namespace egl
