// K1: uint8 BGR frame -> float32 RGB planes at 540x960, normalised (replaces
// cv2.cvtColor(BGR2RGB) + A.Resize(540,960) + A.Normalize() + ToTensorV2 + .float(),
// eagle/models/coordinate_model.py:62-64,221-222,489-491).
//
// The resize is OpenCV's INTER_LINEAR for uint8, reproduced bit for bit in integer arithmetic:
// 11-bit fixed-point tap weights, horizontal pass S[sx]*a0 + S[sx+1]*a1, vertical pass
// ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2, uint8 result BEFORE normalisation (at exact 2x
// decimation OpenCV switches to its 2x2-mean fast path, which this recipe equals identically).
//
// HBM-bound.  One CTA per band of R output rows: the source rows it needs are pulled into shared
// memory with bulk copies on the TMA engine (fully coalesced, no registers), then 240 threads x 4
// interleaved output pixels compute all three channels from shared memory (conflict-free byte reads)
// and write one float per plane, pixel and row (128 contiguous bytes per warp and store).  Every source byte and every output float crosses HBM
// once when the vertical scale is >= 2 (1080p, 4K); at 720p neighbouring output rows share source
// rows and the second read is served by L2.
#include <stdlib.h>

#include "common.cuh"
#include "geometry_core.cuh"

namespace egl {

constexpr int kOutW = EGL_MODEL_W, kOutH = EGL_MODEL_H;
constexpr int kPreThreads = kOutW / 4;  // 240

// ImageNet mean/std of A.Normalize(): float32 mean*255 and 1/(std*255), formed as albucore forms them
__device__ __forceinline__ void norm_consts(float* m, float* d) {
    m[0] = __fmul_rn(0.485f, 255.f); m[1] = __fmul_rn(0.456f, 255.f); m[2] = __fmul_rn(0.406f, 255.f);
    d[0] = __fdiv_rn(1.f, __fmul_rn(0.229f, 255.f));
    d[1] = __fdiv_rn(1.f, __fmul_rn(0.224f, 255.f));
    d[2] = __fdiv_rn(1.f, __fmul_rn(0.225f, 255.f));
}

// OpenCV's tap for output index d along an axis of `src` samples: offset, weights (x2048).
__device__ __forceinline__ void axis_tap(int d, double scale, int* ofs, int* c0, int* c1) {
    float f = (float)__dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    const int s = (int)floorf(f);
    f = __fsub_rn(f, (float)s);
    *ofs = s;
    *c0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));
    *c1 = __float2int_rn(__fmul_rn(f, 2048.f));
}

// R consecutive output rows per CTA.  The 2R source rows they need are requested up front (one
// mbarrier, one bulk copy per distinct row: when the vertical scale is < 2 neighbouring output rows
// share a source row, which is then fetched once and aliased), so a CTA keeps up to 2R rows in
// flight and the per-CTA fixed costs (launch, barrier set-up, tap arithmetic) are amortised R-fold.
//
// kMean4: when the frame is an even integer multiple s of 960x540 (1080p: s = 2, 4K: s = 4) both taps of
// both passes weigh exactly 1024/2048, the fixed-point recipe collapses to the rounded mean of four source
// pixels, (S[y0][x0] + S[y0][x0+1] + S[y0+1][x0] + S[y0+1][x0+1] + 2) >> 2 with x0 = (s/2)(2dx+1) - 1 (the same
// value, bit for bit -- it is also OpenCV's own INTER_AREA shortcut at s = 2), at a third of the
// integer instructions; `half_scale` = s/2 is passed in scale_x's place.
//
// kDual: the same resized uint8 image also leaves as the detector's letterboxed input (ultralytics LetterBox + predictor
// preprocess for imgsz = 960 on a 16:9 frame: INTER_LINEAR resize to 960x540, rows of 114 above and below up to a
// stride-32 multiple, BGR -> RGB, HWC -> CHW, float32, / 255): out2[f][plane][pad_top + dy][dx] = v / 255, the pad rows
// written by the first and the last band of a frame.  One read of the frame feeds both networks.
//
// kPair (exact 2x decimation, 1080p -> 540x960: the headline case): the byte-wise reads of the general code make the
// kernel shared-memory bound -- four LDS.U8 per output value, two wavefronts each because 32 lanes x 6 bytes span 192
// bytes: 8 wavefronts of the SM's one-per-cycle LSU data path per 32 outputs, 98.7 % of its peak in the ncu capture, and
// the limit instead of HBM as soon as the SM clock sags under the power cap.  Here a thread takes two ADJACENT output
// pixels, i.e. 12 contiguous source bytes per row = three aligned 32-bit words (lane stride 3 words: conflict-free, one
// wavefront per load), sorts the bytes with two PRMT per row and forms every four-byte sum with two dp4a: one wavefront
// and about 8 instructions per output value instead of 8 and 15.  The two floats of a plane leave as one 64-bit store.
template <int R, bool kMean4, bool kDual = false, bool kPair = false>
__global__ void __launch_bounds__(kPreThreads) preprocess_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                                 size_t row_stride, size_t frame_stride, double scale_x,
                                                                 double scale_y, int use_bulk, float* __restrict__ out,
                                                                 float* __restrict__ out2 = nullptr, int out2_h = 0, int pad_top = 0) {
    extern __shared__ __align__(128) unsigned char s_rows[];  // 2R row slots, each padded to a multiple of 16 B
    __shared__ uint64_t s_bar;
    constexpr int kBands = kOutH / R;
    static_assert(kOutH % R == 0, "R must divide 540");
    const int tid = threadIdx.x;
    const int dy0 = (blockIdx.x % kBands) * R;
    const int f = blockIdx.x / kBands;
    const int row_bytes = 3 * W;
    const int row_pad = (row_bytes + 15) & ~15;

    // vertical taps of the R rows; slot[] maps (row, tap) to a shared-memory slot, aliasing repeats
    int b0[R], b1[R], ysrc[2 * R], slot[2 * R];
    int nslots = 0;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        int sy;
        if (kMean4) {
            sy = (int)scale_y * (2 * (dy0 + r) + 1) - 1;
            b0[r] = b1[r] = 1024;
        } else {
            axis_tap(dy0 + r, scale_y, &sy, &b0[r], &b1[r]);
        }
        ysrc[2 * r] = min(max(sy, 0), H - 1);  // rows clamp, weights do not
        ysrc[2 * r + 1] = min(max(sy + 1, 0), H - 1);
    }
#pragma unroll
    for (int q = 0; q < 2 * R; ++q) {
        slot[q] = (q > 0 && ysrc[q] == ysrc[q - 1]) ? slot[q - 1] : nslots++;
    }
    const uint8_t* fbase = frames + (size_t)f * frame_stride;
    if (use_bulk) {
        if (tid == 0) {
            mbar_init(&s_bar, 1);
            mbar_fence_init();
            mbar_expect_tx(&s_bar, (uint32_t)nslots * (uint32_t)row_bytes);
#pragma unroll
            for (int q = 0; q < 2 * R; ++q)
                if (q == 0 || slot[q] != slot[q - 1])
                    bulk_g2s(s_rows + (size_t)slot[q] * row_pad, fbase + (size_t)ysrc[q] * row_stride, (uint32_t)row_bytes, &s_bar);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 2 * R; ++q)
            if (q == 0 || slot[q] != slot[q - 1])
                for (int i = tid; i < row_bytes; i += kPreThreads)
                    s_rows[(size_t)slot[q] * row_pad + i] = fbase[(size_t)ysrc[q] * row_stride + i];
    }

    // horizontal taps of this thread's 4 output pixels tid, tid+240, tid+480, tid+720 (independent of
    // the data: overlaps the copy).  Interleaving the pixels over the threads keeps the byte reads of a
    // warp on consecutive shared-memory words (no bank conflicts at any scale) and every store of a
    // warp on one 128-byte line.
    int xo[4], xo1[4], a0[4], a1[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        int s, c0, c1;
        if (kMean4) {
            s = (int)scale_x * (2 * (tid + kPreThreads * k) + 1) - 1;
            c0 = c1 = 1024;
        } else {
            axis_tap(tid + kPreThreads * k, scale_x, &s, &c0, &c1);
        }
        if (s < 0) { s = 0; c0 = 2048; c1 = 0; }
        if (s >= W - 1) { s = W - 1; c0 = 2048; c1 = 0; }
        xo[k] = 3 * s;
        xo1[k] = 3 * min(s + 1, W - 1);
        a0[k] = c0;
        a1[k] = c1;
    }
    float mean[3], den[3];
    norm_consts(mean, den);

    __syncthreads();  // barrier init (bulk) / staged rows (fallback) visible
    if (use_bulk) mbar_wait(&s_bar, 0);
    if constexpr (kPair) {
        static_assert(kMean4, "kPair is the 2x2-mean case");
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const uint32_t* r0 = reinterpret_cast<const uint32_t*>(s_rows + (size_t)slot[2 * r] * row_pad);
            const uint32_t* r1 = reinterpret_cast<const uint32_t*>(s_rows + (size_t)slot[2 * r + 1] * row_pad);
            float* dst = out + ((size_t)f * 3 * kOutH + dy0 + r) * kOutW;
#pragma unroll
            for (int k = 0; k < 2; ++k) {
                const int j = tid + kPreThreads * k;   // output pixels 2j, 2j+1 <- source pixels 4j .. 4j+3 = bytes [12j, 12j+12)
                const uint32_t a0 = r0[3 * j], a1 = r0[3 * j + 1], a2 = r0[3 * j + 2];   // [B0 G0 R0 B1] [G1 R1 B2 G2] [R2 B3 G3 R3]
                const uint32_t c0 = r1[3 * j], c1 = r1[3 * j + 1], c2 = r1[3 * j + 2];
                const uint32_t xa = __byte_perm(a0, a1, 0x5241), xc = __byte_perm(c0, c1, 0x5241);   // [G0 G1 R0 R1]
                const uint32_t ya = __byte_perm(a1, a2, 0x6352), yc = __byte_perm(c1, c2, 0x6352);   // [B2 B3 G2 G3]
                // sums of four bytes + 2 (dp4a with a 0/1 selector), then >> 2: the rounded 2x2 mean
                const int bA = (int)(__dp4a(a0, 0x01000001u, __dp4a(c0, 0x01000001u, 2u)) >> 2);
                const int gA = (int)(__dp4a(xa, 0x00000101u, __dp4a(xc, 0x00000101u, 2u)) >> 2);
                const int rA = (int)(__dp4a(xa, 0x01010000u, __dp4a(xc, 0x01010000u, 2u)) >> 2);
                const int bB = (int)(__dp4a(ya, 0x00000101u, __dp4a(yc, 0x00000101u, 2u)) >> 2);
                const int gB = (int)(__dp4a(ya, 0x01010000u, __dp4a(yc, 0x01010000u, 2u)) >> 2);
                const int rB = (int)(__dp4a(a2, 0x01000001u, __dp4a(c2, 0x01000001u, 2u)) >> 2);
                const int vA[3] = {rA, gA, bA}, vB[3] = {rB, gB, bB};   // planes are R, G, B
#pragma unroll
                for (int plane = 0; plane < 3; ++plane) {
                    __stcs(reinterpret_cast<float2*>(dst + (size_t)plane * kOutH * kOutW + 2 * j),
                           make_float2(__fmul_rn(__fsub_rn((float)vA[plane], mean[plane]), den[plane]),
                                       __fmul_rn(__fsub_rn((float)vB[plane], mean[plane]), den[plane])));
                    if (kDual)
                        __stcs(reinterpret_cast<float2*>(out2 + (((size_t)f * 3 + plane) * out2_h + pad_top + dy0 + r) * kOutW + 2 * j),
                               make_float2(__fdiv_rn((float)vA[plane], 255.f), __fdiv_rn((float)vB[plane], 255.f)));
                }
            }
        }
    } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const uint8_t* r0 = s_rows + (size_t)slot[2 * r] * row_pad;
        const uint8_t* r1 = s_rows + (size_t)slot[2 * r + 1] * row_pad;
        float* dst = out + ((size_t)f * 3 * kOutH + dy0 + r) * kOutW + tid;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {  // c indexes the SOURCE byte (B,G,R); plane = 2 - c (RGB)
                int v;
                if (kMean4) {
                    v = ((int)r0[xo[k] + c] + (int)r0[xo[k] + 3 + c] + (int)r1[xo[k] + c] + (int)r1[xo[k] + 3 + c] + 2) >> 2;
                } else {
                    const int top = (int)r0[xo[k] + c] * a0[k] + (int)r0[xo1[k] + c] * a1[k];
                    const int bot = (int)r1[xo[k] + c] * a0[k] + (int)r1[xo1[k] + c] * a1[k];
                    v = (((b0[r] * (top >> 4)) >> 16) + ((b1[r] * (bot >> 4)) >> 16) + 2) >> 2;
                    v = min(max(v, 0), 255);
                }
                const int plane = 2 - c;
                __stcs(dst + (size_t)plane * kOutH * kOutW + kPreThreads * k, __fmul_rn(__fsub_rn((float)v, mean[plane]), den[plane]));
                if (kDual)
                    __stcs(out2 + (((size_t)f * 3 + plane) * out2_h + pad_top + dy0 + r) * kOutW + tid + kPreThreads * k,
                           __fdiv_rn((float)v, 255.f));
            }
        }
    }
    }
    if (kDual) {
        // letterbox border: rows [0, pad_top) by the first band of the frame, rows [pad_top + 540, out2_h) by the last
        const float border = __fdiv_rn(114.f, 255.f);
        const bool first = dy0 == 0, last = dy0 + R == kOutH;
        if (first || last) {
            const int y_lo = first ? 0 : pad_top + kOutH, y_hi = first ? pad_top : out2_h;
            for (int plane = 0; plane < 3; ++plane)
                for (int y = y_lo; y < y_hi; ++y)
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        __stcs(out2 + (((size_t)f * 3 + plane) * out2_h + y) * kOutW + tid + kPreThreads * k, border);
        }
    }
}

}  // namespace egl

using namespace egl;

static int preprocess_impl(const uint8_t* frames, int F, int H, int W, size_t row_stride, size_t frame_stride, float* out, float* out2,
                           int out2_h, int pad_top, void* stream) {
    if (F == 0) return 0;  // empty batch: nothing to enqueue, pointers may be null
    EGL_REQUIRE(frames && out, EGL_ERR_NULL, "egl_preprocess_u8: null pointer");
    EGL_REQUIRE(F >= 0 && H >= 2 && W >= 2, EGL_ERR_SHAPE, "egl_preprocess_u8: bad shape %dx%d", H, W);
    EGL_REQUIRE(row_stride >= (size_t)3 * W && frame_stride >= row_stride * (size_t)H, EGL_ERR_SHAPE,
                "egl_preprocess_u8: strides smaller than the frame");
    EGL_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15) == 0, EGL_ERR_ALIGN, "egl_preprocess_u8: out must be 16-byte aligned");
    EGL_REQUIRE((long long)F * kOutH < (1ll << 31), EGL_ERR_SHAPE, "egl_preprocess_u8: too many frames for one launch");
    if (F == 0) return 0;
    const int row_bytes = 3 * W;
    const int row_pad = (row_bytes + 15) & ~15;
    // bulk copies need 16-byte aligned rows of a 16-byte multiple length
    const int use_bulk = (row_bytes % 16 == 0) && (row_stride % 16 == 0) && (frame_stride % 16 == 0) &&
                         ((reinterpret_cast<uintptr_t>(frames) & 15) == 0);
    // OpenCV: inv_scale = dsize/ssize (double); scale = 1./inv_scale
    const double scale_x = 1. / ((double)kOutW / (double)W);
    const double scale_y = 1. / ((double)kOutH / (double)H);
    // even integer multiple of the output size -> rounded mean of four (see kMean4)
    bool mean4 = (W % kOutW == 0) && (H % kOutH == 0) && (W / kOutW == H / kOutH) && ((W / kOutW) % 2 == 0);
#ifdef EGL_BENCH_VARIANTS
    static const char* mean4_env = getenv("EGL_PREPROCESS_MEAN4");  // measurement switch: 0 forces the general recipe
    if (mean4_env && atoi(mean4_env) == 0) mean4 = false;
#endif
    const double arg_x = mean4 ? (double)(W / kOutW / 2) : scale_x, arg_y = mean4 ? (double)(H / kOutH / 2) : scale_y;
    // output rows per CTA: 4 unless the row slots of a CTA would exceed ~100 KB of shared memory
    int R = 4;
#ifdef EGL_BENCH_VARIANTS
    static const char* rows_env = getenv("EGL_PREPROCESS_ROWS");  // measurement switch: rows per CTA (1, 2, 3, 4, 6)
    if (rows_env) R = atoi(rows_env);
    if (R != 1 && R != 2 && R != 3 && R != 4 && R != 6) R = 4;
    if (mean4 && R == 3) R = 2;
#endif
    while (R > 1 && (size_t)2 * R * row_pad > 100 * 1024) R = (R == 6) ? 4 : (R == 3 ? 2 : R / 2);
    const size_t smem = (size_t)2 * R * row_pad;
    EGL_REQUIRE(smem <= 200 * 1024, EGL_ERR_SHAPE, "egl_preprocess_u8: frame too wide (%d px)", W);
    cudaStream_t st = (cudaStream_t)stream;
    auto launch = [&](auto kernel, int rows) -> int {
        if (smem > 48 * 1024) {
            int rc = cuda_status(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                                 "egl_preprocess_u8: cudaFuncSetAttribute");
            if (rc) return rc;
        }
        kernel<<<(unsigned)(F * (kOutH / rows)), kPreThreads, smem, st>>>(frames, H, W, row_stride, frame_stride, arg_x, arg_y,
                                                                         use_bulk, out, out2, out2_h, pad_top);
        return 0;
    };
    // exact 2x decimation (1080p): word loads + dp4a (kPair); EGL_PREPROCESS_PAIR=0 in measurement builds selects the byte-wise code
    bool pair = mean4 && W == 2 * kOutW && R == 4;
#ifdef EGL_BENCH_VARIANTS
    static const char* pair_env = getenv("EGL_PREPROCESS_PAIR");
    if (pair_env && atoi(pair_env) == 0) pair = false;
#endif
    int rc;
    if (pair) {   // 1920-pixel rows: eight row slots are 46 KB, so R is always 4 here
        rc = out2 ? launch(preprocess_kernel<4, true, true, true>, 4) : launch(preprocess_kernel<4, true, false, true>, 4);
    } else if (out2) {
        if (R == 3 || R == 6) R = 4;
        if (mean4) {
            switch (R) {
                case 1: rc = launch(preprocess_kernel<1, true, true>, 1); break;
                case 2: rc = launch(preprocess_kernel<2, true, true>, 2); break;
                default: rc = launch(preprocess_kernel<4, true, true>, 4); break;
            }
        } else {
            switch (R) {
                case 1: rc = launch(preprocess_kernel<1, false, true>, 1); break;
                case 2: rc = launch(preprocess_kernel<2, false, true>, 2); break;
                default: rc = launch(preprocess_kernel<4, false, true>, 4); break;
            }
        }
    } else if (mean4) {
        switch (R) {
            case 1: rc = launch(preprocess_kernel<1, true>, 1); break;
            case 2: rc = launch(preprocess_kernel<2, true>, 2); break;
#ifdef EGL_BENCH_VARIANTS
            case 6: rc = launch(preprocess_kernel<6, true>, 6); break;
#endif
            default: rc = launch(preprocess_kernel<4, true>, 4); break;
        }
    } else {
        switch (R) {
            case 1: rc = launch(preprocess_kernel<1, false>, 1); break;
            case 2: rc = launch(preprocess_kernel<2, false>, 2); break;
#ifdef EGL_BENCH_VARIANTS
            case 3: rc = launch(preprocess_kernel<3, false>, 3); break;
            case 6: rc = launch(preprocess_kernel<6, false>, 6); break;
#endif
            default: rc = launch(preprocess_kernel<4, false>, 4); break;
        }
    }
    if (rc) return rc;
    return cuda_status(cudaGetLastError(), "egl_preprocess_u8: kernel launch");
}

extern "C" int egl_preprocess_u8(const uint8_t* frames, int F, int H, int W, size_t row_stride, size_t frame_stride,
                                 float* out, void* stream) {
    return preprocess_impl(frames, F, H, W, row_stride, frame_stride, out, nullptr, 0, 0, stream);
}

extern "C" int egl_preprocess_u8_letterbox(const uint8_t* frames, int F, int H, int W, size_t row_stride, size_t frame_stride,
                                           float* out, float* out_detector, int detector_h, int pad_top, void* stream) {
    if (F == 0) return 0;
    EGL_REQUIRE(out_detector, EGL_ERR_NULL, "egl_preprocess_u8_letterbox: out_detector is null");
    EGL_REQUIRE((reinterpret_cast<uintptr_t>(out_detector) & 15) == 0, EGL_ERR_ALIGN, "egl_preprocess_u8_letterbox: out_detector must be 16-byte aligned");
    EGL_REQUIRE(pad_top >= 0 && detector_h >= pad_top + kOutH && detector_h <= kOutH + 64, EGL_ERR_SHAPE,
                "egl_preprocess_u8_letterbox: need 0 <= pad_top and pad_top + 540 <= detector_h <= 604 (got pad_top %d, detector_h %d)", pad_top, detector_h);
    // The detector image shares the keypoint network's resized frame only when LetterBox's own resize is W x H -> 960 x 540,
    // i.e. round(W r) == 960 and round(H r) == 540 for r = min(960 / H, 960 / W): 16:9 frames at imgsz 960.
    const double r = fmin(960.0 / H, 960.0 / W);
    EGL_REQUIRE((int)nearbyint(W * r) == kOutW && (int)nearbyint(H * r) == kOutH, EGL_ERR_SHAPE,
                "egl_preprocess_u8_letterbox: a %dx%d frame does not letterbox to 960x540 at imgsz 960", W, H);
    return preprocess_impl(frames, F, H, W, row_stride, frame_stride, out, out_detector, detector_h, pad_top, stream);
}
