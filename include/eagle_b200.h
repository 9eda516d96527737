/*
 * eagle_b200.h -- C ABI of the B200-native per-frame geometry path.
 *
 * The reference (nreHieW/Eagle) is pure Python and has no FFI of its own; the entry points below
 * are what a binding for its geometry path attaches to.  Each one replaces the reference
 * statements cited beside it (paths relative to the reference root).  INTEGRATION.md shows the
 * ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *   - extern "C", plain pointers and sizes; no exceptions; no allocation; no global state other
 *     than a thread-local last-error string.
 *   - every pointer is a DEVICE pointer unless its name ends in _host; the caller owns all memory.
 *   - functions only ENQUEUE work on `stream` (cudaStream_t passed as void*) and return at once.
 *   - return value: 0 = enqueued, >0 = argument error (EGL_ERR_*), <0 = -(cudaError_t).
 *   - per-frame soft failures (too few landmarks, degenerate geometry) are reported in the
 *     per-frame status[] array, never through the return code: the reference never raises on this
 *     path either (coordinate_model.py:350-352,366-367).
 */
#ifndef EAGLE_B200_H_
#define EAGLE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EGL_ABI_VERSION 1

#define EGL_NUM_LANDMARKS 57 /* heatmap channels; eagle/utils/pitch.py:1-59 */
#define EGL_ORDER_STRIDE 64  /* bytes per frame in kp_order[] */
#define EGL_MODEL_H 540      /* A.Resize(540, 960); coordinate_model.py:63 */
#define EGL_MODEL_W 960

/* argument errors */
#define EGL_ERR_NULL 1
#define EGL_ERR_SHAPE 2
#define EGL_ERR_ALIGN 3
#define EGL_ERR_MODE 4

/* status[] values written by egl_fit_homography */
#define EGL_FIT_OK 0         /* H, masks valid */
#define EGL_FIT_FEW_POINTS 1 /* < 4 on-plane landmarks: reference sets compute_homography=True (:350-352) */
#define EGL_FIT_NO_MODEL 2   /* no leg of the cascade found a model (cv2.findHomography returned None three times, :363-367) */
#define EGL_FIT_SKIPPED 3    /* egl_fit_homography_masked: the cadence did not ask for a fit on this frame */

/* kp_src[] values: which Python type the reference holds a keypoint value in (it reaches the JSON) */
#define EGL_KP_PY_INT 0    /* tuple of int: decoded (:248) or synthesised (:179) */
#define EGL_KP_NUMPY_INT 1 /* tuple of numpy int64: optical flow (:476) or calibration (:553) */
#define EGL_KP_FLOAT 2     /* list of float: inlier of a successful fit (:361) */

/* info[2] of a frame whose H comes from a later leg of `for method in [cv2.RANSAC, cv2.RHO, cv2.LMEDS]` (:354) */
#define EGL_FIT_LEG_RHO (-2)
#define EGL_FIT_LEG_LMEDS (-3)

/* fit modes */
#define EGL_FIT_CV2_COMPAT 1 /* OpenCV's RNG, sampling, adaptive stopping: picks the model cv2 picks */
#define EGL_FIT_FIXED_K 0    /* K hypotheses per frame from an explicit table or the seeded generator */

int egl_version(void);
const char *egl_last_error(void);

/* Number of SMs of the current device (grid sizing is a multiple of it). */
int egl_sm_count(void);

/* Bit 0: the library was built with -DEGL_BENCH_VARIANTS and holds the alternative kernels / EGL_*_VARIANT
 * environment switches used for A/B measurements.  The shipped build returns 0: one kernel per job. */
int egl_build_flags(void);

/*
 * K1  uint8 BGR frames -> normalised float32 RGB planes for the keypoint network.
 * Replaces cv2.cvtColor(BGR2RGB) + A.Resize(540,960) + A.Normalize() + ToTensorV2 + .float()
 * (coordinate_model.py:62-64, :221-222, :489-491): OpenCV's fixed-point INTER_LINEAR resize
 * (uint8 result, bit-exact), then (v - 255*mean) * (1/(255*std)), HWC -> CHW.
 *   frames  [F][H][row_stride] uint8, pixel = B,G,R; row_stride >= 3*W bytes; frame_stride bytes
 *   out     [F][3][540][960] float32
 */
int egl_preprocess_u8(const uint8_t *frames, int F, int H, int W, size_t row_stride, size_t frame_stride,
                      float *out, void *stream);

/*
 * F3 (second output of K1)  the same pass also writes the DETECTOR's input tensor, so that one read of the uint8 frame
 * feeds both networks.  Replaces what `self.detector_model(frame, ...)` (coordinate_model.py:568) does to the frame before
 * the network sees it -- ultralytics 8.3.184 (uv.lock:1814-1815; third-party, not vendored in the reference):
 * LetterBox(new_shape = imgsz, auto = True, stride = 32) = cv2.resize(INTER_LINEAR) to (round(W r), round(H r)),
 * r = min(imgsz / H, imgsz / W), cv2.copyMakeBorder with 114 up to a stride-32 multiple, then BGR -> RGB, HWC -> CHW,
 * .float(), /= 255 (BasePredictor.preprocess).  Supported where LetterBox's resize is the keypoint network's own, W x H ->
 * 960 x 540 (imgsz 960 on 16:9 frames: 1080p, 720p, 4K ...); anything else returns EGL_ERR_SHAPE.  A maintainer passes
 * the tensor to the model instead of the frame (ultralytics skips its own preprocessing for BCHW float tensors).
 *   out           as egl_preprocess_u8
 *   out_detector  [F][3][detector_h][960] float32, RGB planes, resized uint8 / 255 in rows [pad_top, pad_top + 540),
 *                 114 / 255 elsewhere (for 16:9 at imgsz 960: detector_h = 544, pad_top = 2)
 * Parity: against the same cv2 calls + numpy (oracle/preprocess.py::letterbox_reference_calls); ultralytics itself is
 * absent offline, so the restatement of ITS arithmetic is unpinned.
 */
int egl_preprocess_u8_letterbox(const uint8_t *frames, int F, int H, int W, size_t row_stride, size_t frame_stride,
                                float *out, float *out_detector, int detector_h, int pad_top, void *stream);

/*
 * K2  heatmaps -> landmark pixel positions.
 * Replaces KeypointModel.get_keypoints (keypoint_hrnet.py:583-594: per channel flat argmax, first
 * maximum in row-major order, score = max) and the keypoint post-processing block
 * (coordinate_model.py:229-248: drop score < keypoint_conf, xi = int(x/(w-1) * img_w),
 * yi = int(y/(h-1) * img_h), labels sharing a pixel keep the best score).
 *   hm         [F][57][hm_h][hm_w] float32, 16-byte aligned, hm_h*hm_w % 4 == 0
 *   kp_flat    [F][57] int32   argmax flat index (all channels)
 *   kp_score   [F][57] float   maximum (all channels)
 *   kp_xy      [F][57][2] int32 image pixel (valid for channels listed in kp_order)
 *   kp_order   [F][64] uint8   kept channels in the reference's dict insertion order
 *   kp_count   [F][2] int32    {entries in kp_order, entries that came from the heatmaps}
 */
int egl_decode_heatmaps(const float *hm, int F, int hm_h, int hm_w, int img_w, int img_h, double keypoint_conf,
                        int32_t *kp_flat, float *kp_score, int32_t *kp_xy, uint8_t *kp_order, int32_t *kp_count,
                        void *stream);

/*
 * F3  same as egl_decode_heatmaps, but reads the network's LOGITS (KeypointModel.forward_unnormalized,
 * keypoint_hrnet.py:572-573) and applies the sigmoid of KeypointModel.forward (:565-570) inside the
 * arg-max kernel: one pass over HBM instead of sigmoid (read + write) followed by decode (read).
 * kp_score holds sigmoid(logit); ties of the float sigmoid (saturation at 1.0, plateaus) resolve to the
 * first index exactly as sigmoid-then-np.argmax does.
 */
int egl_decode_logits(const float *logits, int F, int hm_h, int hm_w, int img_w, int img_h, double keypoint_conf,
                      int32_t *kp_flat, float *kp_score, int32_t *kp_xy, uint8_t *kp_order, int32_t *kp_count,
                      void *stream);

/*
 * Sub-pixel landmark positions (an extension: the reference stops at the integer heatmap grid, which
 * alone costs ~1e-2 relative error in H).  Per channel a parabola through the arg-max and its two
 * neighbours per axis, offset clamped to +-0.5 px, 0 on the border; image position
 * (x + d) / (w - 1) * img_w as float.  Parity mode (egl_fit_homography on the integer kp_xy) is unaffected.
 *   kp_flat [F][57] int32 from egl_decode_*;  kp_sub [F][57][2] float32 out (all channels)
 */
int egl_refine_keypoints(const float *hm, int F, int hm_h, int hm_w, int img_w, int img_h, const int32_t *kp_flat,
                         float *kp_sub, void *stream);

/* egl_fit_homography on the sub-pixel positions kp_sub instead of the integer kp_xy (same outputs). */
int egl_fit_homography_subpixel(const float *kp_sub, const int32_t *kp_xy, const uint8_t *kp_order,
                                const int32_t *kp_count, int F, int mode, int K, const uint8_t *hyp, uint64_t seed,
                                double thr, double confidence, double *H, uint64_t *used_mask, uint64_t *inlier_mask,
                                int32_t *status, int32_t *info, void *stream);

/*
 * F1  line-intersection keypoint synthesis, appended to kp_order / kp_xy in place.
 * Replaces CoordinateModel._synthesize_keypoints_with_line_intersections
 * (coordinate_model.py:140-186, with :76-138): per world-y and world-x line family with >= 2
 * detected on-plane members, cv2.fitLine(DIST_L2); intersections that are labelled landmarks not
 * yet present are added (rounded half-to-even), at most max_new (30).
 */
int egl_synthesize_keypoints(int32_t *kp_xy, uint8_t *kp_order, int32_t *kp_count, int F, int max_new, void *stream);

/*
 * Host frames -> device memory: the first step of get_coordinates on a list of frames (coordinate_model.py:188,221 take
 * `frames` as a Python list of pageable HxWx3 uint8 arrays).  frames[i] = host pointer of frame i (any memory, pageable
 * or not), every frame bytes_per_frame long; dst = device buffer of n_frames * bytes_per_frame.  n_threads worker threads
 * copy 4 MiB slices through small page-locked rings of their own and issue the H2D of every slice at once on streams of
 * their own, so the host copy and the DMA interleave at slice granularity.  Synchronous: returns when all
 * frames are on the device (run it on a helper thread to overlap it with other work).  One upload at a time per process.
 */
int egl_upload_frames(const void *const *frames, int n_frames, size_t bytes_per_frame, void *dst, int n_threads);

/*
 * K3  robust image->pitch homography per frame.
 * Replaces the correspondence gather (coordinate_model.py:335-349: on-plane channels of kp_order,
 * in order) and cv2.findHomography(img_pts, world_pts, cv2.RANSAC, 5.0) (:354-357), including
 * OpenCV's least-squares refit on the inliers, Levenberg-Marquardt polish and the final mask
 * recomputed from the refined H.  In mode EGL_FIT_CV2_COMPAT a frame the RANSAC leg leaves without a model goes
 * through the rest of the reference's cascade, cv2.findHomography(..., cv2.RHO, None) and then
 * (..., cv2.LMEDS, None) (:354-357); info[2] says which leg produced H.
 *   mode EGL_FIT_CV2_COMPAT: K = iteration cap (2000 = OpenCV's default), hyp/seed ignored.
 *   mode EGL_FIT_FIXED_K:    K hypotheses per frame; hyp = [F][K][4] uint8 indices into the frame's
 *                            point list, or NULL to draw them from the counter-based generator (seed).
 *   thr       reprojection threshold in pitch metres (5.0); confidence 0.995 = OpenCV's default
 *   H         [F][9] float64 row-major, H[8] == 1 (unchanged when status != EGL_FIT_OK)
 *   used_mask / inlier_mask [F] uint64, bit = channel
 *   status    [F] int32 (EGL_FIT_*)
 *   info      [F][4] int32 {points used, inliers, winning hypothesis index (or EGL_FIT_LEG_*), hypotheses evaluated}
 */
int egl_fit_homography(const int32_t *kp_xy, const uint8_t *kp_order, const int32_t *kp_count, int F, int mode,
                       int K, const uint8_t *hyp, uint64_t seed, double thr, double confidence, double *H,
                       uint64_t *used_mask, uint64_t *inlier_mask, int32_t *status, int32_t *info, void *stream);

/*
 * Cadence: which fit does frame f project with?  Replaces the reference's homography state machine
 * (coordinate_model.py:333 "i % homography_interval == 0 or compute_homography", :350-367 retry on
 * failure, :375-378 H_use = current else previous) evaluated over fits computed for every frame.
 *   status    [F] int32 from egl_fit_homography
 *   interval  homography_interval (>= 1); carry_in = row of H valid before frame 0, or -1
 *   h_index   [F] int32 out: row of H for frame f, -1 = none yet
 *   attempted [F] uint8 out: 1 where the reference would have called findHomography on frame f
 */
int egl_select_homography(const int32_t *status, int F, int interval, int carry_in, int32_t *h_index,
                          uint8_t *attempted, void *stream);

/*
 * The same cadence, one chunk of a clip at a time (so that a long clip streams through fixed buffers and every
 * chunk can be projected and handed to the host while the next is still being decoded).
 *   status      [F] int32: the chunk's fit statuses; the chunk starts at clip frame first_frame, which must be a
 *               multiple of interval (cadence segments then never straddle a chunk boundary)
 *   carry_in    device pointer to the CLIP row of H valid before this chunk (-1 = none), or NULL = none
 *   carry_out   device pointer that receives the row valid after this chunk (may alias carry_in), or NULL
 *   h_index     [F] int32 out: CLIP row of H for every frame of the chunk (H is indexed by clip frame)
 */
int egl_select_homography_chunk(const int32_t *status, int F, int interval, int first_frame, const int32_t *carry_in,
                                int32_t *carry_out, int32_t *h_index, uint8_t *attempted, void *stream);

/*
 * K4  projection of all foot points of every frame + visible-pitch boundaries.
 * Replaces the per-object cv2.perspectiveTransform loop (coordinate_model.py:369-392), the corner
 * projection and find_x_at_y (:396-414, :32-44).
 *   H        [*][9] float64; h_index [F] int32 = row of H to use for frame f, or -1 for "no
 *            homography yet" (reference H_use logic, :375-378); NULL = identity mapping f -> f
 *   pts      [F][P][2] float32 foot points (Bottom_center), npts [F] int32 (<= P)
 *   out_f    [F][P][2] float32 projected pitch coordinates before truncation
 *   out_i    [F][P][2] int64   C-truncated (numpy .astype(int)); INT64_MIN for non-finite
 *   inb      [F][P] uint8      1 = 0 <= X <= 105 and 0 <= Y <= 68 (-> Transformed_Coordinates kept)
 *   bounds   [F][4] float64    x of [bottom_left, top_left, top_right, bottom_right]; all NaN when
 *                              the reference would emit [None]*4
 */
int egl_project_points(const double *H, const int32_t *h_index, const float *pts, const int32_t *npts, int F, int P,
                       int img_w, int img_h, float *out_f, int64_t *out_i, uint8_t *inb, double *bounds,
                       void *stream);

/* ------------------------------------------------------------------------------------------------
 * F4  keypoint propagation between network frames (the sparse keypoint cadence, coordinate_model.py:206).
 *
 * A "keypoint set" is the per-frame dict the reference carries: kp_xy [n][57][2] int32 (by channel),
 * kp_order [n][64] uint8 (insertion order), kp_count [n][2] int32 ({entries, -}), kp_src [n][64] uint8
 * (EGL_KP_* by channel).  Frames are addressed as  index0 + p * frame_step  (p = 0..n-1) inside a
 * device array of frames / pyramids, so that one call advances every chain of a clip by one frame.
 * ------------------------------------------------------------------------------------------------ */

/* Bytes per frame of the gray pyramid egl_gray_pyramid writes (levels stop before one is <= 15 px). */
int64_t egl_pyramid_bytes(int H, int W, int max_level);

/*
 * cv2.cvtColor(frame, COLOR_BGR2GRAY) (coordinate_model.py:281) + the image pyramid
 * cv2.calcOpticalFlowPyrLK builds from it (cv2.pyrDown per level), bit-exact.
 *   pyr  [F][egl_pyramid_bytes] uint8: level 0 (H x W), level 1 ((H+1)/2 x (W+1)/2), ... each 16-byte aligned
 */
int egl_gray_pyramid(const uint8_t *frames, int F, int H, int W, size_t row_stride, size_t frame_stride, int max_level,
                     uint8_t *pyr, void *stream);

/* The same for a strided subset of a clip: frame p is read at frames + p * frame_stride and its pyramid written at
 * pyr + p * pyr_stride (0 = egl_pyramid_bytes, i.e. dense).  With frame_stride = k frames and pyr_stride = k pyramids the
 * call builds the pyramids of one step of every chain (frames s, s + k, s + 2k, ...), which is the order the chain-parallel
 * propagation needs them in: the tracker of round s only waits for steps s and s + 1. */
int egl_gray_pyramid_strided(const uint8_t *frames, int F, int H, int W, size_t row_stride, size_t frame_stride, int max_level,
                             uint8_t *pyr, size_t pyr_stride, void *stream);

/*
 * cv2.calcOpticalFlowPyrLK(prev_gray, curr_gray, prev_points, None, winSize=(15,15), maxLevel,
 * criteria=(EPS|COUNT, max_count, eps)) (coordinate_model.py:431-435, lk_params :65), bit-exact with
 * OpenCV's SSE build (see oracle/optflow.py).  The points of pair p are the entries of keypoint set p.
 *   new_pts [n][64][2] float32, status [n][64] uint8 (1 = found), both indexed by CHANNEL (entries of
 *   channels that are not in the set are left untouched): egl_filter_flow may then be given a subset
 *   of the tracked set -- the inliers a fit selected while the tracker was already running.
 */
int egl_track_keypoints(const uint8_t *pyr, int H, int W, int max_level, const int32_t *kp_xy, const uint8_t *kp_order,
                        const int32_t *kp_count, int n, int prev0, int next0, int frame_step, int max_count, double eps,
                        float *new_pts, uint8_t *status, void *stream);

/*
 * The rest of calculate_optical_flow (coordinate_model.py:438-478): drop lost points, movers with a
 * z-score > 2 (float32 numpy statistics) and points whose mean 3x3 hue changed by more than 25, emit
 * the survivors as a new keypoint set (labels taken by position in the status-filtered list, :446).
 *   frames: BGR uint8; the hue is read from frame hue0 + p * frame_step (the `frame` argument of :419)
 *   prev_*: the set the reference would have flowed (its entries must have been tracked); new_pts / status by channel
 */
int egl_filter_flow(const uint8_t *frames, int H, int W, size_t row_stride, size_t frame_stride, int hue0, int frame_step,
                    const int32_t *prev_xy, const uint8_t *prev_order, const int32_t *prev_count, const float *new_pts,
                    const uint8_t *status, int n, int32_t *out_xy, uint8_t *out_order, int32_t *out_count,
                    uint8_t *out_src, void *stream);

/* a = {**a, **b} per frame (coordinate_model.py:311,320,322,324); apply[n] optional (0 = leave frame alone);
 * b_src NULL = EGL_KP_PY_INT. */
int egl_merge_keypoints(int32_t *a_xy, uint8_t *a_order, int32_t *a_count, uint8_t *a_src, const int32_t *b_xy,
                        const uint8_t *b_order, const int32_t *b_count, const uint8_t *b_src, const uint8_t *apply, int n,
                        void *stream);

/*
 * CoordinateModel.calibrate_keypoints (coordinate_model.py:520-555): a keypoint whose HSV value is
 * below 150 moves to the brightest pixel of frame[y-3:y+3, x-3:x+3].  err[n] is set to 1 for frames
 * on which the reference raises IndexError (:548, a dim keypoint in row or column 0).
 */
int egl_calibrate_keypoints(const uint8_t *frames, int H, int W, size_t row_stride, size_t frame_stride, int frame0,
                            int frame_step, int32_t *kp_xy, const uint8_t *kp_order, const int32_t *kp_count,
                            uint8_t *kp_src, int32_t *err, int n, void *stream);

/* egl_fit_homography restricted to the frames with sched[f] | retry[f] (retry may be NULL): the
 * cadence test of coordinate_model.py:333.  Other frames get status EGL_FIT_SKIPPED. */
int egl_fit_homography_masked(const int32_t *kp_xy, const uint8_t *kp_order, const int32_t *kp_count, int F, int mode,
                              int K, const uint8_t *hyp, uint64_t seed, double thr, double confidence, double *H,
                              uint64_t *used_mask, uint64_t *inlier_mask, int32_t *status, int32_t *info,
                              const uint8_t *sched, const uint8_t *retry, void *stream);

/*
 * What follows a fit attempt (coordinate_model.py:350-367): on success the keypoint set shrinks to
 * the inliers (they become the "Keypoints" of the frame and the seed of the next flow step) and
 * retry[f] clears; on failure retry[f] is set.  fit_ok[f] = 1 where this frame produced a new H.
 */
int egl_commit_fit(uint8_t *kp_order, int32_t *kp_count, uint8_t *kp_src, const int32_t *status,
                   const uint64_t *inlier_mask, const uint8_t *sched, uint8_t *retry, uint8_t *fit_ok, int n, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EAGLE_B200_H_ */
