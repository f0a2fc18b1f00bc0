/*
 * dr3lk.h -- C ABI of the B200-native pyramidal Lucas-Kanade path for kvmanohar22/3DR.
 *
 * The reference has no FFI layer; its boundary for this path is two C++ call surfaces:
 *   (1) cv::calcOpticalFlowPyrLK(prevImg, nextImg, prevPts, nextPts, status, err, winSize, maxLevel,
 *       criteria, flags, minEigThreshold)          -- called at reference src/initialization.cpp:608-613
 *   (2) utils::create_img_pyramid(img, n_levels, pyr) -- declared include/utils.hpp:67, defined
 *       src/utils.cpp:421-430, called from src/frame.cpp:18 (Frame constructor).
 * Every entry point below names the one it replaces.  Plain pointers and sizes only; all functions
 * return DR3LK_OK (0) or a negative DR3LK_E_* code and leave a message in dr3lk_last_error().
 * A context is bound to one CUDA device and one stream; calls on one context must not overlap
 * (use one context per host thread / per GPU).  There is no CPU fallback: without a usable CUDA
 * device dr3lk_create fails with DR3LK_E_CUDA.
 */
#ifndef DR3LK_H_
#define DR3LK_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DR3LK_OK 0
#define DR3LK_E_ARG (-1)         /* cv::Exception -215 (CV_Assert) in the reference path            */
#define DR3LK_E_SIZE (-2)        /* image sizes / types differ, or a shape the path cannot represent */
#define DR3LK_E_CUDA (-3)        /* CUDA runtime error, no device, out of memory                     */
#define DR3LK_E_UNSUPPORTED (-4) /* shapes on which the reference itself overruns its buffers        */

/* cv::TermCriteria::Type */
#define DR3LK_TERM_COUNT 1
#define DR3LK_TERM_EPS 2
/* cv::OPTFLOW_* flags */
#define DR3LK_USE_INITIAL_FLOW 4
#define DR3LK_GET_MIN_EIGENVALS 8

/* rounding / walk of utils::reduce_to_half (reference src/utils.cpp:382-419) */
#define DR3LK_BOX_AUTO_X86 0 /* as the reference behaves on x86: SSE2 path iff cols % 16 == 0 and 16-B aligned data */
#define DR3LK_BOX_TRUNC 1    /* scalar / NEON arithmetic (a+b+c+d)/4, src/utils.cpp:411 and 353-372                  */
#define DR3LK_BOX_SSE2 2     /* halfSampleSSE2 double round-up averaging, src/utils.cpp:324-350                      */

#define DR3LK_MAX_LEVELS 16

typedef struct dr3lk_ctx dr3lk_ctx;

/* ---- context ------------------------------------------------------------------------------------ */
int dr3lk_create(dr3lk_ctx** out, int device);
void dr3lk_destroy(dr3lk_ctx* ctx);
const char* dr3lk_last_error(const dr3lk_ctx* ctx); /* ctx may be NULL: message of the last failed create */
/* Run on an existing cudaStream_t (e.g. torch's current stream); NULL restores the context's own stream.  Switching to a
 * different stream first waits for the work queued on the current one (the context's scratch buffers are ordered by the
 * stream they were last used on). */
int dr3lk_set_stream(dr3lk_ctx* ctx, void* cuda_stream);
int dr3lk_synchronize(dr3lk_ctx* ctx);
/* Checked build only (`make -C 3dr_b200/csrc checked` -> lib/libdr3lk_checked.so, -DDR3LK_CHECKED): the specialised LK kernels
 * verify every staged rectangle against the allocation it is copied from, every shared-memory load against its region and
 * the staged rows, every output index against the batch, and the pyramid kernels every store (level pixels, mirrored apron
 * pixels, derivatives) against the destination image (compute-sanitizer is not available on the GPU pool).  Returns
 * out4 = {violations, kind of the first, its detail, checks executed} since the last read and resets them;
 * DR3LK_E_UNSUPPORTED in the default build, which contains none of the checks. */
int dr3lk_debug_check_read(dr3lk_ctx* ctx, unsigned long long* out4);
/* Parity hook for the corner detector: arc length of the FAST segment test used by dr3lk_fast_detect / dr3lk_init_first_frame
 * on this context.  10 (the default) is the reference's fast_corner_detect_10 (src/features.cpp:55-66); 9 runs the same
 * kernels as OpenCV's FAST-9, which is how the tests pin them against cv2 (the `fast` library the reference links is not
 * available to test against).  Not meant for production use. */
int dr3lk_debug_set_fast_arc(dr3lk_ctx* ctx, int arc);
/* Number of CUDA kernels this context has launched since creation (bench.py's gpu_launches). */
uint64_t dr3lk_launch_count(const dr3lk_ctx* ctx);
/* Kernel timing with CUDA events on the launching stream (bench.py's roofline leg).  While profiling is on, every
 * LK launch and every pyramid build is bracketed by events; dr3lk_profile_read waits for them, returns the summed
 * durations (milliseconds) and launch counts since the last read, and clears them. */
int dr3lk_set_profiling(dr3lk_ctx* ctx, int on);
int dr3lk_profile_read(dr3lk_ctx* ctx, float* lk_ms, int* lk_launches, float* pyramid_ms, int* pyramid_builds);
/* Pinned host memory for callers that want the host-buffer entry points to overlap copies with compute.  The single-pair
 * entry points (dr3lk_calc_optical_flow_pyr_lk, dr3lk_pyramid_create, dr3lk_track_frame) upload an image as it is -- one
 * contiguous copy, no packing into the context's own pinned mirror -- when its rows are continuous (step == w, the usual
 * cv::Mat) or sit at the aligned pitch (w + 15) / 16 * 16 (for the two-image call: both frames at the same step).  From
 * page-locked memory (this allocator, dr3lk_host_register, cudaHostAlloc, cudaHostRegister) that copy is an asynchronous DMA:
 * 85 us per call on a KITTI 1241x376 pair from C++; from pageable memory the driver stages it: 114 us.  Rows at any other
 * step (ROIs) are packed first (120 us); results are identical in all three cases. */
void* dr3lk_host_alloc(size_t bytes);
void dr3lk_host_free(void* p);
/* Page-locks memory the caller already owns (cudaHostRegister / cudaHostUnregister), e.g. the data of an existing continuous
 * cv::Mat or a camera ring buffer: register once when the buffer is created (it costs a fraction of a millisecond), unregister
 * before freeing it.  Registered images take the same direct path as dr3lk_host_alloc memory. */
int dr3lk_host_register(void* p, size_t bytes);
int dr3lk_host_unregister(void* p);

/* ---- (2) box pyramid: utils::create_img_pyramid, reference src/utils.cpp:421-430 ------------------ */
/* Host buffers.  img: h rows of `step` bytes, CV_8UC1.  out_levels[l-1] receives level l (l = 1..n_levels-1),
 * continuous (w>>l) x (h>>l); level 0 stays the caller's image, as in the reference (shallow copy).
 * mode AUTO_X86 also looks at the 16-byte alignment of img, like reduce_to_half does. */
int dr3lk_box_pyramid(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels,
                      uint8_t* const* out_levels, int mode);
/* Device buffers, `batch` images `image_stride` bytes apart.  out_levels_dev[l-1] is a device buffer of
 * batch * (w>>l)*(h>>l) bytes (images back to back, continuous). */
int dr3lk_box_pyramid_device(dr3lk_ctx* ctx, const uint8_t* img_dev, int w, int h, size_t pitch, size_t image_stride,
                             int batch, int n_levels, uint8_t* const* out_levels_dev, int mode);

/* ---- (1) cv::calcOpticalFlowPyrLK, reference call site src/initialization.cpp:608-613 ------------- */
/* One frame pair, host buffers, OpenCV semantics:
 *   prev/next  CV_8UC1 w x h, row steps in bytes;  prev_pts n x 2 float (x, y);
 *   next_pts   n x 2 float, output; also input (initial guess) when flags has DR3LK_USE_INITIAL_FLOW;
 *   status     n bytes, 1 = tracked;  err n floats or NULL (like an unneeded OutputArray);
 *   win/max_level/criteria/flags/min_eig_threshold as in OpenCV (maxCount clamped to [0,100], eps to [0,10]).
 * n == 0 returns DR3LK_OK without touching the outputs.  Errors mirror OpenCV's CV_Assert conditions
 * (max_level < 0, win <= 2 -> DR3LK_E_ARG).  err of a lost point is written as 0 (OpenCV leaves it
 * uninitialised).  Synchronous: outputs are complete on return. */
int dr3lk_calc_optical_flow_pyr_lk(dr3lk_ctx* ctx, const uint8_t* prev, size_t prev_step, const uint8_t* next,
                                   size_t next_step, int w, int h, const float* prev_pts, float* next_pts,
                                   uint8_t* status, float* err, int n, int win_w, int win_h, int max_level,
                                   int crit_type, int crit_max_count, double crit_eps, int flags,
                                   double min_eig_threshold);

/* Batched form, device buffers: `batch` independent frame pairs of identical size.  Pair b's images are at
 * prev_dev + b*image_stride / next_dev + b*image_stride (rows `pitch` bytes apart).  Its points are
 * prev_pts_dev[pts_offset[b] .. pts_offset[b+1]) (pts_offset: HOST array of batch+1 ints, pts_offset[0] == 0).
 * stats_dev (optional, may be NULL): one uint32 per point -- bits 0..15 LK iterations executed over all levels,
 * bits 16..23 number of levels whose template window was built, bit 24 set when the final error pass ran
 * (the inputs of the algorithmic-bytes model).  Asynchronous on the context's stream. */
int dr3lk_track_batch(dr3lk_ctx* ctx, const uint8_t* prev_dev, const uint8_t* next_dev, int w, int h, size_t pitch,
                      size_t image_stride, int batch, const float* prev_pts_dev, float* next_pts_dev,
                      uint8_t* status_dev, float* err_dev, const int* pts_offset, uint32_t* stats_dev, int win_w,
                      int win_h, int max_level, int crit_type, int crit_max_count, double crit_eps, int flags,
                      double min_eig_threshold);

/* Batched form, HOST buffers (the end-to-end path): same layout as dr3lk_track_batch but every pointer is host
 * memory (pinned memory from dr3lk_host_alloc lets copies overlap compute).  The batch is cut into chunks of
 * `chunk_pairs` pairs (0 = choose) that are pipelined H2D -> pyramids -> LK -> D2H over several streams.
 * Synchronous: outputs are complete on return. */
int dr3lk_track_batch_host(dr3lk_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t step,
                           size_t image_stride, int batch, const float* prev_pts, float* next_pts, uint8_t* status,
                           float* err, const int* pts_offset, uint32_t* stats, int chunk_pairs, int win_w, int win_h,
                           int max_level, int crit_type, int crit_max_count, double crit_eps, int flags,
                           double min_eig_threshold);

/* ---- multi-GPU form of the host-buffer batch ---------------------------------------------------------- */
/* The reference is one C++ process (src/handler.cpp:31-48); to use several GPUs from one process a dr3lk_multi owns one
 * context per listed device (a device may be listed more than once) and dr3lk_multi_track_batch_host runs one worker
 * thread per context.  Frame pairs are independent, so the batch is cut into contiguous blocks (dr3lk_shard_range: pair p
 * belongs to rank floor(p * world / n_pairs)); every device builds the pyramids of its own pairs and copies its results
 * straight into its slice of the caller's arrays: no collective, no exchange between devices.  Arguments and results are
 * those of dr3lk_track_batch_host (bit-identical to a single-device call, whatever the number of devices). */
typedef struct dr3lk_multi dr3lk_multi;
int dr3lk_multi_create(dr3lk_multi** out, const int* devices, int n_devices);
void dr3lk_multi_destroy(dr3lk_multi* m);
int dr3lk_multi_size(const dr3lk_multi* m);
dr3lk_ctx* dr3lk_multi_context(dr3lk_multi* m, int i); /* context of rank i (owned by m), e.g. for dr3lk_launch_count */
const char* dr3lk_multi_last_error(const dr3lk_multi* m);
void dr3lk_shard_range(int n_pairs, int rank, int world, int* lo, int* hi);
int dr3lk_multi_track_batch_host(dr3lk_multi* m, const uint8_t* prev, const uint8_t* next, int w, int h, size_t step,
                                 size_t image_stride, int batch, const float* prev_pts, float* next_pts, uint8_t* status,
                                 float* err, const int* pts_offset, uint32_t* stats, int chunk_pairs, int win_w, int win_h,
                                 int max_level, int crit_type, int crit_max_count, double crit_eps, int flags,
                                 double min_eig_threshold);

/* ---- the pyramids calcOpticalFlowPyrLK builds internally (buildOpticalFlowPyramid + calcScharrDeriv) ---- */
/* Level sizes with OpenCV's early stop; ws/hs need max_level+1 entries.  Returns the effective maxLevel. */
int dr3lk_lk_level_sizes(int w, int h, int win_w, int win_h, int max_level, int* ws, int* hs);
/* Host in / host out, for parity checks: out_levels[l] (l = 0..eff. maxLevel) continuous w_l x h_l uint8;
 * out_derivs[l] continuous w_l x h_l x 2 int16 (Ix, Iy interleaved) or out_derivs == NULL.
 * *eff_max_level receives the effective maxLevel. */
int dr3lk_build_lk_pyramid(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int win_w, int win_h,
                           int max_level, uint8_t* const* out_levels, int16_t* const* out_derivs, int* eff_max_level);

/* ---- SURVEY.md 8(f) "next" rows ---------------------------------------------------------------------------- */

/* f-2: pyramid caching.  The reference passes only level-0 images to cv::calcOpticalFlowPyrLK
 * (src/initialization.cpp:608-609), so OpenCV rebuilds both Gaussian pyramids and the reference frame's Scharr
 * derivatives on every call -- also on every retry against the same reference frame (src/handler.cpp:67-72).
 * A dr3lk_pyramid is the device-resident equivalent of OpenCV's precomputed-pyramid input form: Gaussian levels
 * (with OpenCV's early stop for win/max_level) plus Scharr derivatives, built once per frame.
 * Lifetime: a pyramid belongs to the context that created it.  dr3lk_destroy(ctx) releases the device memory of every
 * pyramid of that context that is still alive and orphans it: the handle stays valid for exactly one thing,
 * dr3lk_pyramid_destroy (any order of destruction is legal, e.g. from a garbage collector or a static destructor);
 * every other call that is handed an orphaned pyramid fails with DR3LK_E_ARG. */
typedef struct dr3lk_pyramid dr3lk_pyramid;
int dr3lk_pyramid_create(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int win_w, int win_h, int max_level,
                         dr3lk_pyramid** out);
void dr3lk_pyramid_destroy(dr3lk_pyramid* pyr);
int dr3lk_pyramid_levels(const dr3lk_pyramid* pyr); /* effective maxLevel + 1 */
/* Same semantics and results as dr3lk_calc_optical_flow_pyr_lk on the two source images; win must be the one the
 * pyramids were built for and max_level is clamped to what they hold. */
int dr3lk_calc_optical_flow_pyr_lk_cached(dr3lk_ctx* ctx, const dr3lk_pyramid* prev, const dr3lk_pyramid* next,
                                          const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n, int win_w,
                                          int win_h, int max_level, int crit_type, int crit_max_count, double crit_eps, int flags,
                                          double min_eig_threshold);

/* Streaming form of the same row, for the per-frame loop of the reference (src/handler.cpp:31-48: one new image per call,
 * tracked against a frame that is already on the device -- the anchored reference frame of Init::process_second_frame or
 * simply the previous frame of a chain).  ONE call = one H2D copy (new image + points), the new frame's pyramid, LK, one
 * D2H copy, one synchronisation; results are identical to dr3lk_calc_optical_flow_pyr_lk on the two source images.
 * keep_next: 0 = discard the new frame's pyramid; 1 = return it in *next_out with Gaussian levels only (it can be the
 * `next` side of later calls); 2 = return it with derivatives (it is the `prev` of the following call). */
int dr3lk_track_frame(dr3lk_ctx* ctx, const dr3lk_pyramid* prev, const uint8_t* next_img, size_t next_step,
                      const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n, int win_w, int win_h,
                      int max_level, int crit_type, int crit_max_count, double crit_eps, int flags, double min_eig_threshold,
                      int keep_next, dr3lk_pyramid** next_out);

/* f-3: the step right after the LK call, reference src/initialization.cpp:615-635 -- drop the points with status == 0
 * (order preserved, like the erase loop), disparity = ||ref - cur|| (double), and the unit bearing vector of the current
 * point, Pinhole::cam2world (src/camera.cpp:25-41): ((u-cx)/fx, (v-cy)/fy, 1) normalised for an undistorted camera; with
 * distortion (d0..d4 = k1 k2 p1 p2 k3 in `distortion`, active when fabs(d0) > 1e-7 like Pinhole::_distortion,
 * src/camera.cpp:17) the pixel first goes through cv::undistortPoints' five fixed-point iterations with the float K / D the
 * reference's constructor builds.  distortion may be NULL (= undistorted).
 * Host buffers; out_ref/out_cur hold n x 2 floats, out_disparity n doubles, out_bearing n x 3 doubles (may be NULL).
 * *n_kept receives the number of surviving points. */
int dr3lk_filter_tracks(dr3lk_ctx* ctx, const float* ref_pts, const float* cur_pts, const uint8_t* status, int n, double fx,
                        double fy, double cx, double cy, const double* distortion, float* out_ref, float* out_cur,
                        double* out_disparity, double* out_bearing, int* n_kept);

/* f-1 / a-10: the prevPts provider -- feature_detection::FastDetector::detect (reference src/features.cpp:43-98) on
 * the Frame's box pyramid (utils::create_img_pyramid, n_levels levels, rounding `box_mode`): FAST-10 corners
 * (threshold fast_threshold, 20 in the reference), fast_corner_score_10, fast_nonmax_3x3, then per grid cell of
 * cell_size pixels the corner with the best utils::shi_tomasi_score above detection_threshold
 * (Config::cell_size() = 30, Config::min_harris_corner_score() = 20.0, src/config.cpp:9-12).
 * occupancy: ceil(w/cell) * ceil(h/cell) bytes, non-zero = cell already taken (NULL = all free).
 * Outputs, in grid-cell order like the reference's feature list: out_xy level-0 pixel coordinates (2 ints each),
 * out_level, out_score; capacity ceil(w/cell)*ceil(h/cell) entries each.  Host buffers, synchronous. */
int dr3lk_fast_detect(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, int cell_size,
                      int fast_threshold, double detection_threshold, int box_mode, const uint8_t* occupancy, int* out_xy,
                      int* out_level, float* out_score, int* n_out);

/* f-4: scoring of RANSAC fundamental-matrix hypotheses -- InitHelper::CheckFundamental (reference
 * src/initialization.cpp:171-249) for n_hyp matrices at once, as FindFundamental (src/initialization.cpp:81-133)
 * needs it for its 200 hypotheses.  F21: n_hyp x 9 floats (row-major 3x3, already de-normalised: T2^T * Fn * T1);
 * pts1 / pts2: n x 2 floats (matched key points of frames 1 and 2); sigma as in the reference (1.0).
 * out_scores: n_hyp floats; out_inliers: n_hyp x n bytes (may be NULL); *best receives the index FindFundamental would
 * keep (first hypothesis with the strictly largest score, -1 when every score is <= 0).  The fp32 operations and the
 * accumulation order over the matches are the reference's, so scores are bit-identical to a scalar evaluation. */
int dr3lk_score_fundamental(dr3lk_ctx* ctx, const float* F21, int n_hyp, const float* pts1, const float* pts2, int n, float sigma,
                            float* out_scores, uint8_t* out_inliers, int* best);

/* ---- the two-frame initialiser front end in three calls, device-resident between its steps ------------------------
 * reference: init::Init::process_first_frame / process_second_frame, src/initialization.cpp:546-661.  The four f-rows above
 * as separate host-buffer calls cost four uploads / downloads / synchronisations per frame pair; these three entry points
 * chain them on the device and return bit-identical results.
 *
 * dr3lk_init_first_frame (src/initialization.cpp:546-585 + src/frame.cpp:13-20): ONE upload of the image; the Frame's box
 * pyramid and FastDetector::detect on it (arguments and outputs of dr3lk_fast_detect) AND, from the same device copy, the LK
 * pyramid of the frame for the tracking calls that follow (as dr3lk_pyramid_create with win / max_level); ONE download.
 *
 * dr3lk_init_second_frame (src/initialization.cpp:587-657): ONE upload (new image, kps_ref, kps_cur), the new frame's Gaussian
 * pyramid, cv::calcOpticalFlowPyrLK against the reference pyramid (arguments of dr3lk_track_frame), then the erase-by-status
 * loop, disparities and cam2world bearings of lines 615-635 (arguments and outputs of dr3lk_filter_tracks); ONE download.
 * out_status / out_err (n entries each, may be NULL) are the raw LK outputs.  The compacted (ref, cur) points stay on the device.
 *
 * dr3lk_init_score_fundamental (src/initialization.cpp:81-133, 171-249): scores hypotheses against the points the last
 * dr3lk_init_second_frame left on the device -- only the matrices go up, the scores (and inlier masks, n_hyp x n_kept bytes or
 * NULL) come back.  Results as dr3lk_score_fundamental on the compacted points. */
int dr3lk_init_first_frame(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, int cell_size,
                           int fast_threshold, double detection_threshold, int box_mode, const uint8_t* occupancy, int win_w,
                           int win_h, int max_level, int* out_xy, int* out_level, float* out_score, int* n_out,
                           dr3lk_pyramid** ref_pyramid);
int dr3lk_init_second_frame(dr3lk_ctx* ctx, const dr3lk_pyramid* ref, const uint8_t* cur_img, size_t cur_step, const float* kps_ref,
                            const float* kps_cur, int n, int win_w, int win_h, int max_level, int crit_type, int crit_max_count,
                            double crit_eps, int flags, double min_eig_threshold, double fx, double fy, double cx, double cy,
                            const double* distortion, float* out_ref, float* out_cur, double* out_disparity, double* out_bearing,
                            uint8_t* out_status, float* out_err, int* n_kept);
int dr3lk_init_score_fundamental(dr3lk_ctx* ctx, const float* F21, int n_hyp, float sigma, float* out_scores, uint8_t* out_inliers,
                                 int* best);

#ifdef __cplusplus
}
#endif
#endif /* DR3LK_H_ */
