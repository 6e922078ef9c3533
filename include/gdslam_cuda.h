/*
 * gdslam_cuda.h — C ABI of the B200-native GD-SLAM front-end (GeoMaskMaker + ORBextractor hot path).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch/OpenCV types.  A thin C++ shim
 * with the reference's unchanged class declarations (gd-slam_b200/host/) forwards to these entry points,
 * so Tracking.cc / rgbd_tum.cc of the reference compile and run untouched.
 *
 * Reference interface each group replaces (paths relative to the GD-SLAM tree):
 *   gd_geomask_create   <- GeoMaskMaker::GeoMaskMaker(Mat K, Mat DistCoef, float DepthMapFactor)
 *                          include/GeoMaskMaker.h:94, src/GeoMaskMaker.cc:39-70   (called src/Tracking.cc:137)
 *   gd_geomask_push     <- GeoMaskMaker::AddNewImage(Mat rgb, Mat depth, Mat label, Mat originlabel)
 *                          include/GeoMaskMaker.h:96, src/GeoMaskMaker.cc:409-429 (called src/Tracking.cc:242)
 *   gd_geomask_mask     <- GeoMaskMaker::GetNoGMMmask(Mat& mask)
 *                          include/GeoMaskMaker.h:98, src/GeoMaskMaker.cc:167-408 (called src/Tracking.cc:245)
 *                          The pose (R,T) that GetRt() (:77-156) estimates on the CPU is an INPUT here.
 *   gd_orb_create       <- ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 *                          include/ORBextractor.h:51, src/ORBextractor.cc:410-470 (called src/Tracking.cc:108)
 *   gd_orb_extract      <- ORBextractor::operator()(InputArray image, InputArray mask, vector<KeyPoint>&, OutputArray)
 *                          include/ORBextractor.h:59, src/ORBextractor.cc:1043-1105 (called src/Frame.cc:419-425)
 *   gd_frontend_*       <- the per-frame sequence of Tracking::GrabImageRGBD_GD, src/Tracking.cc:212-252
 *                          (cvtColor, Frame()->ORB, AddNewImage, GetNoGMMmask) for `batch` independent
 *                          RGB-D streams stepped together on one GPU; this is what the throughput benchmark drives.
 *   gd_stage_*          <- single stages of the above with host buffers in/out, used by the parity tests to check
 *                          every kernel against the CPU oracle through this ABI.
 *
 * Conventions
 *   - every function returns 0 on success, a negative GD_E* code otherwise, and never throws; the message of the
 *     last failure on the calling thread is available from gd_last_error().
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails with GD_ENODEVICE.
 *   - images are row-major with a byte stride ("step"); BGR is 8UC3 in imread order; depth is 32FC1 in metres
 *     (already multiplied by 1/DepthMapFactor, src/Tracking.cc:234-235); masks are 8UC1 with 1 = static,
 *     0 = dynamic, exactly the matrix GetNoGMMmask hands back.
 *   - the library owns all device memory; the caller owns every host buffer; no pointer is retained across calls.
 *   - a handle is bound to one device and one CUDA stream; calls on one handle must be ordered by the caller
 *     (the reference calls them from the single tracking thread); distinct handles are independent.
 *   - `batch` streams in one handle advance in lockstep: array arguments carry one entry per stream.
 */
#ifndef GDSLAM_CUDA_H_
#define GDSLAM_CUDA_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GD_OK 0
#define GD_EINVAL (-1)    /* bad argument */
#define GD_ENODEVICE (-2) /* no CUDA device / device index out of range */
#define GD_ECUDA (-3)     /* a CUDA runtime call or kernel failed */
#define GD_ENOMEM (-4)
#define GD_ECAPACITY (-5) /* an output buffer given by the caller is too small (results clamped, counts set) */
#define GD_EINTERNAL (-6) /* an internal device-side capacity bound was exceeded (results truncated; reported once) */

#define GD_ABI_VERSION 2

#if defined(__GNUC__)
#define GD_API __attribute__((visibility("default")))
#else
#define GD_API
#endif

/* mirrors cv::KeyPoint field order (pt.x, pt.y, size, angle, response, octave, class_id): 28 bytes */
typedef struct gd_keypoint {
    float x, y;
    float size;
    float angle;
    float response;
    int32_t octave;
    int32_t class_id;
} gd_keypoint;

typedef struct gd_geomask gd_geomask_t;
typedef struct gd_orb gd_orb_t;
typedef struct gd_frontend gd_frontend_t;

GD_API const char* gd_last_error(void);
GD_API int gd_abi_version(void);
GD_API int gd_device_count(int* count);
/* name (<= len bytes), SM count and total global memory of a device */
GD_API int gd_device_info(int device, char* name, int len, int* sm_count, size_t* total_mem);

/* page-locked host memory for callers that want true asynchronous H2D/D2H (bench e2e leg) */
GD_API int gd_host_alloc(void** ptr, size_t bytes);
GD_API int gd_host_free(void* ptr);
/* measurement aid: rate of `iters` back-to-back pinned cudaMemcpyAsync of `bytes` (to_device != 0: host->device) on `device`,
 * CUDA-event timed.  Run on all GPUs of a box at once it gives the host-side ceiling of the end-to-end (host buffer) path. */
GD_API int gd_probe_copy(int device, size_t bytes, int iters, int to_device, double* gb_per_s);

/* ------------------------------------------------------------------ GeoMaskMaker ------------------ */
/* K: 3x3 row-major; dist: k1,k2,p1,p2[,k3] or NULL (ndist 0).  Non-zero distortion builds the undistorted-pixel
 * LUT of GeoMaskMaker.cc:56-69 on the host (iterative undistortPoints).  width/height replace the literals
 * 640/480 of GeoMaskMaker.cc:54-55. */
GD_API int gd_geomask_create(gd_geomask_t** out, const float K[9], const float* dist, int ndist, float depth_factor,
                      int width, int height, int device, int batch);
GD_API void gd_geomask_destroy(gd_geomask_t* h);
/* AddNewImage: one new frame per stream.  bgr[b], depth_m[b]: host pointers. */
GD_API int gd_geomask_push(gd_geomask_t* h, const uint8_t* const* bgr, size_t bgr_step, const float* const* depth_m,
                    size_t depth_step);
/* GetNoGMMmask: R (batch x 9), T (batch x 3): pose with P_cur = R * P_ref + T between the buffered pair (t-5, t);
 * pose_valid[b] == 0 reproduces GetRt()'s failure path (all-ones mask, GeoMaskMaker.cc:179-185); fewer than six
 * pushed frames reproduce the warm-up path (:171-175).  Blocks until mask_out[b] (H x W, values {0,1}) is written. */
GD_API int gd_geomask_mask(gd_geomask_t* h, const float* R, const float* T, const int* pose_valid, uint8_t* const* mask_out,
                    size_t mask_step);
/* number of frames pushed so far (image_count analogue) */
GD_API int gd_geomask_frames(const gd_geomask_t* h);
/* GeoMaskMaker::GetRt (GeoMaskMaker.cc:77-141) for the buffered pair (t-5, t): after gd_geomask_enable_getrt every pushed frame
 * also gets its cv::ORB features (cached per ring slot); gd_geomask_getrt_points matches the pair and returns, per stream, the
 * objectPoints (up to 100 x 3 floats) and imagePixels (100 x 2) of solvePnPRansac (:148) and their number.  The frames of the
 * pair must have been pushed after the stage was enabled. */
GD_API int gd_geomask_enable_getrt(gd_geomask_t* h);
GD_API int gd_geomask_getrt_points(gd_geomask_t* h, float* const* object_points, float* const* image_pixels, int* n_points);

/* intermediate products of the LAST gd_geomask_mask call of stream `stream`, for parity tests */
enum gd_debug_what {
    GD_DBG_FLOW = 0,     /* f32  H*W*2 (x,y interleaved)  = GetFlow()          */
    GD_DBG_DIST = 1,     /* f32  H*W   dist_image before normalize            */
    GD_DBG_EDGE_REF = 2, /* u8   H*W   GetEdge(_firstDepth)                   */
    GD_DBG_EDGE_CUR = 3, /* u8   H*W   GetEdge(_secondDepth)                  */
    GD_DBG_GRAY_CUR = 4, /* u8   H*W   BGR2GRAY of the newest frame           */
    GD_DBG_MINMAX = 5,   /* f32  2     min, max of dist_image                 */
    GD_DBG_LUT = 6       /* f32  H*W*2 undistortedPoint LUT of the ctor (GeoMaskMaker.cc:56-69); error when D == 0 */
};
GD_API int gd_geomask_debug_fetch(gd_geomask_t* h, int what, int stream, void* dst, size_t dst_bytes);

/* ------------------------------------------------------------------ ORBextractor ------------------ */
GD_API int gd_orb_create(gd_orb_t** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                  int max_width, int max_height, int device, int batch);
GD_API void gd_orb_destroy(gd_orb_t* h);
/* operator(): gray[b] 8UC1 host images of size w x h.  kps[b] (capacity entries) and desc[b] (capacity x 32 bytes)
 * receive the keypoints level by level in the reference's order; n_out[b] the count (can exceed nfeatures by a few,
 * SURVEY B-9).  GD_ECAPACITY if capacity is too small (n_out still set). */
GD_API int gd_orb_extract(gd_orb_t* h, const uint8_t* const* gray, size_t gray_step, int w, int h_, gd_keypoint* const* kps,
                   uint8_t* const* desc, int capacity, int* n_out);
/* pyramid level of the last extraction (mvImagePyramid analogue, include/ORBextractor.h:85) */
GD_API int gd_orb_fetch_level(gd_orb_t* h, int stream, int level, uint8_t* dst, size_t dst_step, int* w, int* h_);
GD_API int gd_orb_level_size(const gd_orb_t* h, int level, int* w, int* h_);
GD_API int gd_orb_features_per_level(const gd_orb_t* h, int* n_per_level /* nlevels ints */);

/* ------------------------------------------------------------------ batched front-end -------------- */
typedef struct gd_frontend_config {
    float K[9];
    float dist[5];
    int ndist;
    float depth_factor;
    int width, height;
    int device;
    int batch; /* independent RGB-D streams stepped together */
    /* ORB settings (TUM3.yaml:41-54): 1500, 1.2, 8, 20, 7 */
    int nfeatures;
    float scale_factor;
    int nlevels;
    int ini_th_fast, min_th_fast;
    int orb_gray_order; /* 1: RGB2GRAY on the BGR bytes (Camera.RGB=1, Tracking.cc:219-225); 0: BGR2GRAY */
    int kp_capacity;    /* per-stream keypoint capacity of the result buffers (>= nfeatures + 3*nlevels) */
    int staged_slots;   /* number of device-resident input slots for gd_frontend_stage (0 = none) */
    int getrt;          /* 1: run GeoMaskMaker::GetRt's feature / matching half (GeoMaskMaker.cc:77-141) as a stage of every step */
} gd_frontend_config;

GD_API int gd_frontend_create(gd_frontend_t** out, const gd_frontend_config* cfg);
GD_API void gd_frontend_destroy(gd_frontend_t* h);
/* one frame per stream, host buffers in, host buffers out (H2D and D2H inside the call):
 * ORB on the new frame, AddNewImage, GetNoGMMmask.  Any of mask_out / kps / desc / n_kp may be NULL (skipped). */
GD_API int gd_frontend_step(gd_frontend_t* h, const uint8_t* const* bgr, size_t bgr_step, const float* const* depth_m,
                     size_t depth_step, const float* R, const float* T, const int* pose_valid,
                     uint8_t* const* mask_out, size_t mask_step, gd_keypoint* const* kps, uint8_t* const* desc,
                     int* n_kp);
/* SURVEY section 8 row (f)-4, depth ingest: same as gd_frontend_step but the depth arrives as the raw 16-bit TUM image
 * and is converted on the device exactly like Tracking.cc:234-235 (imDepth.convertTo(CV_32F, 1/DepthMapFactor)):
 * metres = (float)v * (1.0f / depth_factor).  Halves the depth H2D bytes. */
GD_API int gd_frontend_step_u16(gd_frontend_t* h, const uint8_t* const* bgr, size_t bgr_step, const uint16_t* const* depth_raw,
                                size_t depth_step, const float* R, const float* T, const int* pose_valid,
                                uint8_t* const* mask_out, size_t mask_step, gd_keypoint* const* kps, uint8_t* const* desc,
                                int* n_kp);
/* device-resident variant: upload frames into slot `slot` once, then step from HBM (no PCIe in the step);
 * results stay on the device until gd_frontend_fetch. */
GD_API int gd_frontend_stage(gd_frontend_t* h, int slot, const uint8_t* const* bgr, size_t bgr_step,
                      const float* const* depth_m, size_t depth_step);
GD_API int gd_frontend_step_staged(gd_frontend_t* h, int slot, const float* R, const float* T, const int* pose_valid);
GD_API int gd_frontend_fetch(gd_frontend_t* h, uint8_t* const* mask_out, size_t mask_step, gd_keypoint* const* kps,
                      uint8_t* const* desc, int* n_kp);
/* SURVEY section 8 row (f)-2, the step right after both kernels — Frame::Frame's mask erosion + keypoint filter
 * (src/Frame.cc:258-282) for the second Frame() of GrabImageRGBD_GD (src/Tracking.cc:252): erode(new mask, 31x31 ellipse),
 * keep keypoint i iff eroded((int)pt.y,(int)pt.x) == 1, order preserved.  Uses the mask and keypoints of the last step. */
GD_API int gd_frontend_fetch_filtered(gd_frontend_t* h, gd_keypoint* const* kps, uint8_t* const* desc, int* n_kp);
/* SURVEY section 8 row (f)-3 — Frame::UndistortKeyPoints + ComputeImageBounds (src/Frame.cc:576-636), ComputeStereoFromRGBD
 * (:815-837) and AssignFeaturesToGrid / PosInGrid (:402-417, :553-565) for the FILTERED keypoints of the last
 * gd_frontend_fetch_filtered call.  The camera is the handle's (K, dist); like the reference, it counts as distorted iff
 * dist[0] != 0 (then mvKeysUn = cv::undistortPoints of the keypoints and the grid bounds are the undistorted image corners;
 * otherwise mvKeysUn == mvKeys and the bounds are the image).  depth[b][i] / uright[b][i]: mvDepth / mvuRight (-1 when d <= 0;
 * the depth is read at the distorted keypoint, uRight uses the undistorted x), bf = Camera.bf.  Grid: cell = col * 48 + row
 * (mGrid[col][row]); cell_start[b] has 64*48+1 entries, cell_items[b] lists the keypoint indices of each cell in increasing
 * order.  keys_un[b] (the _un variant, optional): mvKeysUn positions, 2 floats per keypoint.  All per-keypoint output arrays
 * must hold kp_capacity entries. */
GD_API int gd_frontend_fetch_stereo_grid(gd_frontend_t* h, float bf, float* const* depth, float* const* uright,
                                         int* const* cell_start, int* const* cell_items);
GD_API int gd_frontend_fetch_stereo_grid_un(gd_frontend_t* h, float bf, float* const* depth, float* const* uright,
                                            int* const* cell_start, int* const* cell_items, float* const* keys_un);
/* SURVEY section 8 row (f)-1 — GeoMaskMaker::GetRt (src/GeoMaskMaker.cc:77-156) as a resident stage (config.getrt = 1): every
 * step computes cv::ORB(2000, 1.2, 8, 31, 0, 2) features of the new frame once (kept per ring slot), matches them against the
 * frame five steps back (BFMatcher NORM_HAMMING, crossCheck), sorts, keeps the first 100, undistorts, looks the depth up and
 * back-projects: exactly the objectPoints / imagePixels the reference hands to cv::solvePnPRansac (:148).  That call and
 * cv::Rodrigues stay with the caller (GD-SLAM links OpenCV):
 *   - gd_frontend_fetch_getrt returns the points of the last step (up to 100 x 3 / 100 x 2 floats per stream; n_points[b] = 0
 *     during the warm-up); GetRt returns false when n_points < 20 (:143-146);
 *   - with a pose hook installed, a step called with R == NULL and T == NULL asks the hook for the pose of every stream with
 *     at least 20 points (return non-zero when R (row-major 3x3) and T were written; zero = "no pose" = all-ones mask). */
typedef int (*gd_pose_hook_fn)(void* user, int stream, const float* object_points, const float* image_pixels, int n_points,
                               float R[9], float T[3]);
GD_API int gd_frontend_set_pose_hook(gd_frontend_t* h, gd_pose_hook_fn hook, void* user);
GD_API int gd_frontend_fetch_getrt(gd_frontend_t* h, float* const* object_points, float* const* image_pixels, int* n_points);
GD_API int gd_frontend_sync(gd_frontend_t* h);
/* CUDA-event timing on the handle's own stream: begin records an event, end records + synchronises and returns ms */
GD_API int gd_frontend_timer_begin(gd_frontend_t* h);
GD_API int gd_frontend_timer_end(gd_frontend_t* h, float* ms);
/* number of this library's kernels launched on the handle since creation, and a per-kernel-family event profile
 * (name[i] is a static string; ms[i] accumulated device time of family i over the profiled steps) */
GD_API int gd_frontend_launch_count(gd_frontend_t* h, long long* launches);
GD_API int gd_frontend_profile(gd_frontend_t* h, int enable);
GD_API int gd_frontend_profile_read(gd_frontend_t* h, int max_entries, const char** names, float* ms, long long* launches,
                             int* n_entries);
GD_API int gd_frontend_debug_fetch(gd_frontend_t* h, int what, int stream, void* dst, size_t dst_bytes);
/* flush L2 by writing a scratch buffer larger than the L2 (bench hygiene) */
GD_API int gd_frontend_flush_l2(gd_frontend_t* h);

/* ------------------------------------------------------------------ single stages (parity harness) -- */
GD_API int gd_stage_gray(int device, const uint8_t* bgr, size_t bgr_step, int w, int h, int order, uint8_t* gray);
GD_API int gd_stage_depth_edge(int device, const float* depth_m, int w, int h, const float K[9], uint8_t* edge);
/* lut: NULL or w*h*2 floats.  dist and mask may each be NULL. */
GD_API int gd_stage_mahalanobis(int device, const float* flow, const float* depth_ref, const float* depth_cur,
                         const uint8_t* edge_ref, const uint8_t* edge_cur, const float* lut, int w, int h,
                         const float K[9], const float R[9], const float T[3], float* dist, uint8_t* mask,
                         float* minmax);
GD_API int gd_stage_farneback(int device, const uint8_t* prev, const uint8_t* next, int w, int h, float* flow);
/* polynomial expansion of pyramid level k: out = 5 planes (lh*lw each) in OpenCV channel order */
GD_API int gd_stage_polyexp(int device, const uint8_t* gray, int w, int h, int k, float* out, int* lw, int* lh);

/* keep flags of the Frame-ctor filter (src/Frame.cc:258-282) for n keypoints against an 8UC1 mask (w x h, dense rows) */
GD_API int gd_stage_erode_filter(int device, const uint8_t* mask, int w, int h, const gd_keypoint* kps, int n, uint8_t* keep);
/* ORB pyramid (ComputePyramid, ORBextractor.cc:1107-1132): `out` receives the levels tightly packed one after another,
 * level_sizes = nlevels x (w, h) */
GD_API int gd_stage_orb_pyramid(int device, const uint8_t* gray, int w, int h, int nlevels, float scale, uint8_t* out,
                                int* level_sizes);
/* cell loop of ComputeKeyPointsOctTree (ORBextractor.cc:789-829) on `gray` taken as ONE pyramid level:
 * out = (x, y, response) float triples relative to the 16-px border, in vToDistributeKeys order */
GD_API int gd_stage_fast_cells(int device, const uint8_t* gray, int w, int h, int ini_th, int min_th, float* out, int capacity,
                               int* n);
/* GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) of an 8-bit image (ORBextractor.cc:1086) */
GD_API int gd_stage_gaussian7(int device, const uint8_t* gray, int w, int h, uint8_t* out);

/* ---- building blocks of GeoMaskMaker::GetRt (src/GeoMaskMaker.cc:77-156), SURVEY 8(f)-1, with host buffers: each bit-exact
 * against the cv2-pinned restatement oracle/getrt_proto.py.  They run the kernels of the resident stage (gd_frontend getrt /
 * gd_geomask_getrt_points) on a temporary one-stream instance. */
/* Everything GeoMaskMaker::GetRt does before solvePnPRansac (GeoMaskMaker.cc:82-141) for one host image pair: cv::ORB features
 * of both gray images, the Hamming cross-check matcher, the reference's sort / first-100 selection, undistortPoints (dist:
 * k1 k2 p1 p2 [k3], may be all zero or NULL), depth look-up and back-projection, all on the device.  object_points: up to
 * 100 x 3 floats (metres, first camera), image_pixels: up to 100 x 2 floats (second image), in the reference's order.  GetRt then
 * returns false when *n_points < 20 (:143-146) and otherwise calls cv::solvePnPRansac(objectPoints, imagePixels, K, D, rvec, T) +
 * cv::Rodrigues (:148-150), which stay in the caller. */
GD_API int gd_getrt_points(int device, const uint8_t* gray_first, const uint8_t* gray_second, int w, int h, const float* depth_first_m,
                           const float K[9], const float* dist, int ndist, float* object_points, float* image_pixels, int* n_points);
/* cv::ORB::create(nfeatures, 1.2f, 8, 31, 0, 2)->detectAndCompute(gray) of GeoMaskMaker.cc:82-90: keypoints (cv::KeyPoint layout) and
 * 32-byte descriptors in OpenCV's own order.  Pyramid, FAST, the two KeyPointsFilter::retainBest orderings (libstdc++'s
 * std::nth_element / std::partition restated for one device thread), Harris, blur, orientation, descriptors: all on the device */
GD_API int gd_stage_cvorb_detect_and_compute(int device, const uint8_t* gray, int w, int h, int nfeatures, gd_keypoint* kps,
                                             uint8_t* desc, int capacity, int* n);
/* cv::FAST(threshold, nonmaxSuppression) on a whole 8-bit image as cv::ORB runs it per level: kept[y*w+x] = S' (= response + 1)
 * at the corners that survive the 8-neighbour suppression, 0 elsewhere; raster order of the non-zeros = cv::FAST's output order */
GD_API int gd_stage_fast_whole(int device, const uint8_t* gray, int w, int h, int threshold, uint8_t* kept);
/* cv::resize(INTER_LINEAR_EXACT) between the pyramid levels of cv::ORB (GeoMaskMaker.cc:82, cv::ORB::detectAndCompute) */
GD_API int gd_stage_resize_linear_exact(int device, const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
/* GaussianBlur(7x7, sigma 2, BORDER_REFLECT_101) as cv::ORB gets it on a pyramid submatrix: the float separable path */
GD_API int gd_stage_gaussian7_float(int device, const uint8_t* src, int w, int h, uint8_t* dst);
/* HarrisResponses(blockSize 7, k 0.04) of cv::ORB at integer level coordinates (at least 4 px from the border) */
GD_API int gd_stage_harris(int device, const uint8_t* img, int w, int h, const int* xs, const int* ys, int n, float* out);
/* BFMatcher(NORM_HAMMING, crossCheck = true)->match(des_first, des_second) (GeoMaskMaker.cc:92-94): 32-byte descriptors,
 * matches ordered by query index, distance = number of differing bits */
GD_API int gd_stage_hamming_crosscheck(int device, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* query_idx,
                                       int* train_idx, int* distance, int capacity, int* n_matches);

#ifdef __cplusplus
}
#endif
#endif /* GDSLAM_CUDA_H_ */
