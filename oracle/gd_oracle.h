/*
 * gd_oracle.h — CPU oracle for the GD-SLAM GeoMaskMaker + ORBextractor hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is product code: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load it.  The product path (gd-slam_b200/csrc, include/gdslam_cuda.h) never links,
 * imports or calls anything declared here.
 *
 * It is a plain-C/C++ restatement of the reference's own loops, with the un-vendored
 * OpenCV primitives restated with OpenCV-4.13 semantics.  Each
 * function cites the reference file:line (relative to the GD-SLAM tree) it follows.
 *
 * Pinning: the reference has no tests / golden vectors, so the
 * oracle is pinned against (i) literal per-pixel evaluation with Python cv2 4.13
 * (tests/golden/make_golden.py -> tests/golden/ fixtures) for every OpenCV primitive and
 * for the Mahalanobis / depth-edge arithmetic, (ii) cv2.calcOpticalFlowFarneback output,
 * (iii) the reference's src/ORBextractor.cc compiled verbatim against a stand-in cv header
 * (oracle/_ref, built by oracle/Makefile where /root/reference exists).
 *
 * Build: -O2 -ffp-contract=off (no FMA contraction; the two fused operations OpenCV 4.13
 * really performs with FMA — scaleAdd and convertTo-with-scale — are written as explicit
 * fmaf()).
 */
#ifndef GD_ORACLE_H_
#define GD_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- geomask_oracle.c ------------------------------------------------------------- */

/* cvtColor 8UC3 -> 8UC1.  order 0: COLOR_BGR2GRAY (GeoMaskMaker.cc:163-164),
 * order 1: COLOR_RGB2GRAY applied to the same bytes (Tracking.cc:219-225, Camera.RGB=1). */
void gdo_gray_u8(const uint8_t* src, size_t src_step, int w, int h, int order, uint8_t* dst);

/* cv::invert of a 3x3 (closed form, f64 determinant + cofactors). Returns 0 if singular. */
int gdo_inv3_f32(const float* m, float* out);
int gdo_inv3_f64(const double* m, double* out);

/* GeoMaskMaker::GetEdge (GeoMaskMaker.cc:854-964).  depth: metres f32, K: 3x3 f32.
 * edge out: u8 {0,255}. */
void gdo_depth_edge(const float* depth, int w, int h, const float* K, uint8_t* edge);

/* Main per-pixel loop of GetNoGMMmask (GeoMaskMaker.cc:190-272).  lut may be NULL
 * (identity undistortion = TUM3) or w*h*2 floats (x,y) like undistortedPoint.
 * dist out: f32 w*h (zeros where never written).  written (optional, may be NULL): u8 {0,1}.
 * src_index (optional): int32 index of the winning source pixel or -1. */
void gdo_mahalanobis(const float* flow /*w*h*2*/, const float* depth_ref, const float* depth_cur,
                     const uint8_t* edge_ref, const uint8_t* edge_cur, const float* lut,
                     int w, int h, const float* K, const float* R, const float* T,
                     float* dist, uint8_t* written, int32_t* src_index);

/* normalize(NORM_MINMAX,0,255) + convertTo(8U) + (<20)/255 (GeoMaskMaker.cc:276-277,405-407).
 * d8 (optional) receives the 8-bit normalised image; minmax (optional) receives {min,max}. */
void gdo_normalize_threshold(const float* dist, int w, int h, uint8_t* mask, uint8_t* d8, float* minmax);

/* "next" row (f)-2, Frame ctor (src/Frame.cc:258-282): erode(mask, 31x31 ellipse) and the keypoint filter */
void gdo_ellipse31(int* j1, int* j2);
void gdo_erode31(const uint8_t* mask, int w, int h, uint8_t* out);
int gdo_erode_filter(const uint8_t* mask, int w, int h, const float* kps, int n, uint8_t* keep);

/* "next" row (f)-3, Frame::ComputeStereoFromRGBD + AssignFeaturesToGrid (src/Frame.cc:815-837, 402-417, 553-565), D = 0 */
void gdo_undistort_point(const float* K, const float* D, float u, float v, float* ou, float* ov);
void gdo_stereo_grid(const float* depth_m, int w, int h, const float* kps, int n, float bf, const float* K, const float* D,
                     float* depth_out, float* uright, int* cell_start, int* cell_items, float* un_out, float* bounds_out);

/* depth2std (GeoMaskMaker.cc:1386-1391) */
float gdo_depth2std(float depth, float fu);

/* ---- farneback_oracle.c ----------------------------------------------------------- */

/* cv::calcOpticalFlowFarneback(prev,next,flow,pyr_scale,levels,winsize,iterations,poly_n,
 * poly_sigma,flags=0) as called at GeoMaskMaker.cc:165 with (0.5,3,15,3,5,1.2,0).
 * prev/next: 8UC1 w*h; flow out: w*h*2 f32. */
void gdo_farneback(const uint8_t* prev, const uint8_t* next, int w, int h, double pyr_scale, int levels,
                   int winsize, int iterations, int poly_n, double poly_sigma, float* flow);

/* the per-image half of it: blur + resize + FarnebackPolyExp for pyramid level k.
 * out: lw*lh*5 f32 (interleaved like OpenCV's CV_32FC5). lw/lh returned. */
void gdo_farneback_polyexp_level(const uint8_t* img, int w, int h, double pyr_scale, int k, int poly_n,
                                 double poly_sigma, float* out, int* lw, int* lh);

/* ---- whole GetNoGMMmask on one (ref,cur) pair, pose given (GeoMaskMaker.cc:167-277,405-407) */
void gdo_geomask_pair(const uint8_t* bgr_ref, const uint8_t* bgr_cur, const float* depth_ref,
                      const float* depth_cur, int w, int h, const float* K, const float* R, const float* T,
                      uint8_t* mask, float* flow_out /*optional*/, float* dist_out /*optional*/);

#ifdef __cplusplus
}
#endif
#endif /* GD_ORACLE_H_ */
