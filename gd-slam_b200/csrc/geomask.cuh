// geomask.cuh — launchers of the GeoMaskMaker kernels (K0 gray, K2a depth edge, K2b Mahalanobis scatter,
// K3 min-max + normalise + threshold).  All launchers are batched over `batch` independent streams
// (blockIdx.z / blockIdx.y = stream) and asynchronous on `s`.
#pragma once
#include "gd_internal.h"

namespace gd {

// per-stream pose block read by K2b / K3 (uploaded once per step)
struct PoseDev {
    float R[9];
    float T[3];
    float RK[9];  // R * inv(K), f32 gemm semantics of OpenCV (GeoMaskMaker.cc:241)
    int valid;    // 0 -> all-ones mask (GetRt failure path / warm-up)
    int epoch;    // scatter-key epoch of this step (see KeyFormat); uploaded with the pose, never part of a captured graph
    int pad;
};

// 64-bit scatter key of K2b/K3: [ epoch | source index + 1 | bits(value) ].  The source index orders the writers of one
// target by raster position (atomicMax = last raster-order writer wins, GeoMaskMaker.cc:269); the epoch in the top bits
// makes every key of an earlier step smaller than any key of this step and lets the readers treat stale entries as
// "not written" (0.0f), so the key image is never cleared between frames (it was 8 bytes/pixel of stores per frame).
// When the epoch wraps (every 2^epoch_bits - 1 steps) the host clears the image once.
struct KeyFormat {
    int shift;      // 32 + index bits
    int epoch_max;  // 2^(64 - shift) - 1
};
KeyFormat make_key_format(size_t n_px);

// camera constants shared by all streams of a handle (kernel argument by value)
struct CamConst {
    float fu, fv, cu, cv;
    float rfu, rfv;  // RN(1/fu), RN(1/fv): reciprocals for the correctly rounded division by a constant
    float std_k;     // ((1/fu)^2 * 0.5) * 0.5, the depth-independent factor of Depth2Std (GeoMaskMaker.cc:1386-1391)
    float Ki[9];    // inv(K) in f32  (GeoMaskMaker.cc:200)
    double Kid[9];  // inv((double)K)  (GeoMaskMaker.cc:888)
};

// cv::undistortPoints(pt, K, D, noArray(), K) for one f32 point: normalise (multiplying by 1/fx like OpenCV), 5 fixed-point
// iterations (the default TermCriteria(COUNT, 5)), re-project, all in f64; D = k1 k2 p1 p2 k3.  Used by GetRt's matched
// points (GeoMaskMaker.cc:104-110), Frame::UndistortKeyPoints and ComputeImageBounds (Frame.cc:576-636).  Compile the callers
// without FMA contraction.
struct UndistortArgs {
    double fx, fy, cx, cy;
    double k[5];
};
#ifdef __CUDACC__
__host__ __device__
#endif
inline void undistort_point_cv(const UndistortArgs& a, float u, float v, float* ou, float* ov)
{
    const double ifx = 1.0 / a.fx, ify = 1.0 / a.fy;
    double x = ((double)u - a.cx) * ifx, y = ((double)v - a.cy) * ify;
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((a.k[4] * r2 + a.k[1]) * r2 + a.k[0]) * r2);
        if (icdist < 0) {
            x = ((double)u - a.cx) * ifx;
            y = ((double)v - a.cy) * ify;
            break;
        }
        const double dX = 2 * a.k[2] * x * y + a.k[3] * (r2 + 2 * x * x);
        const double dY = a.k[2] * (r2 + 2 * y * y) + 2 * a.k[3] * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ou = (float)(a.fx * x + a.cx);
    *ov = (float)(a.fy * y + a.cy);
}
void make_undistort_args(const float K[9], const float* dist, int ndist, UndistortArgs* a);

void make_cam_const(const float K[9], CamConst* c);
void make_pose(const float K[9], const float R[9], const float T[3], int valid, int epoch, PoseDev* p);

// K0: 8UC3 -> 8UC1 (cv::cvtColor 8-bit, 15-bit fixed point).  Either output may be null.
// gray_bgr2gray: rows of w bytes, stream stride gray_stride_b.  gray_orb: row pitch orb_pitch, stream stride orb_stride_b.
int launch_gray(const uint8_t* bgr, size_t bgr_step, size_t bgr_stride_b, int w, int h, int batch, uint8_t* gray_bgr2gray,
                size_t gray_stride_b, uint8_t* gray_orb, int orb_order, size_t orb_pitch, size_t orb_stride_b, cudaStream_t s,
                LaunchStats* st);

// K2a: GetEdge.  depth f32 [b][h][w] -> edge u8 {0,255}
int launch_depth_edge(const float* depth, size_t depth_stride_b, int w, int h, int batch, const CamConst& cam,
                      uint8_t* edge, size_t edge_stride_b, cudaStream_t s, LaunchStats* st);

// K2b: fused back-projection + J S J^T + 3x3 inverse + Mahalanobis + scatter (64-bit atomicMax keys)
int launch_mahalanobis(const float2* flow, size_t flow_stride_b, const float* depth_ref, const float* depth_cur,
                       size_t depth_stride_b, const uint8_t* edge_ref, const uint8_t* edge_cur, size_t edge_stride_b,
                       const float2* lut, int w, int h, int batch, const CamConst& cam, const PoseDev* poses, KeyFormat kf,
                       unsigned long long* keys, size_t keys_stride_b, cudaStream_t s, LaunchStats* st);

// K3 pass 1: per-stream min / max of the resolved dist image.  minmax_bits: [batch][GD_MM_WORDS] u32, must hold 0xFFFFFFFF
// on entry (launch_minmax_reset).  Encoding: [0] = min(bits(v)), [1] = min(~bits(v)), [2] = 0 when pixel 0 holds a NaN
// (cv::normalize's min/max scan starts from element 0 and ignores NaN everywhere else: a NaN there poisons the whole image).
constexpr int GD_MM_WORDS = 4;
int launch_minmax_reset(unsigned int* minmax_bits, int batch, cudaStream_t s);
int launch_minmax(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses, KeyFormat kf,
                  unsigned int* minmax_bits, cudaStream_t s, LaunchStats* st);
// K3 pass 2: normalise (0..255), round half even, < 20 -> mask {1,0}.  The keys are left alone (epoch tagged).
int launch_normalize_mask(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch,
                          const unsigned int* minmax_bits, const PoseDev* poses, KeyFormat kf, uint8_t* mask,
                          size_t mask_stride_b, cudaStream_t s, LaunchStats* st);
// K3 in one pass (thread-block cluster per stream, partial min/max exchanged through distributed shared memory); *used is
// false when the image is too large for a slice to stay in shared memory: the caller then runs the two kernels above
int launch_minmax_mask_cluster(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses,
                               KeyFormat kf, unsigned int* minmax_bits, uint8_t* mask, size_t mask_stride_b, cudaStream_t s,
                               LaunchStats* st, bool* used);
// debug / parity only: the resolved f32 dist image of the last step (GeoMaskMaker.cc:269, before normalize)
int launch_resolve_dist(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses,
                        KeyFormat kf, float* dist_out, size_t dist_stride_b, cudaStream_t s);
// "next" row (f)-2: Frame ctor erosion + keypoint filter (src/Frame.cc:258-282)
struct EllipseSE {
    int j1[31], j2[31];  // [j1, j2) columns of each row of cv::getStructuringElement(MORPH_ELLIPSE, 31x31)
};
void make_ellipse31(EllipseSE* se);
// keep[b][i] = eroded(mask_b)((int)kp.y, (int)kp.x) == 1 for the first n_kp[b] (or n_fixed) keypoints of every stream
int launch_erode_filter(const uint8_t* mask, size_t mask_stride_b, int w, int h, int batch, const gd_keypoint* kps, size_t cap,
                        const int* n_kp, int n_fixed, uint8_t* keep, cudaStream_t s, LaunchStats* st);
// order-preserving compaction of the kept keypoints + descriptors
int launch_compact_keypoints(const gd_keypoint* kps, const uint8_t* desc, const uint8_t* keep, size_t cap, int batch,
                             const int* n_kp, gd_keypoint* out_kps, uint8_t* out_desc, int* out_n, cudaStream_t s, LaunchStats* st);
// "next" row (f)-4: raw 16-bit depth -> metres, (float)v * inv_factor (Tracking.cc:234-235)
int launch_depth_u16_to_m(const uint16_t* raw, size_t raw_stride_b, float* depth, size_t depth_stride_b, size_t n, int batch,
                          float inv_factor, cudaStream_t s, LaunchStats* st);
// "next" row (f)-3: UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD / AssignFeaturesToGrid (Frame.cc:576-636,
// 815-837, 402-417, 553-565), one CTA per stream.  und: the camera (distorted iff und->k[0] != 0, the reference's own test);
// un_out (optional): mvKeysUn positions (x, y) per keypoint.
int launch_stereo_grid(const float* depth, size_t depth_stride_b, int w, int h, int batch, const gd_keypoint* kps, size_t cap,
                       const int* n_kp, float bf, const UndistortArgs& und, float* depth_out, float* uright, int* cell_start,
                       int* cell_items, float2* un_out, cudaStream_t s, LaunchStats* st);
// all-ones mask (warm-up path)
int launch_fill_u8(uint8_t* dst, size_t n, uint8_t v, cudaStream_t s, LaunchStats* st);

}  // namespace gd
