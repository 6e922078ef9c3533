// farneback.cuh — dense optical flow (cv::calcOpticalFlowFarneback, flags = 0) as a batched CUDA pipeline.
// Same algorithm and parameters the reference invokes at src/GeoMaskMaker.cc:165: (0.5, 3, 15, 3, 5, 1.2, 0).
#pragma once
#include <cuda.h>  // CUtensorMap (type only)

#include "gd_internal.h"

namespace gd {

constexpr int FB_MAX_LEVELS = 8;
constexpr int FB_MAX_KSIZE = 41;
constexpr int FB_POLY_N = 5;
constexpr int FB_WIN = 15;

struct FbLevel {
    int w, h;           // level size
    int ksize;          // Gaussian taps applied on the FULL-RES image before resampling
    float taps[FB_MAX_KSIZE];
    double scale_x, scale_y;  // src/dst ratios of cv::resize
    size_t r_off;       // offset (floats) of this level's R planes inside one image's R pyramid
    size_t i_off;       // offset (floats) of this level inside the I scratch
    size_t f_off;       // offset (float2) inside the flow scratch
};

struct FbPlan {
    int w = 0, h = 0;
    int nlevels = 0;     // number of pyramid levels actually used (k = nlevels-1 .. 0), incl. level 0
    int iterations = 3;
    FbLevel lv[FB_MAX_LEVELS];
    size_t r_floats = 0;   // floats of one image's whole R pyramid (5 planes per level)
    size_t i_floats = 0;   // floats of one image's I scratch (all levels) + the row-pass scratch
    size_t hrow_off = 0;   // offset (floats) of the float2 row-pass scratch inside the I scratch
    size_t f_float2 = 0;   // float2 of one flow scratch (all levels)
    size_t m_floats = 0;   // floats of the M scratch (5 planes of the largest level)
    // FarnebackPolyExp constants
    float g[2 * FB_POLY_N + 1], xg[2 * FB_POLY_N + 1], xxg[2 * FB_POLY_N + 1];
    double ig11, ig03, ig33, ig55;
};

int fb_make_plan(int w, int h, double pyr_scale, int levels, int iterations, int poly_n, double poly_sigma, int winsize,
                 FbPlan* plan);

// per-image half: gray u8 [b][h][w] -> R pyramid (planar, 5 planes per level) [b][r_floats]
// scratch_I: [b][i_floats]
int fb_launch_pyramid_polyexp(const FbPlan& plan, const uint8_t* gray, size_t gray_stride_b, int batch, float* scratch_I,
                              size_t i_stride_b, float* R, size_t r_stride_b, cudaStream_t s, LaunchStats* st);

// M scratch of the split flow form and how the box/solve kernel reads it.
//   M[0]            [batch][m_floats] (or a smaller group window of m_bytes): UpdateMatrices output
//   M[1]            second buffer of the same size: enables the fused "box/solve + next iteration's matrices" kernel
//                   (M ping-pong, no intermediate flow, 4 launches per level instead of 6); nullptr = plain split form
//   tmap[level][i]  rank-3 TMA descriptors (x, y, plane) of M[i] at every level whose rows are 16-byte aligned: the
//                   84 x 46 channel tile of the box filter arrives by one cp.async.bulk.tensor per channel
// GD_FLOW_NEXT=0 / GD_FLOW_TMA=0 switch the two features off, GD_FLOW_BOX_F64=1 / GD_FLOW_NBUF=5 select the box kernel's
// accumulation type and prefetch depth (A/B measurements).
struct FbFlowBuffers {
    float* M[2] = {nullptr, nullptr};
    size_t m_bytes = 0;
    bool fuse_next = false, use_tma = false;
    bool box_f32 = true;  // 15 x 15 box sums in f32 tree form (default) or f64 running sums like OpenCV (GD_FLOW_BOX_F64=1)
    int min_blocks = 3;   // register budget of the f32 box kernel: 2 or 3 resident CTAs per SM (GD_FLOW_MB)
    int nbuf = 2;         // channel tiles a CTA of the box kernel keeps in flight (GD_FLOW_NBUF = 2 or 5)
    bool tmap_ok[FB_MAX_LEVELS][2] = {};
    alignas(64) CUtensorMap tmap[FB_MAX_LEVELS][2];
};
int fb_prepare_flow_buffers(const FbPlan& plan, int batch, float* M0, float* M1, size_t m_bytes_each, FbFlowBuffers* fb);

// per-pair half: R0, R1 -> flow (level 0, float2 [b][h][w]).  flow scratch: two buffers [b][f_float2].
// On return *final points at the level-0 flow (inside flowA or flowB).
// fbuf: the M scratch (split form); nullptr selects the older single-kernel flow iteration.
int fb_launch_flow(const FbPlan& plan, const float* R0, const float* R1, size_t r_stride_b, int batch, float2* flowA,
                   float2* flowB, size_t f_stride_b, const FbFlowBuffers* fbuf, const float2** final_flow, cudaStream_t s,
                   LaunchStats* st);

}  // namespace gd
