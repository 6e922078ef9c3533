// getrt.cuh — GeoMaskMaker::GetRt (GD-SLAM src/GeoMaskMaker.cc:77-156) up to the solvePnPRansac call as a resident, batched stage
// (SURVEY 8f-1): cv::ORB(2000, 1.2, 8, 31, 0, 2) features computed ONCE per frame and kept per ring slot (the reference
// extracts both images on every call although `first` was `second` five frames earlier, :419-428), Hamming cross-check
// matching, the reference's std::sort / first-100 selection, undistortPoints, depth look-up and back-projection — all on the
// device, bit-identical to OpenCV 4.13 + libstdc++ (stdalgo.cuh), no per-call allocation.
#pragma once
#include "geomask.cuh"
#include "orb.cuh"

namespace gd {

constexpr int GETRT_LEVELS = 8;
constexpr int GETRT_NFEATURES = 2000;  // ORB::create(2000, ...), GeoMaskMaker.cc:82
constexpr int GETRT_EDGE = 31;         // edgeThreshold
constexpr int GETRT_TOP = 100;         // good_matchs(matches.begin(), matches.begin() + 100), :97

struct GetRtCore {
    int device = 0, batch = 0, w = 0, h = 0, ring = 0;
    cudaStream_t stream = nullptr;  // borrowed: the caller sets it before enqueueing (front-end: a forked side stream)
    LaunchStats* stats = nullptr;
    CamConst cam;
    bool distorted = false;
    double Kd[4];       // fx, fy, cx, cy as doubles of the f32 K (cv::undistortPoints works in double)
    double dist[5];     // k1 k2 p1 p2 k3
    CvPyrArgs pyr_args;
    int nper[GETRT_LEVELS];
    size_t pyr_bytes = 0;  // one stream's pyramid (dense levels, 256-byte aligned offsets)
    int rows_total = 0;    // sum of level heights (row-count array)
    int n1_cap = 0, n2_cap = 0, sel_cap = 0, feat_cap = 0;
    size_t select_smem = 0;

    DevBuf pyr, kept, blur;         // [B][pyr_bytes]
    DevBuf rowcnt;                  // [B][rows_total] int
    DevBuf tabs;                    // cv::resize(INTER_LINEAR_EXACT) tables of levels 1..7: x table then y table (ushort4)
    int tab_x[GETRT_LEVELS] = {0}, tab_y[GETRT_LEVELS] = {0};
    DevBuf sel;                     // [B][levels][sel_cap] uint2 (level pixel index, bits of the Harris response)
    DevBuf sel_n;                   // [B][levels] int
    DevBuf feat_kp, feat_desc, feat_n;  // [ring][B][feat_cap] cv::KeyPoint records / [..][32] descriptors / [ring][B] counts
    DevBuf nn, dd;                  // [B][2][feat_cap] nearest neighbour index / distance of both matching directions
    DevBuf colkey;                  // [B][feat_cap] packed column minima of the distance matrix (direction 1)
    DevBuf out_obj, out_pix, out_cnt;   // [B][100][3] f32, [B][100][2] f32, [B] int: what GetRt hands to solvePnPRansac
    DevBuf err;                     // [1] int: capacity overflow flags of the selection kernel
    PinnedBuf h_out;                // pinned copy of out_obj | out_pix | out_cnt | err

    int nfeatures = GETRT_NFEATURES;
    int init(const float K[9], const float* dist_coef, int ndist, int width, int height, int device_, int batch_, int ring_slots,
             int nfeatures_ = GETRT_NFEATURES);
    // cv::ORB::detectAndCompute of the new frame of every stream (gray: [B] dense w x h images, stride gray_stride_b) into
    // ring slot `slot`
    int enqueue_features(const uint8_t* gray, size_t gray_stride_b, int slot);
    // BFMatcher cross-check match(first = ref slot, second = cur slot), sort, first 100, undistort, depth look-up, back-projection
    int enqueue_match(int ref_slot, int cur_slot, const float* depth_ref, size_t depth_stride_b);
    // async D2H of the points into the pinned block; `*_host` point into it after the caller synchronises `stream`
    int enqueue_fetch();
    const float* host_obj(int b) const { return h_out.as<float>() + (size_t)b * GETRT_TOP * 3; }
    const float* host_pix(int b) const { return h_out.as<float>() + (size_t)batch * GETRT_TOP * 3 + (size_t)b * GETRT_TOP * 2; }
    const int* host_cnt() const { return reinterpret_cast<const int*>(h_out.as<float>() + (size_t)batch * GETRT_TOP * 5); }
    int host_err() const { return host_cnt()[batch]; }
    gd_keypoint* slot_kp(int slot) { return feat_kp.as<gd_keypoint>() + (size_t)slot * batch * feat_cap; }
    uint8_t* slot_desc(int slot) { return feat_desc.as<uint8_t>() + (size_t)slot * batch * feat_cap * 32; }
    int* slot_n(int slot) { return feat_n.as<int>() + (size_t)slot * batch; }
};

}  // namespace gd
