// geomask_core.cuh — device-resident state of `batch` GeoMaskMaker streams (ring buffer of per-image products)
// and the per-frame sequence AddNewImage / GetNoGMMmask (GD-SLAM src/GeoMaskMaker.cc:167-277, 405-429).
#pragma once
#include <memory>

#include "farneback.cuh"
#include "geomask.cuh"
#include "getrt.cuh"

namespace gd {

constexpr int GD_RING = 6;  // inter_frame_size (5) + 1, include/GeoMaskMaker.h:55

struct GeoMaskCore {
    int device = 0, batch = 0, w = 0, h = 0;
    size_t n = 0;  // pixels per image (padded strides below)
    float K[9];
    CamConst cam;
    FbPlan plan;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    LaunchStats* stats = nullptr;
    int frames = 0;  // frames pushed so far
    bool has_lut = false;

    // device memory (HBM layout: stream-major, ring-slot-minor)
    DevBuf bgr;        // [B][h][w][3]   newest frame (input staging)
    DevBuf gray;       // [B][n]         BGR2GRAY of the newest frame
    DevBuf depth;      // [B][RING][n]   f32 metres
    DevBuf edge;       // [B][RING][n]   u8 {0,255}
    DevBuf R;          // [B][RING][r_floats]  Farnebäck polynomial expansion pyramid
    DevBuf scratchI;   // [B][i_floats]
    DevBuf flowA, flowB;  // [B][f_float2]
    DevBuf Mbuf, Mbuf2;   // [B][m_floats] UpdateMatrices output of the current level (split flow form), ping-pong pair
    FbFlowBuffers flow_bufs;
    bool split_flow = true;
    DevBuf keys;       // [B][n] u64 scatter keys, epoch tagged (KeyFormat): never cleared between frames
    KeyFormat keyfmt;
    int epoch = 0;     // epoch of the last uploaded pose block
    bool epoch_used = false;
    DevBuf minmax;     // [B][GD_MM_WORDS] u32
    DevBuf poses;      // [B] PoseDev
    DevBuf mask;       // [B][n] u8
    DevBuf dist;       // [B][n] f32 resolved dist image: allocated and filled by debug_fetch(GD_DBG_DIST) only
    DevBuf lut;        // [n] float2 or empty
    // pinned staging of the pose blocks: a ring, because the H2D copy reads the buffer when it EXECUTES and a caller of the
    // device-resident path may enqueue several steps without synchronising (each slot is guarded by an event)
    static constexpr int POSE_SLOTS = 8;
    PinnedBuf h_poses;  // [POSE_SLOTS][B] PoseDev
    cudaEvent_t pose_ev[POSE_SLOTS] = {};
    bool pose_pending[POSE_SLOTS] = {};
    int pose_slot = 0;
    const float2* last_flow = nullptr;
    int last_ref_slot = -1, last_cur_slot = -1;

    size_t n_pad = 0;  // n rounded up to 64 elements

    int init(const float K_[9], const float* dist_coef, int ndist, int width, int height, int device_, int batch_,
             cudaStream_t s, LaunchStats* st);
    // new frame already resident in this->bgr and in depth slot (frames % RING): computes the per-image products
    // optional second stream for the depth-edge kernel (independent of the flow chain); the caller orders it:
    // forked after the depth upload, joined before enqueue_mask()
    cudaStream_t edge_stream = nullptr;
    // graph replay of the two launch sequences for stand-alone handles (gd_geomask_*); disabled inside the batched front-end
    GraphCache push_graphs, mask_graphs;
    // GeoMaskMaker::GetRt as a resident stage (row f-1): cv::ORB features of every pushed frame kept per ring slot, the pair
    // (t-5, t) matched and back-projected on demand.  Off until enable_getrt().
    std::unique_ptr<GetRtCore> getrt;
    float dist_coef[5] = {0, 0, 0, 0, 0};
    int ndist = 0;
    long long feat_frame[GD_RING];  // frame number whose features sit in the slot (-1: none)
    cudaStream_t getrt_stream = nullptr;  // optional side stream for the GetRt chain (the caller forks / joins it)
    int enable_getrt();
    int enqueue_getrt_match();            // match + points + D2H of the buffered pair; needs features of both slots
    bool getrt_pair_ready() const;
    int enqueue_flow();                   // first half of enqueue_mask(): Farneback of the buffered pair
    int enqueue_mask_tail();              // second half: Mahalanobis scatter, min/max, normalise + threshold
    int enqueue_push(int slot, bool gray_done);
    int push_resident(bool gray_done = false);
    float* depth_slot_ptr(int slot) { return depth.as<float>() + (size_t)slot * n_pad; }
    size_t depth_stride_b() const { return (size_t)GD_RING * n_pad; }
    int cur_slot() const { return frames % GD_RING; }  // slot the NEXT push writes
    // GetNoGMMmask for all streams; result stays in this->mask.  R: [B][9], T: [B][3], valid: [B] (host)
    int compute_mask(const float* R, const float* T, const int* pose_valid);
    // pose blocks -> device (async, never part of a captured graph); frames_pushed = number of frames pushed when the
    // mask will be evaluated
    int upload_poses(const float* R, const float* T, const int* pose_valid, int frames_pushed);
    int enqueue_mask();  // everything GetNoGMMmask launches (graph capturable); expects upload_poses() before it
    int debug_fetch(int what, int stream_idx, void* dst, size_t dst_bytes);
    ~GeoMaskCore();
};

}  // namespace gd
