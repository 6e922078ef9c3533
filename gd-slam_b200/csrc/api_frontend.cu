// api_frontend.cu — C ABI: gd_frontend_* — the per-frame sequence of Tracking::GrabImageRGBD_GD
// (GD-SLAM src/Tracking.cc:212-252: cvtColor, Frame() -> ORBextractor, AddNewImage, GetNoGMMmask) for `batch`
// independent RGB-D streams stepped together on one GPU, one CUDA stream, BGR uploaded once per frame.
#include "geomask_core.cuh"
#include "orb.cuh"

#include <cstdlib>
#include <map>
#include <tuple>
#include <vector>

namespace gd {
int orb_fetch_results(OrbCore& c, gd_keypoint* const* kps, uint8_t* const* desc, int capacity, int* n_out);
}

struct gd_frontend {
    gd_frontend_config cfg;
    gd::LaunchStats stats;
    cudaStream_t stream = nullptr;
    gd::GeoMaskCore geo;
    gd::OrbCore orb;
    gd::DevBuf staged_bgr;    // [slots][B][n_pad*3]
    gd::DevBuf staged_depth;  // [slots][B][n_pad] f32
    gd::DevBuf l2_scratch;
    gd::DevBuf raw_depth;  // [B][n_pad] u16 staging of gd_frontend_step_u16
    gd::DevBuf sg_depth, sg_uright, sg_start, sg_items, sg_un;  // row f-3 outputs
    bool filtered_ready = false;
    gd::DevBuf keep, filt_kp, filt_desc, filt_n;  // Frame-ctor filter (row f-2): [B][cap] flags / records, [B] counts
    gd::PinnedBuf h_n;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool results_ready = false;
    // The ORB chain (pyramid, FAST, quadtree, blur, describe) and the GeoMask chain (pyramids, flow, edges, Mahalanobis)
    // only share the gray conversion: they run on two streams forked after K0 and joined at the end of the step, so the
    // many small ORB launches fill the SMs the flow kernels leave idle.  GD_OVERLAP=0 serialises them on one stream.
    cudaStream_t aux_stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_edge = nullptr, ev_join = nullptr;
    bool overlap = true;
    // GetRt stage (cfg.getrt): its chain (cv::ORB features of the new frame, matching against the frame five steps back, the
    // 100 points of solvePnPRansac and their D2H) runs on a third forked stream.  With a pose hook installed and no pose
    // given, the step waits for the points in the middle, asks the hook for (R, T) and only then enqueues the Mahalanobis half.
    cudaStream_t rt_stream = nullptr;
    cudaEvent_t ev_rt = nullptr;
    gd_pose_hook_fn pose_hook = nullptr;
    void* pose_hook_user = nullptr;
    bool getrt_points_ready = false;
    // CUDA graphs of the per-frame device work, one per (ring phase, input buffer): the launch sequence of a step is
    // static, so small batches (launch bound: 51 launches per frame) replay a graph instead of re-issuing every launch
    struct GraphEntry {
        cudaGraphExec_t exec = nullptr;
        long long launches = 0;
    };
    std::map<std::tuple<int, const void*, size_t>, GraphEntry> graphs;
    bool use_graphs = false;
    int graph_fallbacks = 0;  // captures / instantiations that failed (the reason is left in gd_last_error)
    ~gd_frontend()
    {
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (ev_edge) cudaEventDestroy(ev_edge);
        if (ev_rt) cudaEventDestroy(ev_rt);
        if (rt_stream) cudaStreamDestroy(rt_stream);
        if (aux_stream) cudaStreamDestroy(aux_stream);
        // cores do not own the shared stream
        for (auto& kv : graphs)
            if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
        if (stream) cudaStreamDestroy(stream);
    }
};

using namespace gd;

// per-frame device work once the new frame sits in geo.bgr and in the depth ring slot.
// part 1 = everything up to (and including) the flow; part 2 = Mahalanobis + mask + joins.  A plain step runs both back to
// back (and a graph captures both); the pose-hook path fetches the GetRt points between them.
static int frontend_enqueue_head(gd_frontend* h, const uint8_t* bgr_dev, size_t bgr_stride_b, bool* forked)
{
    GeoMaskCore& g = h->geo;
    PdlScope pdl_scope(g.batch);
    OrbCore& o = h->orb;
    // K0: both grays in one pass over the BGR bytes (BGR2GRAY for the flow, cfg.orb_gray_order for ORB level 0)
    GD_TRY(launch_gray(bgr_dev, (size_t)g.w * 3, bgr_stride_b, g.w, g.h, g.batch, g.gray.as<uint8_t>(), g.n_pad, o.level0(0),
                       h->cfg.orb_gray_order, (size_t)o.plan.lv[0].pitch, o.plan.pyr_bytes, h->stream, &h->stats));
    // (launch_gray above and everything below is what a graph replays)
    const bool fork = h->overlap && !h->stats.profiling;  // the per-family event profile wants serialised kernels
    *forked = fork;
    if (!fork) {
        GD_TRY(o.extract_resident());       // Frame() -> ORBextractor::operator()   (Tracking.cc:238)
        GD_TRY(g.push_resident(true));      // AddNewImage                          (Tracking.cc:242)
        if (g.getrt && g.getrt_pair_ready()) GD_TRY(g.enqueue_getrt_match());
        GD_TRY(g.enqueue_flow());           // GetNoGMMmask, first half               (Tracking.cc:245)
        return GD_OK;
    }
    // aux stream: depth edges (only needed by the Mahalanobis kernel at the end) then the ORB chain;
    // rt stream: the GetRt chain; main stream: flow pyramids, polynomial expansion, flow iterations
    GD_CUDA(cudaEventRecord(h->ev_fork, h->stream));
    GD_CUDA(cudaStreamWaitEvent(h->aux_stream, h->ev_fork, 0));
    if (g.getrt) GD_CUDA(cudaStreamWaitEvent(h->rt_stream, h->ev_fork, 0));
    g.edge_stream = h->aux_stream;
    g.getrt_stream = g.getrt ? h->rt_stream : nullptr;
    int rc = g.push_resident(true);
    g.edge_stream = nullptr;
    if (rc == GD_OK && g.getrt && g.getrt_pair_ready()) rc = g.enqueue_getrt_match();
    g.getrt_stream = nullptr;
    if (rc != GD_OK) return rc;
    if (g.getrt) GD_CUDA(cudaEventRecord(h->ev_rt, h->rt_stream));
    GD_CUDA(cudaEventRecord(h->ev_edge, h->aux_stream));
    o.stream = h->aux_stream;
    rc = o.extract_resident();
    o.stream = h->stream;
    if (rc != GD_OK) return rc;
    GD_CUDA(cudaEventRecord(h->ev_join, h->aux_stream));
    GD_TRY(g.enqueue_flow());
    return GD_OK;
}

static int frontend_enqueue_tail(gd_frontend* h, bool forked)
{
    GeoMaskCore& g = h->geo;
    PdlScope pdl_scope(g.batch);
    if (forked) GD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_edge, 0));
    GD_TRY(g.enqueue_mask_tail());
    if (forked) {
        GD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_join, 0));
        if (g.getrt) GD_CUDA(cudaStreamWaitEvent(h->stream, h->ev_rt, 0));
    }
    h->results_ready = true;
    h->filtered_ready = false;
    h->getrt_points_ready = g.getrt && g.getrt_pair_ready();
    return GD_OK;
}

static int frontend_enqueue(gd_frontend* h, const uint8_t* bgr_dev, size_t bgr_stride_b)
{
    bool forked = false;
    GD_TRY(frontend_enqueue_head(h, bgr_dev, bgr_stride_b, &forked));
    return frontend_enqueue_tail(h, forked);
}

// GeoMaskMaker::GetRt through the caller's solvePnPRansac: wait for the points of this step, ask the hook for every stream
// with at least 20 points (GeoMaskMaker.cc:143-146), build the pose arrays
static int frontend_poses_from_hook(gd_frontend* h, std::vector<float>& R, std::vector<float>& T, std::vector<int>& valid)
{
    GeoMaskCore& g = h->geo;
    const int B = g.batch;
    R.assign((size_t)9 * B, 0.f);
    T.assign((size_t)3 * B, 0.f);
    valid.assign((size_t)B, 0);
    if (!g.getrt_pair_ready()) return GD_OK;  // warm-up: no pair yet
    GD_CUDA(cudaEventSynchronize(h->ev_rt));
    if (g.getrt->host_err() != 0) {
        set_error("GetRt stage: capacity overflow of the selection lists (flags %d)", g.getrt->host_err());
        return GD_EINTERNAL;
    }
    for (int b = 0; b < B; ++b) {
        const int n = g.getrt->host_cnt()[b];
        if (n < 20) continue;  // "small feature match.": GetRt returns false -> all-ones mask
        valid[b] = h->pose_hook(h->pose_hook_user, b, g.getrt->host_obj(b), g.getrt->host_pix(b), n, &R[9 * b], &T[3 * b]) ? 1 : 0;
    }
    return GD_OK;
}

static int frontend_compute(gd_frontend* h, const uint8_t* bgr_dev, size_t bgr_stride_b, const float* R, const float* T,
                            const int* pose_valid)
{
    GeoMaskCore& g = h->geo;
    if (!R && !T && h->pose_hook && g.getrt) {
        // pose from the GetRt stage + the caller's solvePnPRansac: plain launches, one host synchronisation in the middle
        bool forked = false;
        GD_TRY(frontend_enqueue_head(h, bgr_dev, bgr_stride_b, &forked));
        if (!forked && g.getrt_pair_ready()) GD_CUDA(cudaEventRecord(h->ev_rt, h->stream));
        std::vector<float> Rh, Th;
        std::vector<int> vh;
        GD_TRY(frontend_poses_from_hook(h, Rh, Th, vh));
        GD_TRY(g.upload_poses(Rh.data(), Th.data(), vh.data(), g.frames));
        return frontend_enqueue_tail(h, forked);
    }
    // the frame of this step is pushed before the mask is evaluated; the pose copy stays outside any captured graph
    GD_TRY(g.upload_poses(R, T, pose_valid, g.frames + 1));
    // graphs only in steady state (every code path has run un-captured at least once: lazy attribute setup, ring full)
    // and never while the per-family event profile is on (events + synchronisation inside the launch scopes)
    if (!h->use_graphs || h->stats.profiling || g.frames < 2 * GD_RING) return frontend_enqueue(h, bgr_dev, bgr_stride_b);
    const auto key = std::make_tuple(g.frames % GD_RING, (const void*)bgr_dev, bgr_stride_b);
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        const int frames0 = g.frames;
        const long long l0 = h->stats.launches;
        cudaGraph_t graph = nullptr;
        GD_CUDA(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = frontend_enqueue(h, bgr_dev, bgr_stride_b);
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &graph);
        if (rc != GD_OK || ce != cudaSuccess || !graph) {
            if (graph) cudaGraphDestroy(graph);
            // not fatal, but recorded: gd_last_error() tells why this handle runs plain launches from now on
            set_error("CUDA graph capture of the front-end step failed (enqueue rc %d, %s): falling back to plain launches", rc,
                      cudaGetErrorString(ce));
            cudaGetLastError();
            h->use_graphs = false;  // fall back to plain launches for good
            h->graph_fallbacks += 1;
            g.frames = frames0;
            h->stats.launches = l0;
            return frontend_enqueue(h, bgr_dev, bgr_stride_b);
        }
        gd_frontend::GraphEntry e;
        e.launches = h->stats.launches - l0;
        const cudaError_t ie = cudaGraphInstantiate(&e.exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) {
            set_error("cudaGraphInstantiate of the front-end step failed (%s): falling back to plain launches", cudaGetErrorString(ie));
            cudaGetLastError();
            h->use_graphs = false;
            h->graph_fallbacks += 1;
            g.frames = frames0;
            h->stats.launches = l0;
            return frontend_enqueue(h, bgr_dev, bgr_stride_b);
        }
        h->stats.launches = l0;  // the capture issued nothing; the replay below is what runs
        g.frames = frames0;
        it = h->graphs.emplace(key, e).first;
    }
    GD_CUDA(cudaGraphLaunch(it->second.exec, h->stream));
    h->stats.launches += it->second.launches;
    // host-side state the enqueue path advances
    if (g.getrt) g.feat_frame[g.frames % GD_RING] = g.frames;
    g.frames += 1;
    g.last_cur_slot = (g.frames - 1) % GD_RING;
    g.last_ref_slot = (g.frames - GD_RING) % GD_RING;
    h->results_ready = true;
    h->filtered_ready = false;
    h->getrt_points_ready = g.getrt && g.getrt_pair_ready();
    return GD_OK;
}

extern "C" {

int gd_frontend_create(gd_frontend_t** out, const gd_frontend_config* cfg)
{
    GD_REQUIRE(out && cfg, "null argument");
    *out = nullptr;
    GD_TRY(select_device(cfg->device));
    gd_frontend* h = new (std::nothrow) gd_frontend();
    if (!h) return GD_ENOMEM;
    h->cfg = *cfg;
    int r = GD_OK;
    do {
        int main_prio = 0;
        {  // GD_MAIN_PRIO=1: the flow stream at the highest stream priority (A/B switch)
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char* e = std::getenv("GD_MAIN_PRIO");
            if (e && std::atoi(e) > 0) main_prio = hi;
        }
        if (cudaStreamCreateWithPriority(&h->stream, cudaStreamNonBlocking, main_prio) != cudaSuccess) {
            set_error("cudaStreamCreate failed");
            r = GD_ECUDA;
            break;
        }
        if ((r = h->geo.init(cfg->K, cfg->dist, cfg->ndist, cfg->width, cfg->height, cfg->device, cfg->batch, h->stream, &h->stats)) != GD_OK) break;
        if ((r = h->orb.init(cfg->nfeatures, cfg->scale_factor, cfg->nlevels, cfg->ini_th_fast, cfg->min_th_fast, cfg->width,
                             cfg->height, cfg->device, cfg->batch, h->stream, &h->stats)) != GD_OK)
            break;
        // the front-end captures its whole step itself: the per-core graph caches must never capture inside that capture
        h->geo.push_graphs.enabled = h->geo.mask_graphs.enabled = h->orb.graphs.enabled = false;
        if (cfg->staged_slots > 0) {
            const size_t B = (size_t)cfg->batch, S = (size_t)cfg->staged_slots;
            if ((r = h->staged_bgr.alloc(S * B * h->geo.n_pad * 3)) != GD_OK) break;
            if ((r = h->staged_depth.alloc(S * B * h->geo.n_pad * sizeof(float))) != GD_OK) break;
        }
        if ((r = h->h_n.alloc(sizeof(int) * (size_t)(cfg->batch + 1))) != GD_OK) break;
        {  // graphs pay off when a step is launch bound (small batches); GD_GRAPHS=0/1 overrides
            const char* e = std::getenv("GD_GRAPHS");
            h->use_graphs = e ? std::atoi(e) != 0 : cfg->batch <= 8;
        }
        {
            const char* e = std::getenv("GD_OVERLAP");
            h->overlap = e ? std::atoi(e) != 0 : true;
        }
        // GD_AUX_PRIO: stream priority of the auxiliary (depth edges + ORB) stream relative to the flow stream: 1 = higher,
        // -1 = lower, 0 / unset = same (A/B switch)
        int aux_prio = 0;
        {
            int lo = 0, hi = 0;  // lo = numerically largest = least priority
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            const char* e = std::getenv("GD_AUX_PRIO");
            const int want = e ? std::atoi(e) : 0;
            aux_prio = want > 0 ? hi : (want < 0 ? lo : 0);
            if (aux_prio < hi) aux_prio = hi;
            if (aux_prio > lo) aux_prio = lo;
        }
        if (cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, aux_prio) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_edge, cudaEventDisableTiming) != cudaSuccess ||
            cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
            set_error("cudaStreamCreate/cudaEventCreate failed");
            r = GD_ECUDA;
            break;
        }
        if (cfg->getrt) {
            if ((r = h->geo.enable_getrt()) != GD_OK) break;
            if (cudaStreamCreateWithFlags(&h->rt_stream, cudaStreamNonBlocking) != cudaSuccess ||
                cudaEventCreateWithFlags(&h->ev_rt, cudaEventDisableTiming) != cudaSuccess) {
                set_error("cudaStreamCreate/cudaEventCreate failed");
                r = GD_ECUDA;
                break;
            }
        }
        if (cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
            set_error("cudaEventCreate failed");
            r = GD_ECUDA;
            break;
        }
    } while (0);
    if (r != GD_OK) {
        delete h;
        return r;
    }
    *out = h;
    return GD_OK;
}

void gd_frontend_destroy(gd_frontend_t* h)
{
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->aux_stream) cudaStreamSynchronize(h->aux_stream);
    if (h->rt_stream) cudaStreamSynchronize(h->rt_stream);
    delete h;
}

static int upload_frames(gd_frontend* h, uint8_t* bgr_dst, float* depth_dst, size_t depth_stride_b, const uint8_t* const* bgr,
                         size_t bgr_step, const float* const* depth_m, size_t depth_step)
{
    GeoMaskCore& g = h->geo;
    GD_REQUIRE(bgr_step >= (size_t)g.w * 3 && depth_step >= (size_t)g.w * sizeof(float), "step smaller than a row");
    // one copy for the whole batch when the caller's frames are densely packed back to back (pinned batch buffers)
    bool packed_bgr = bgr_step == (size_t)g.w * 3 && g.n_pad == g.n, packed_d = depth_step == (size_t)g.w * 4 && g.n_pad == g.n;
    for (int b = 0; b < g.batch; ++b) {
        GD_REQUIRE(bgr[b] && depth_m[b], "null image pointer");
        if (b > 0) {
            packed_bgr = packed_bgr && bgr[b] == bgr[b - 1] + g.n * 3;
            packed_d = packed_d && depth_m[b] == depth_m[b - 1] + g.n;
        }
    }
    if (packed_bgr) {
        GD_CUDA(cudaMemcpyAsync(bgr_dst, bgr[0], (size_t)g.batch * g.n * 3, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int b = 0; b < g.batch; ++b)
            GD_CUDA(cudaMemcpy2DAsync(bgr_dst + (size_t)b * g.n_pad * 3, (size_t)g.w * 3, bgr[b], bgr_step, (size_t)g.w * 3, g.h,
                                      cudaMemcpyHostToDevice, h->stream));
    }
    if (packed_d && depth_stride_b == g.n) {
        GD_CUDA(cudaMemcpyAsync(depth_dst, depth_m[0], (size_t)g.batch * g.n * 4, cudaMemcpyHostToDevice, h->stream));
    } else if (packed_d) {
        GD_CUDA(cudaMemcpy2DAsync(depth_dst, depth_stride_b * 4, depth_m[0], g.n * 4, g.n * 4, g.batch, cudaMemcpyHostToDevice,
                                  h->stream));
    } else {
        for (int b = 0; b < g.batch; ++b)
            GD_CUDA(cudaMemcpy2DAsync(depth_dst + (size_t)b * depth_stride_b, (size_t)g.w * 4, depth_m[b], depth_step,
                                      (size_t)g.w * 4, g.h, cudaMemcpyHostToDevice, h->stream));
    }
    return GD_OK;
}

int gd_frontend_fetch(gd_frontend_t* h, uint8_t* const* mask_out, size_t mask_step, gd_keypoint* const* kps, uint8_t* const* desc,
                      int* n_kp)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_REQUIRE(h->results_ready, "no step has run yet");
    GeoMaskCore& g = h->geo;
    OrbCore& o = h->orb;
    const int cap = h->cfg.kp_capacity > 0 ? std::min(h->cfg.kp_capacity, o.plan.kp_capacity) : o.plan.kp_capacity;
    int* hn = h->h_n.as<int>();
    if (mask_out) {
        GD_REQUIRE(mask_step >= (size_t)g.w, "mask_step smaller than a row");
        bool packed = mask_step == (size_t)g.w && g.n_pad == g.n;
        for (int b = 1; b < g.batch && packed; ++b) packed = mask_out[b] == mask_out[b - 1] + g.n;
        if (packed && mask_out[0]) {
            GD_CUDA(cudaMemcpyAsync(mask_out[0], g.mask.p, (size_t)g.batch * g.n, cudaMemcpyDeviceToHost, h->stream));
        } else {
            for (int b = 0; b < g.batch; ++b)
                if (mask_out[b])
                    GD_CUDA(cudaMemcpy2DAsync(mask_out[b], mask_step, g.mask.as<uint8_t>() + (size_t)b * g.n_pad, (size_t)g.w,
                                              (size_t)g.w, g.h, cudaMemcpyDeviceToHost, h->stream));
        }
    }
    // keypoints: copy the fixed capacity (count is data dependent; one synchronisation instead of two)
    for (int b = 0; b < g.batch; ++b) {
        if (kps && kps[b])
            GD_CUDA(cudaMemcpyAsync(kps[b], o.out_kp.as<gd_keypoint>() + (size_t)b * o.plan.kp_capacity, sizeof(gd_keypoint) * cap,
                                    cudaMemcpyDeviceToHost, h->stream));
        if (desc && desc[b])
            GD_CUDA(cudaMemcpyAsync(desc[b], o.out_desc.as<uint8_t>() + (size_t)b * o.plan.kp_capacity * 32, (size_t)32 * cap,
                                    cudaMemcpyDeviceToHost, h->stream));
    }
    GD_CUDA(cudaMemcpyAsync(hn, o.out_n.p, sizeof(int) * g.batch, cudaMemcpyDeviceToHost, h->stream));
    GD_CUDA(cudaMemcpyAsync(hn + g.batch, o.err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    GD_CUDA(cudaStreamSynchronize(h->stream));
    int rc = GD_OK;
    if (hn[g.batch] != 0) {  // reported once, then cleared; counts and (clamped) records are still delivered
        set_error("ORB kernel reported an internal capacity overflow (flags %d): results are truncated", hn[g.batch]);
        GD_CUDA(cudaMemsetAsync(o.err.p, 0, sizeof(int), h->stream));
        rc = GD_EINTERNAL;
    }
    for (int b = 0; b < g.batch; ++b) {
        if (n_kp) n_kp[b] = hn[b];
        if (rc == GD_OK && hn[b] > cap && ((kps && kps[b]) || (desc && desc[b]))) {
            set_error("keypoint capacity %d too small for %d keypoints", cap, hn[b]);
            rc = GD_ECAPACITY;
        }
    }
    return rc;
}

int gd_frontend_step(gd_frontend_t* h, const uint8_t* const* bgr, size_t bgr_step, const float* const* depth_m, size_t depth_step,
                     const float* R, const float* T, const int* pose_valid, uint8_t* const* mask_out, size_t mask_step,
                     gd_keypoint* const* kps, uint8_t* const* desc, int* n_kp)
{
    GD_REQUIRE(h && bgr && depth_m, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GeoMaskCore& g = h->geo;
    const int slot = g.cur_slot();
    GD_TRY(upload_frames(h, g.bgr.as<uint8_t>(), g.depth_slot_ptr(slot), g.depth_stride_b(), bgr, bgr_step, depth_m, depth_step));
    GD_TRY(frontend_compute(h, g.bgr.as<uint8_t>(), g.n_pad * 3, R, T, pose_valid));
    return gd_frontend_fetch(h, mask_out, mask_step, kps, desc, n_kp);
}

int gd_frontend_step_u16(gd_frontend_t* h, const uint8_t* const* bgr, size_t bgr_step, const uint16_t* const* depth_raw,
                         size_t depth_step, const float* R, const float* T, const int* pose_valid, uint8_t* const* mask_out,
                         size_t mask_step, gd_keypoint* const* kps, uint8_t* const* desc, int* n_kp)
{
    GD_REQUIRE(h && bgr && depth_raw, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GeoMaskCore& g = h->geo;
    GD_REQUIRE(bgr_step >= (size_t)g.w * 3 && depth_step >= (size_t)g.w * 2, "step smaller than a row");
    GD_REQUIRE(h->cfg.depth_factor != 0.f, "depth_factor is zero");
    if (!h->raw_depth.p) GD_TRY(h->raw_depth.alloc((size_t)g.batch * g.n_pad * sizeof(uint16_t)));
    const int slot = g.cur_slot();
    // one copy per plane for the whole batch when the caller's frames are densely packed back to back (pinned batch buffers)
    bool packed_bgr = bgr_step == (size_t)g.w * 3 && g.n_pad == g.n, packed_d = depth_step == (size_t)g.w * 2 && g.n_pad == g.n;
    for (int b = 0; b < g.batch; ++b) {
        GD_REQUIRE(bgr[b] && depth_raw[b], "null image pointer");
        if (b > 0) {
            packed_bgr = packed_bgr && bgr[b] == bgr[b - 1] + g.n * 3;
            packed_d = packed_d && depth_raw[b] == depth_raw[b - 1] + g.n;
        }
    }
    if (packed_bgr) {
        GD_CUDA(cudaMemcpyAsync(g.bgr.p, bgr[0], (size_t)g.batch * g.n * 3, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int b = 0; b < g.batch; ++b)
            GD_CUDA(cudaMemcpy2DAsync(g.bgr.as<uint8_t>() + (size_t)b * g.n_pad * 3, (size_t)g.w * 3, bgr[b], bgr_step, (size_t)g.w * 3,
                                      g.h, cudaMemcpyHostToDevice, h->stream));
    }
    if (packed_d) {
        GD_CUDA(cudaMemcpyAsync(h->raw_depth.p, depth_raw[0], (size_t)g.batch * g.n * 2, cudaMemcpyHostToDevice, h->stream));
    } else {
        for (int b = 0; b < g.batch; ++b)
            GD_CUDA(cudaMemcpy2DAsync(h->raw_depth.as<uint16_t>() + (size_t)b * g.n_pad, (size_t)g.w * 2, depth_raw[b], depth_step,
                                      (size_t)g.w * 2, g.h, cudaMemcpyHostToDevice, h->stream));
    }
    const float inv_factor = 1.0f / h->cfg.depth_factor;  // mDepthMapFactor = 1.0f/mDepthMapFactor, Tracking.cc:130-134
    GD_TRY(launch_depth_u16_to_m(h->raw_depth.as<uint16_t>(), g.n_pad, g.depth_slot_ptr(slot), g.depth_stride_b(), g.n, g.batch,
                                 inv_factor, h->stream, &h->stats));
    GD_TRY(frontend_compute(h, g.bgr.as<uint8_t>(), g.n_pad * 3, R, T, pose_valid));
    return gd_frontend_fetch(h, mask_out, mask_step, kps, desc, n_kp);
}

int gd_frontend_stage(gd_frontend_t* h, int slot, const uint8_t* const* bgr, size_t bgr_step, const float* const* depth_m,
                      size_t depth_step)
{
    GD_REQUIRE(h && bgr && depth_m, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GD_REQUIRE(slot >= 0 && slot < h->cfg.staged_slots, "slot out of range");
    GeoMaskCore& g = h->geo;
    const size_t B = (size_t)g.batch;
    GD_TRY(upload_frames(h, h->staged_bgr.as<uint8_t>() + (size_t)slot * B * g.n_pad * 3,
                         h->staged_depth.as<float>() + (size_t)slot * B * g.n_pad, g.n_pad, bgr, bgr_step, depth_m, depth_step));
    GD_CUDA(cudaStreamSynchronize(h->stream));
    return GD_OK;
}

int gd_frontend_step_staged(gd_frontend_t* h, int slot, const float* R, const float* T, const int* pose_valid)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_REQUIRE(slot >= 0 && slot < h->cfg.staged_slots, "slot out of range");
    GeoMaskCore& g = h->geo;
    const size_t B = (size_t)g.batch;
    const int ring = g.cur_slot();
    // the depth image joins the stream's ring (it is read again five frames later); BGR is consumed in place
    GD_CUDA(cudaMemcpy2DAsync(g.depth_slot_ptr(ring), g.depth_stride_b() * 4, h->staged_depth.as<float>() + (size_t)slot * B * g.n_pad,
                              g.n_pad * 4, g.n * 4, g.batch, cudaMemcpyDeviceToDevice, h->stream));
    return frontend_compute(h, h->staged_bgr.as<uint8_t>() + (size_t)slot * B * g.n_pad * 3, g.n_pad * 3, R, T, pose_valid);
}

int gd_frontend_fetch_filtered(gd_frontend_t* h, gd_keypoint* const* kps, uint8_t* const* desc, int* n_kp)
{
    GD_REQUIRE(h && n_kp, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GD_REQUIRE(h->results_ready, "no step has run yet");
    GeoMaskCore& g = h->geo;
    OrbCore& o = h->orb;
    const size_t cap = (size_t)o.plan.kp_capacity, B = (size_t)g.batch;
    if (!h->keep.p) {
        GD_TRY(h->keep.alloc(B * cap));
        GD_TRY(h->filt_kp.alloc(B * cap * sizeof(gd_keypoint)));
        GD_TRY(h->filt_desc.alloc(B * cap * 32));
        GD_TRY(h->filt_n.alloc(B * sizeof(int)));
    }
    GD_TRY(launch_erode_filter(g.mask.as<uint8_t>(), g.n_pad, g.w, g.h, g.batch, o.out_kp.as<gd_keypoint>(), cap, o.out_n.as<int>(), 0,
                               h->keep.as<uint8_t>(), h->stream, &h->stats));
    GD_TRY(launch_compact_keypoints(o.out_kp.as<gd_keypoint>(), o.out_desc.as<uint8_t>(), h->keep.as<uint8_t>(), cap, g.batch,
                                    o.out_n.as<int>(), h->filt_kp.as<gd_keypoint>(), h->filt_desc.as<uint8_t>(), h->filt_n.as<int>(),
                                    h->stream, &h->stats));
    const int ucap = h->cfg.kp_capacity > 0 ? std::min(h->cfg.kp_capacity, o.plan.kp_capacity) : o.plan.kp_capacity;
    int* hn = h->h_n.as<int>();
    for (int b = 0; b < g.batch; ++b) {
        if (kps && kps[b])
            GD_CUDA(cudaMemcpyAsync(kps[b], h->filt_kp.as<gd_keypoint>() + (size_t)b * cap, sizeof(gd_keypoint) * ucap,
                                    cudaMemcpyDeviceToHost, h->stream));
        if (desc && desc[b])
            GD_CUDA(cudaMemcpyAsync(desc[b], h->filt_desc.as<uint8_t>() + (size_t)b * cap * 32, (size_t)32 * ucap, cudaMemcpyDeviceToHost,
                                    h->stream));
    }
    GD_CUDA(cudaMemcpyAsync(hn, h->filt_n.p, sizeof(int) * g.batch, cudaMemcpyDeviceToHost, h->stream));
    GD_CUDA(cudaStreamSynchronize(h->stream));
    for (int b = 0; b < g.batch; ++b) n_kp[b] = hn[b];
    h->filtered_ready = true;
    return GD_OK;
}

int gd_frontend_fetch_stereo_grid(gd_frontend_t* h, float bf, float* const* depth, float* const* uright, int* const* cell_start,
                                  int* const* cell_items)
{
    return gd_frontend_fetch_stereo_grid_un(h, bf, depth, uright, cell_start, cell_items, nullptr);
}

int gd_frontend_fetch_stereo_grid_un(gd_frontend_t* h, float bf, float* const* depth, float* const* uright, int* const* cell_start,
                                     int* const* cell_items, float* const* keys_un)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_REQUIRE(h->filtered_ready, "call gd_frontend_fetch_filtered first");
    GeoMaskCore& g = h->geo;
    OrbCore& o = h->orb;
    const size_t cap = (size_t)o.plan.kp_capacity, B = (size_t)g.batch;
    const int ncell = 64 * 48 + 1;
    if (!h->sg_depth.p) {
        GD_TRY(h->sg_depth.alloc(B * cap * sizeof(float)));
        GD_TRY(h->sg_uright.alloc(B * cap * sizeof(float)));
        GD_TRY(h->sg_start.alloc(B * ncell * sizeof(int)));
        GD_TRY(h->sg_items.alloc(B * cap * sizeof(int)));
        GD_TRY(h->sg_un.alloc(B * cap * sizeof(float2)));
    }
    UndistortArgs und;
    make_undistort_args(h->cfg.K, h->cfg.dist, h->cfg.ndist, &und);
    const int cur = (g.frames - 1) % GD_RING;  // depth image of the newest frame
    GD_TRY(launch_stereo_grid(g.depth_slot_ptr(cur), g.depth_stride_b(), g.w, g.h, g.batch, h->filt_kp.as<gd_keypoint>(), cap,
                              h->filt_n.as<int>(), bf, und, h->sg_depth.as<float>(), h->sg_uright.as<float>(), h->sg_start.as<int>(),
                              h->sg_items.as<int>(), h->sg_un.as<float2>(), h->stream, &h->stats));
    const int ucap = h->cfg.kp_capacity > 0 ? std::min(h->cfg.kp_capacity, o.plan.kp_capacity) : o.plan.kp_capacity;
    for (int b = 0; b < g.batch; ++b) {
        if (depth && depth[b])
            GD_CUDA(cudaMemcpyAsync(depth[b], h->sg_depth.as<float>() + (size_t)b * cap, sizeof(float) * ucap, cudaMemcpyDeviceToHost, h->stream));
        if (uright && uright[b])
            GD_CUDA(cudaMemcpyAsync(uright[b], h->sg_uright.as<float>() + (size_t)b * cap, sizeof(float) * ucap, cudaMemcpyDeviceToHost, h->stream));
        if (cell_start && cell_start[b])
            GD_CUDA(cudaMemcpyAsync(cell_start[b], h->sg_start.as<int>() + (size_t)b * ncell, sizeof(int) * ncell, cudaMemcpyDeviceToHost, h->stream));
        if (cell_items && cell_items[b])
            GD_CUDA(cudaMemcpyAsync(cell_items[b], h->sg_items.as<int>() + (size_t)b * cap, sizeof(int) * ucap, cudaMemcpyDeviceToHost, h->stream));
        if (keys_un && keys_un[b])
            GD_CUDA(cudaMemcpyAsync(keys_un[b], h->sg_un.as<float2>() + (size_t)b * cap, sizeof(float2) * ucap, cudaMemcpyDeviceToHost, h->stream));
    }
    GD_CUDA(cudaStreamSynchronize(h->stream));
    return GD_OK;
}

int gd_stage_erode_filter(int device, const uint8_t* mask, int w, int h, const gd_keypoint* kps, int n, uint8_t* keep)
{
    GD_REQUIRE(mask && kps && keep && w > 0 && h > 0 && n >= 0, "bad argument");
    GD_TRY(select_device(device));
    if (n == 0) return GD_OK;
    DevBuf dm, dk, dkeep;
    GD_TRY(dm.alloc((size_t)w * h));
    GD_TRY(dk.alloc((size_t)n * sizeof(gd_keypoint)));
    GD_TRY(dkeep.alloc((size_t)n));
    GD_CUDA(cudaMemcpy(dm.p, mask, (size_t)w * h, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dk.p, kps, (size_t)n * sizeof(gd_keypoint), cudaMemcpyHostToDevice));
    GD_TRY(launch_erode_filter(dm.as<uint8_t>(), 0, w, h, 1, dk.as<gd_keypoint>(), (size_t)n, nullptr, n, dkeep.as<uint8_t>(), 0, nullptr));
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(keep, dkeep.p, (size_t)n, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_frontend_set_pose_hook(gd_frontend_t* h, gd_pose_hook_fn hook, void* user)
{
    GD_REQUIRE(h, "null handle");
    GD_REQUIRE(!hook || h->geo.getrt, "the pose hook needs the GetRt stage (gd_frontend_config.getrt = 1)");
    h->pose_hook = hook;
    h->pose_hook_user = user;
    return GD_OK;
}

int gd_frontend_fetch_getrt(gd_frontend_t* h, float* const* object_points, float* const* image_pixels, int* n_points)
{
    GD_REQUIRE(h && n_points, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GeoMaskCore& g = h->geo;
    GD_REQUIRE(g.getrt, "GetRt stage not enabled (gd_frontend_config.getrt = 1)");
    GD_CUDA(cudaStreamSynchronize(h->stream));  // the step joined the GetRt chain into the main stream
    if (h->rt_stream) GD_CUDA(cudaStreamSynchronize(h->rt_stream));
    for (int b = 0; b < g.batch; ++b) n_points[b] = 0;
    if (!h->getrt_points_ready) return GD_OK;  // warm-up: no buffered pair yet
    if (g.getrt->host_err() != 0) {
        set_error("GetRt stage: capacity overflow of the selection lists (flags %d)", g.getrt->host_err());
        return GD_EINTERNAL;
    }
    for (int b = 0; b < g.batch; ++b) {
        const int n = g.getrt->host_cnt()[b];
        n_points[b] = n;
        if (object_points && object_points[b]) std::memcpy(object_points[b], g.getrt->host_obj(b), sizeof(float) * 3 * n);
        if (image_pixels && image_pixels[b]) std::memcpy(image_pixels[b], g.getrt->host_pix(b), sizeof(float) * 2 * n);
    }
    return GD_OK;
}

int gd_frontend_sync(gd_frontend_t* h)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_CUDA(cudaStreamSynchronize(h->stream));
    return GD_OK;
}

int gd_frontend_timer_begin(gd_frontend_t* h)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_CUDA(cudaEventRecord(h->ev0, h->stream));
    return GD_OK;
}

int gd_frontend_timer_end(gd_frontend_t* h, float* ms)
{
    GD_REQUIRE(h && ms, "null argument");
    GD_TRY(select_device(h->cfg.device));
    GD_CUDA(cudaEventRecord(h->ev1, h->stream));
    GD_CUDA(cudaEventSynchronize(h->ev1));
    GD_CUDA(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return GD_OK;
}

int gd_frontend_launch_count(gd_frontend_t* h, long long* launches)
{
    GD_REQUIRE(h && launches, "null argument");
    *launches = h->stats.launches;
    return GD_OK;
}

int gd_frontend_profile(gd_frontend_t* h, int enable)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    GD_CUDA(cudaStreamSynchronize(h->stream));
    h->stats.profiling = enable != 0;
    if (enable) h->stats.fam.clear();
    return GD_OK;
}

int gd_frontend_profile_read(gd_frontend_t* h, int max_entries, const char** names, float* ms, long long* launches, int* n_entries)
{
    GD_REQUIRE(h && n_entries, "null argument");
    const int n = (int)h->stats.fam.size();
    *n_entries = n;
    for (int i = 0; i < n && i < max_entries; ++i) {
        if (names) names[i] = h->stats.fam[i].name;
        if (ms) ms[i] = h->stats.fam[i].ms;
        if (launches) launches[i] = h->stats.fam[i].launches;
    }
    return GD_OK;
}

int gd_frontend_debug_fetch(gd_frontend_t* h, int what, int stream, void* dst, size_t dst_bytes)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    return h->geo.debug_fetch(what, stream, dst, dst_bytes);
}

int gd_frontend_flush_l2(gd_frontend_t* h)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->cfg.device));
    const size_t bytes = (size_t)256 << 20;  // > 126 MB L2
    if (!h->l2_scratch.p) GD_TRY(h->l2_scratch.alloc(bytes));
    GD_CUDA(cudaMemsetAsync(h->l2_scratch.p, 0x5a, bytes, h->stream));
    return GD_OK;
}

}  // extern "C"
