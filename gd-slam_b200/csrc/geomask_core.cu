// geomask_core.cu — per-frame sequencing of the GeoMaskMaker kernels over a batch of streams.
#include "geomask_core.cuh"

#include <cmath>
#include <cstdlib>

namespace gd {

// cv::undistortPoints(pts, K, D, noArray, P=K) on the integer pixel grid (GeoMaskMaker.cc:56-69).
// 5 fixed-point iterations, f64, like OpenCV's default criteria.  Only used when D != 0 (TUM1/TUM2);
// TUM3 (D = 0) maps the grid onto itself exactly and takes the LUT-free path.
static void build_undistort_lut(const float K[9], const float* d, int nd, int w, int h, std::vector<float>& lut)
{
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    double k[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < nd && i < 5; ++i) k[i] = d[i];
    lut.resize((size_t)w * h * 2);
    for (int v = 0; v < h; ++v)
        for (int u = 0; u < w; ++u) {
            double x = (u - cx) / fx, y = (v - cy) / fy;
            const double x0 = x, y0 = y;
            for (int it = 0; it < 5; ++it) {
                const double r2 = x * x + y * y;
                double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
                if (icdist < 0) {
                    x = x0;
                    y = y0;
                    break;
                }
                const double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
                const double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
                x = (x0 - dX) * icdist;
                y = (y0 - dY) * icdist;
            }
            lut[2 * ((size_t)v * w + u)] = (float)(x * fx + cx);
            lut[2 * ((size_t)v * w + u) + 1] = (float)(y * fy + cy);
        }
}

static int dist_coef_in_range(int n) { return n < 0 ? 0 : (n > 5 ? 5 : n); }

int GeoMaskCore::init(const float K_[9], const float* dist_coef, int ndist, int width, int height, int device_,
                      int batch_, cudaStream_t s, LaunchStats* st)
{
    GD_REQUIRE(width >= 16 && height >= 16 && batch_ >= 1, "bad size / batch");
    GD_TRY(select_device(device_));
    device = device_;
    batch = batch_;
    w = width;
    h = height;
    n = (size_t)w * h;
    n_pad = align_up(n, 64);
    stats = st;
    std::memcpy(K, K_, sizeof(K));
    make_cam_const(K, &cam);
    ndist = dist_coef_in_range(ndist);
    for (int i = 0; i < ndist && dist_coef; ++i) this->dist_coef[i] = dist_coef[i];
    if (!dist_coef) ndist = 0;
    this->ndist = ndist;
    for (int i = 0; i < GD_RING; ++i) feat_frame[i] = -1;
    GD_TRY(fb_make_plan(w, h, 0.5, 3, 3, 5, 1.2, 15, &plan));  // GeoMaskMaker.cc:165
    if (s) {
        stream = s;
    } else {
        GD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    const size_t B = (size_t)batch;
    GD_TRY(bgr.alloc(B * n_pad * 3));
    GD_TRY(gray.alloc(B * n_pad));
    GD_TRY(depth.alloc(B * GD_RING * n_pad * sizeof(float)));
    GD_TRY(edge.alloc(B * GD_RING * n_pad));
    GD_TRY(R.alloc(B * GD_RING * plan.r_floats * sizeof(float)));
    GD_TRY(scratchI.alloc(B * plan.i_floats * sizeof(float)));
    GD_TRY(flowA.alloc(B * plan.f_float2 * sizeof(float2)));
    GD_TRY(flowB.alloc(B * plan.f_float2 * sizeof(float2)));
    {  // GD_FLOW_FUSED=1 selects the older single-kernel flow iteration; default (0 / unset): split form (matrices + box/solve)
        const char* e = std::getenv("GD_FLOW_FUSED");
        split_flow = !(e && std::atoi(e) != 0);
    }
    if (split_flow) {
        // M scratch of the split flow form.  GD_M_L2_MB=<n> processes the streams in groups whose M fits n MB (so that it
        // could stay L2 resident between the two kernels); measured on B200 at batch 32: 24/48/96 MB groups are 12/5/2 %
        // SLOWER than one launch over all streams (smaller grids, more tails), so the default is no grouping.
        const char* e = std::getenv("GD_M_L2_MB");
        const size_t one = plan.m_floats * sizeof(float);
        const size_t budget = e ? (size_t)std::max(1, std::atoi(e)) << 20 : B * one;
        GD_TRY(Mbuf.alloc(std::max(one, std::min(B * one, budget / one * one))));
        const bool whole = Mbuf.bytes >= B * one;
        if (whole) GD_TRY(Mbuf2.alloc(Mbuf.bytes));
        GD_TRY(fb_prepare_flow_buffers(plan, batch, Mbuf.as<float>(), whole ? Mbuf2.as<float>() : nullptr, Mbuf.bytes, &flow_bufs));
    }
    GD_TRY(keys.alloc(B * n_pad * sizeof(unsigned long long)));
    GD_TRY(minmax.alloc(B * GD_MM_WORDS * sizeof(unsigned)));
    keyfmt = make_key_format(n);
    GD_TRY(poses.alloc(B * sizeof(PoseDev)));
    GD_TRY(mask.alloc(B * n_pad));
    GD_TRY(h_poses.alloc(POSE_SLOTS * B * sizeof(PoseDev)));
    for (int i = 0; i < POSE_SLOTS; ++i) GD_CUDA(cudaEventCreateWithFlags(&pose_ev[i], cudaEventDisableTiming));
    GD_CUDA(cudaMemsetAsync(keys.p, 0, keys.bytes, stream));
    GD_CUDA(cudaMemsetAsync(depth.p, 0, depth.bytes, stream));
    GD_CUDA(cudaMemsetAsync(edge.p, 0, edge.bytes, stream));
    bool any = false;
    for (int i = 0; i < ndist && dist_coef; ++i) any = any || dist_coef[i] != 0.f;
    if (any) {
        std::vector<float> l;
        build_undistort_lut(K, dist_coef, ndist, w, h, l);
        GD_TRY(lut.alloc(l.size() * sizeof(float)));
        GD_CUDA(cudaMemcpyAsync(lut.p, l.data(), l.size() * sizeof(float), cudaMemcpyHostToDevice, stream));
        GD_CUDA(cudaStreamSynchronize(stream));
        has_lut = true;
    }
    GD_CUDA(cudaStreamSynchronize(stream));
    return GD_OK;
}

GeoMaskCore::~GeoMaskCore()
{
    for (int i = 0; i < POSE_SLOTS; ++i)
        if (pose_ev[i]) cudaEventDestroy(pose_ev[i]);
    if (own_stream && stream) cudaStreamDestroy(stream);
}

int GeoMaskCore::push_resident(bool gray_done)
{
    const int slot = frames % GD_RING;
    // graphs only once every path has run un-captured (lazy set-up) and the ring is in steady state
    const bool replay = frames >= 2 * GD_RING && !edge_stream;
    const int rc = replay ? push_graphs.run((unsigned long long)slot * 2 + (gray_done ? 1 : 0), stream, stats,
                                            [&] { return enqueue_push(slot, gray_done); })
                          : enqueue_push(slot, gray_done);
    if (rc != GD_OK) return rc;
    if (getrt) feat_frame[slot] = frames;
    frames += 1;
    return GD_OK;
}

int GeoMaskCore::enqueue_push(int slot, bool gray_done)
{
    PdlScope pdl_scope(batch);
    // K0: gray for the flow (the batched front-end computes it together with the ORB gray)
    if (!gray_done)
        GD_TRY(launch_gray(bgr.as<uint8_t>(), (size_t)w * 3, n_pad * 3, w, h, batch, gray.as<uint8_t>(), n_pad, nullptr, 0, 0, 0,
                           stream, stats));
    // K1a: blur + resample + polynomial expansion, all levels, into the ring slot
    GD_TRY(fb_launch_pyramid_polyexp(plan, gray.as<uint8_t>(), n_pad, batch, scratchI.as<float>(), plan.i_floats,
                                     R.as<float>() + (size_t)slot * plan.r_floats, (size_t)GD_RING * plan.r_floats, stream, stats));
    // K2a: depth edges of the new depth image
    GD_TRY(launch_depth_edge(depth_slot_ptr(slot), depth_stride_b(), w, h, batch, cam,
                             edge.as<uint8_t>() + (size_t)slot * n_pad, (size_t)GD_RING * n_pad,
                             edge_stream ? edge_stream : stream, stats));
    // GetRt (when enabled): cv::ORB features of the new frame into the slot's feature cache
    if (getrt) {
        getrt->stream = getrt_stream ? getrt_stream : stream;
        getrt->stats = stats;
        GD_TRY(getrt->enqueue_features(gray.as<uint8_t>(), n_pad, slot));
    }
    return GD_OK;
}

int GeoMaskCore::enable_getrt()
{
    if (getrt) return GD_OK;
    GD_TRY(select_device(device));
    std::unique_ptr<GetRtCore> c(new (std::nothrow) GetRtCore());
    if (!c) return GD_ENOMEM;
    GD_TRY(c->init(K, ndist ? dist_coef : nullptr, ndist, w, h, device, batch, GD_RING));
    getrt = std::move(c);
    // graphs captured before this point do not contain the feature launches
    push_graphs.enabled = false;
    return GD_OK;
}

bool GeoMaskCore::getrt_pair_ready() const
{
    if (!getrt || frames < GD_RING) return false;
    const int cur = (frames - 1) % GD_RING, ref = (frames - GD_RING) % GD_RING;
    return feat_frame[cur] == (long long)frames - 1 && feat_frame[ref] == (long long)frames - GD_RING;
}

int GeoMaskCore::enqueue_getrt_match()
{
    GD_REQUIRE(getrt, "GetRt stage not enabled");
    GD_REQUIRE(getrt_pair_ready(), "features of the buffered pair are not available (enable GetRt before pushing the frames)");
    const int cur = (frames - 1) % GD_RING, ref = (frames - GD_RING) % GD_RING;
    getrt->stream = getrt_stream ? getrt_stream : stream;
    getrt->stats = stats;
    GD_TRY(getrt->enqueue_match(ref, cur, depth_slot_ptr(ref), depth_stride_b()));
    return getrt->enqueue_fetch();
}

int GeoMaskCore::compute_mask(const float* Rm, const float* Tm, const int* pose_valid)
{
    GD_TRY(upload_poses(Rm, Tm, pose_valid, frames));
    if (frames < 2 * GD_RING) return enqueue_mask();
    last_cur_slot = (frames - 1) % GD_RING;  // host state enqueue_mask() sets; a replay does not run it
    last_ref_slot = (frames - GD_RING) % GD_RING;
    return mask_graphs.run((unsigned long long)(frames % GD_RING), stream, stats, [&] { return enqueue_mask(); });
}

int GeoMaskCore::upload_poses(const float* Rm, const float* Tm, const int* pose_valid, int frames_pushed)
{
    const bool started = frames_pushed >= GD_RING;  // start_flag, GeoMaskMaker.cc:419-428
    const int slot = pose_slot;
    pose_slot = (pose_slot + 1) % POSE_SLOTS;
    if (pose_pending[slot]) GD_CUDA(cudaEventSynchronize(pose_ev[slot]));  // the copy that last read this slot has run
    PoseDev* hp = h_poses.as<PoseDev>() + (size_t)slot * batch;
    // scatter-key epoch of this step (KeyFormat): 1 .. epoch_max, the key image is cleared once per wrap-around
    epoch = epoch % keyfmt.epoch_max + 1;
    if (epoch == 1 && epoch_used) GD_CUDA(cudaMemsetAsync(keys.p, 0, keys.bytes, stream));
    epoch_used = true;
    for (int b = 0; b < batch; ++b) {
        const int valid = started && (!pose_valid || pose_valid[b]) ? 1 : 0;
        static const float I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, Z3[3] = {0, 0, 0};
        make_pose(K, Rm ? Rm + 9 * b : I3, Tm ? Tm + 3 * b : Z3, valid, epoch, hp + b);
    }
    GD_CUDA(cudaMemcpyAsync(poses.p, hp, sizeof(PoseDev) * batch, cudaMemcpyHostToDevice, stream));
    GD_CUDA(cudaEventRecord(pose_ev[slot], stream));
    pose_pending[slot] = true;
    return GD_OK;
}

// device half: everything GetNoGMMmask enqueues on the stream (capturable into a CUDA graph)
int GeoMaskCore::enqueue_mask()
{
    PdlScope pdl_scope(batch);
    GD_TRY(enqueue_flow());
    return enqueue_mask_tail();
}

int GeoMaskCore::enqueue_flow()
{
    PdlScope pdl_scope(batch);
    const bool started = frames >= GD_RING;
    if (!started) {
        last_flow = nullptr;
        return GD_OK;
    }
    const int cur = (frames - 1) % GD_RING;
    const int ref = (frames - GD_RING) % GD_RING;  // frame t-5
    last_ref_slot = ref;
    last_cur_slot = cur;
    const size_t rs = (size_t)GD_RING * plan.r_floats;
    return fb_launch_flow(plan, R.as<float>() + (size_t)ref * plan.r_floats, R.as<float>() + (size_t)cur * plan.r_floats, rs,
                          batch, flowA.as<float2>(), flowB.as<float2>(), plan.f_float2, split_flow ? &flow_bufs : nullptr, &last_flow,
                          stream, stats);
}

int GeoMaskCore::enqueue_mask_tail()
{
    PdlScope pdl_scope(batch);
    const bool started = frames >= GD_RING;
    if (!started) {  // warm-up: all-ones mask (:171-175)
        return launch_fill_u8(mask.as<uint8_t>(), (size_t)batch * n_pad, 1, stream, stats);
    }
    const int cur = last_cur_slot, ref = last_ref_slot;
    GD_TRY(launch_mahalanobis(last_flow, plan.f_float2, depth_slot_ptr(ref), depth_slot_ptr(cur), depth_stride_b(),
                              edge.as<uint8_t>() + (size_t)ref * n_pad, edge.as<uint8_t>() + (size_t)cur * n_pad,
                              (size_t)GD_RING * n_pad, has_lut ? lut.as<float2>() : nullptr, w, h, batch, cam,
                              poses.as<PoseDev>(), keyfmt, keys.as<unsigned long long>(), n_pad, stream, stats));
    bool clustered = false;
    GD_TRY(launch_minmax_mask_cluster(keys.as<unsigned long long>(), n_pad, (int)n, batch, poses.as<PoseDev>(), keyfmt,
                                      minmax.as<unsigned>(), mask.as<uint8_t>(), n_pad, stream, stats, &clustered));
    if (clustered) return GD_OK;
    GD_TRY(launch_minmax_reset(minmax.as<unsigned>(), batch, stream));
    GD_TRY(launch_minmax(keys.as<unsigned long long>(), n_pad, (int)n, batch, poses.as<PoseDev>(), keyfmt, minmax.as<unsigned>(),
                         stream, stats));
    GD_TRY(launch_normalize_mask(keys.as<unsigned long long>(), n_pad, (int)n, batch, minmax.as<unsigned>(), poses.as<PoseDev>(),
                                 keyfmt, mask.as<uint8_t>(), n_pad, stream, stats));
    return GD_OK;
}

int GeoMaskCore::debug_fetch(int what, int b, void* dst, size_t dst_bytes)
{
    GD_REQUIRE(b >= 0 && b < batch && dst, "bad stream index / dst");
    const void* src = nullptr;
    size_t bytes = 0;
    switch (what) {
        case GD_DBG_FLOW:
            GD_REQUIRE(last_flow, "no flow computed yet");
            src = last_flow + (size_t)b * plan.f_float2;
            bytes = n * sizeof(float2);
            break;
        case GD_DBG_DIST:
            // resolved on demand from the key image of the last step (the per-frame path only writes the mask)
            GD_REQUIRE(last_cur_slot >= 0, "no pair evaluated yet");
            if (!dist.p) GD_TRY(dist.alloc((size_t)batch * n_pad * sizeof(float)));
            GD_TRY(launch_resolve_dist(keys.as<unsigned long long>(), n_pad, (int)n, batch, poses.as<PoseDev>(), keyfmt,
                                       dist.as<float>(), n_pad, stream));
            src = dist.as<float>() + (size_t)b * n_pad;
            bytes = n * sizeof(float);
            break;
        case GD_DBG_EDGE_REF:
        case GD_DBG_EDGE_CUR: {
            GD_REQUIRE(last_cur_slot >= 0, "no pair evaluated yet");
            const int slot = what == GD_DBG_EDGE_REF ? last_ref_slot : last_cur_slot;
            src = edge.as<uint8_t>() + ((size_t)b * GD_RING + slot) * n_pad;
            bytes = n;
            break;
        }
        case GD_DBG_LUT:
            GD_REQUIRE(has_lut, "no LUT: the distortion coefficients are zero (identity path)");
            src = lut.p;
            bytes = n * sizeof(float2);
            break;
        case GD_DBG_GRAY_CUR:
            src = gray.as<uint8_t>() + (size_t)b * n_pad;
            bytes = n;
            break;
        case GD_DBG_MINMAX: {
            GD_REQUIRE(dst_bytes >= 2 * sizeof(float), "dst too small");
            unsigned bits[GD_MM_WORDS];
            GD_CUDA(cudaStreamSynchronize(stream));
            GD_CUDA(cudaMemcpy(bits, minmax.as<unsigned>() + GD_MM_WORDS * b, sizeof(bits), cudaMemcpyDeviceToHost));
            unsigned mn = bits[0], mx = ~bits[1];
            if (bits[2] == 0u) mn = mx = 0x7FC00000u;  // NaN at pixel 0: cv's scan returns NaN for both
            std::memcpy(dst, &mn, 4);
            std::memcpy((char*)dst + 4, &mx, 4);
            return GD_OK;
        }
        default:
            set_error("unknown debug selector %d", what);
            return GD_EINVAL;
    }
    GD_REQUIRE(dst_bytes >= bytes, "dst too small");
    GD_CUDA(cudaStreamSynchronize(stream));
    GD_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
    return GD_OK;
}

}  // namespace gd
