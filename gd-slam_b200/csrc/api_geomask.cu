// api_geomask.cu — C ABI: gd_geomask_* (GeoMaskMaker drop-in) and the gd_stage_* parity harness for its kernels.
#include "geomask_core.cuh"

struct gd_geomask {
    gd::GeoMaskCore core;
    gd::LaunchStats stats;
};

using namespace gd;

extern "C" {

int gd_geomask_create(gd_geomask_t** out, const float K[9], const float* dist, int ndist, float depth_factor, int width,
                      int height, int device, int batch)
{
    (void)depth_factor;  // the reference stores it but never uses it on this path (depth arrives in metres)
    GD_REQUIRE(out && K, "null argument");
    *out = nullptr;
    gd_geomask* h = new (std::nothrow) gd_geomask();
    if (!h) return GD_ENOMEM;
    int r = h->core.init(K, dist, ndist, width, height, device, batch, nullptr, &h->stats);
    if (r != GD_OK) {
        delete h;
        return r;
    }
    h->core.push_graphs.enabled = h->core.mask_graphs.enabled = GraphCache::env_default(batch);
    *out = h;
    return GD_OK;
}

void gd_geomask_destroy(gd_geomask_t* h)
{
    if (!h) return;
    cudaSetDevice(h->core.device);
    if (h->core.stream) cudaStreamSynchronize(h->core.stream);
    delete h;
}

int gd_geomask_push(gd_geomask_t* h, const uint8_t* const* bgr, size_t bgr_step, const float* const* depth_m,
                    size_t depth_step)
{
    GD_REQUIRE(h && bgr && depth_m, "null argument");
    GeoMaskCore& c = h->core;
    GD_TRY(select_device(c.device));
    GD_REQUIRE(bgr_step >= (size_t)c.w * 3 && depth_step >= (size_t)c.w * sizeof(float), "step smaller than a row");
    const int slot = c.cur_slot();
    for (int b = 0; b < c.batch; ++b) {
        GD_REQUIRE(bgr[b] && depth_m[b], "null image pointer");
        GD_CUDA(cudaMemcpy2DAsync(c.bgr.as<uint8_t>() + (size_t)b * c.n_pad * 3, (size_t)c.w * 3, bgr[b], bgr_step,
                                  (size_t)c.w * 3, c.h, cudaMemcpyHostToDevice, c.stream));
        GD_CUDA(cudaMemcpy2DAsync(c.depth_slot_ptr(slot) + (size_t)b * c.depth_stride_b(), (size_t)c.w * sizeof(float),
                                  depth_m[b], depth_step, (size_t)c.w * sizeof(float), c.h, cudaMemcpyHostToDevice, c.stream));
    }
    GD_TRY(c.push_resident());
    // the caller may reuse its buffers as soon as we return (the reference deep-copies inside AddNewImage)
    GD_CUDA(cudaStreamSynchronize(c.stream));
    return GD_OK;
}

int gd_geomask_mask(gd_geomask_t* h, const float* R, const float* T, const int* pose_valid, uint8_t* const* mask_out,
                    size_t mask_step)
{
    GD_REQUIRE(h && mask_out, "null argument");
    GeoMaskCore& c = h->core;
    GD_TRY(select_device(c.device));
    GD_REQUIRE(mask_step >= (size_t)c.w, "mask_step smaller than a row");
    GD_TRY(c.compute_mask(R, T, pose_valid));
    for (int b = 0; b < c.batch; ++b) {
        GD_REQUIRE(mask_out[b], "null mask pointer");
        GD_CUDA(cudaMemcpy2DAsync(mask_out[b], mask_step, c.mask.as<uint8_t>() + (size_t)b * c.n_pad, (size_t)c.w, (size_t)c.w,
                                  c.h, cudaMemcpyDeviceToHost, c.stream));
    }
    GD_CUDA(cudaStreamSynchronize(c.stream));
    return GD_OK;
}

int gd_geomask_frames(const gd_geomask_t* h) { return h ? h->core.frames : GD_EINVAL; }

int gd_geomask_enable_getrt(gd_geomask_t* h)
{
    GD_REQUIRE(h, "null handle");
    return h->core.enable_getrt();
}

int gd_geomask_getrt_points(gd_geomask_t* h, float* const* object_points, float* const* image_pixels, int* n_points)
{
    GD_REQUIRE(h && n_points, "null argument");
    GeoMaskCore& c = h->core;
    GD_TRY(select_device(c.device));
    GD_REQUIRE(c.getrt, "GetRt stage not enabled (gd_geomask_enable_getrt)");
    for (int b = 0; b < c.batch; ++b) n_points[b] = 0;
    if (c.frames < GD_RING) return GD_OK;  // fewer than six frames: no pair (the reference's warm-up, :171-175)
    GD_TRY(c.enqueue_getrt_match());
    GD_CUDA(cudaStreamSynchronize(c.stream));
    if (c.getrt->host_err() != 0) {
        set_error("GetRt stage: capacity overflow of the selection lists (flags %d)", c.getrt->host_err());
        return GD_EINTERNAL;
    }
    for (int b = 0; b < c.batch; ++b) {
        const int n = c.getrt->host_cnt()[b];
        n_points[b] = n;
        if (object_points && object_points[b]) std::memcpy(object_points[b], c.getrt->host_obj(b), sizeof(float) * 3 * n);
        if (image_pixels && image_pixels[b]) std::memcpy(image_pixels[b], c.getrt->host_pix(b), sizeof(float) * 2 * n);
    }
    return GD_OK;
}

int gd_geomask_debug_fetch(gd_geomask_t* h, int what, int stream, void* dst, size_t dst_bytes)
{
    GD_REQUIRE(h, "null handle");
    GD_TRY(select_device(h->core.device));
    return h->core.debug_fetch(what, stream, dst, dst_bytes);
}

// ----------------------------------------------------------------------------------------------- stages
int gd_stage_gray(int device, const uint8_t* bgr, size_t bgr_step, int w, int h, int order, uint8_t* gray)
{
    GD_REQUIRE(bgr && gray && w > 0 && h > 0 && bgr_step >= (size_t)w * 3, "bad argument");
    GD_TRY(select_device(device));
    DevBuf in, out;
    GD_TRY(in.alloc(bgr_step * h));
    GD_TRY(out.alloc((size_t)w * h));
    GD_CUDA(cudaMemcpy(in.p, bgr, bgr_step * h, cudaMemcpyHostToDevice));
    if (order == 0)
        GD_TRY(launch_gray(in.as<uint8_t>(), bgr_step, 0, w, h, 1, out.as<uint8_t>(), 0, nullptr, 0, 0, 0, 0, nullptr));
    else
        GD_TRY(launch_gray(in.as<uint8_t>(), bgr_step, 0, w, h, 1, nullptr, 0, out.as<uint8_t>(), order, (size_t)w, 0, 0, nullptr));
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(gray, out.p, (size_t)w * h, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_depth_edge(int device, const float* depth_m, int w, int h, const float K[9], uint8_t* edge)
{
    GD_REQUIRE(depth_m && edge && K && w > 2 && h > 2, "bad argument");
    GD_TRY(select_device(device));
    const size_t n = (size_t)w * h;
    DevBuf d, e;
    GD_TRY(d.alloc(n * sizeof(float)));
    GD_TRY(e.alloc(n));
    GD_CUDA(cudaMemcpy(d.p, depth_m, n * sizeof(float), cudaMemcpyHostToDevice));
    CamConst cam;
    make_cam_const(K, &cam);
    GD_TRY(launch_depth_edge(d.as<float>(), 0, w, h, 1, cam, e.as<uint8_t>(), 0, 0, nullptr));
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(edge, e.p, n, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_mahalanobis(int device, const float* flow, const float* depth_ref, const float* depth_cur,
                         const uint8_t* edge_ref, const uint8_t* edge_cur, const float* lut, int w, int h, const float K[9],
                         const float R[9], const float T[3], float* dist, uint8_t* mask, float* minmax)
{
    GD_REQUIRE(flow && depth_ref && depth_cur && edge_ref && edge_cur && K && R && T && w > 0 && h > 0, "bad argument");
    GD_TRY(select_device(device));
    const size_t n = (size_t)w * h, np = align_up(n, 64);
    DevBuf dflow, dr, dc, er, ec, dl, keys, mm, pose, dmask, ddist;
    GD_TRY(dflow.alloc(n * 8));
    GD_TRY(dr.alloc(n * 4));
    GD_TRY(dc.alloc(n * 4));
    GD_TRY(er.alloc(n));
    GD_TRY(ec.alloc(n));
    GD_TRY(keys.alloc(np * 8));
    GD_TRY(mm.alloc(GD_MM_WORDS * 4));
    GD_TRY(pose.alloc(sizeof(PoseDev)));
    GD_TRY(dmask.alloc(np));
    GD_TRY(ddist.alloc(np * 4));
    GD_CUDA(cudaMemcpy(dflow.p, flow, n * 8, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dr.p, depth_ref, n * 4, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dc.p, depth_cur, n * 4, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(er.p, edge_ref, n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(ec.p, edge_cur, n, cudaMemcpyHostToDevice));
    if (lut) {
        GD_TRY(dl.alloc(n * 8));
        GD_CUDA(cudaMemcpy(dl.p, lut, n * 8, cudaMemcpyHostToDevice));
    }
    GD_CUDA(cudaMemset(keys.p, 0, np * 8));
    CamConst cam;
    make_cam_const(K, &cam);
    PoseDev p;
    make_pose(K, R, T, 1, 1, &p);
    const KeyFormat kf = make_key_format(n);
    GD_CUDA(cudaMemcpy(pose.p, &p, sizeof(p), cudaMemcpyHostToDevice));
    GD_TRY(launch_mahalanobis(dflow.as<float2>(), 0, dr.as<float>(), dc.as<float>(), 0, er.as<uint8_t>(), ec.as<uint8_t>(), 0,
                              lut ? dl.as<float2>() : nullptr, w, h, 1, cam, pose.as<PoseDev>(), kf,
                              keys.as<unsigned long long>(), 0, 0, nullptr));
    bool clustered = false;
    GD_TRY(launch_minmax_mask_cluster(keys.as<unsigned long long>(), 0, (int)n, 1, pose.as<PoseDev>(), kf, mm.as<unsigned>(),
                                      dmask.as<uint8_t>(), 0, 0, nullptr, &clustered));
    if (!clustered) {
        GD_TRY(launch_minmax_reset(mm.as<unsigned>(), 1, 0));
        GD_TRY(launch_minmax(keys.as<unsigned long long>(), 0, (int)n, 1, pose.as<PoseDev>(), kf, mm.as<unsigned>(), 0, nullptr));
        GD_TRY(launch_normalize_mask(keys.as<unsigned long long>(), 0, (int)n, 1, mm.as<unsigned>(), pose.as<PoseDev>(), kf,
                                     dmask.as<uint8_t>(), 0, 0, nullptr));
    }
    GD_TRY(launch_resolve_dist(keys.as<unsigned long long>(), 0, (int)n, 1, pose.as<PoseDev>(), kf, ddist.as<float>(), 0, 0));
    GD_CUDA(cudaDeviceSynchronize());
    if (dist) GD_CUDA(cudaMemcpy(dist, ddist.p, n * 4, cudaMemcpyDeviceToHost));
    if (mask) GD_CUDA(cudaMemcpy(mask, dmask.p, n, cudaMemcpyDeviceToHost));
    if (minmax) {
        unsigned bits[GD_MM_WORDS];
        GD_CUDA(cudaMemcpy(bits, mm.p, sizeof(bits), cudaMemcpyDeviceToHost));
        bits[1] = ~bits[1];
        if (bits[2] == 0u) bits[0] = bits[1] = 0x7FC00000u;
        std::memcpy(minmax, bits, 8);
    }
    return GD_OK;
}

int gd_stage_polyexp(int device, const uint8_t* gray, int w, int h, int k, float* out, int* lw, int* lh)
{
    GD_REQUIRE(gray && out && lw && lh, "null argument");
    GD_TRY(select_device(device));
    FbPlan plan;
    GD_TRY(fb_make_plan(w, h, 0.5, 3, 3, 5, 1.2, 15, &plan));
    GD_REQUIRE(k >= 0 && k < plan.nlevels, "level out of range");
    const size_t n = (size_t)w * h;
    DevBuf g, I, R;
    GD_TRY(g.alloc(n));
    GD_TRY(I.alloc(plan.i_floats * 4));
    GD_TRY(R.alloc(plan.r_floats * 4));
    GD_CUDA(cudaMemcpy(g.p, gray, n, cudaMemcpyHostToDevice));
    GD_TRY(fb_launch_pyramid_polyexp(plan, g.as<uint8_t>(), 0, 1, I.as<float>(), 0, R.as<float>(), 0, 0, nullptr));
    GD_CUDA(cudaDeviceSynchronize());
    const FbLevel& L = plan.lv[k];
    *lw = L.w;
    *lh = L.h;
    // device layout: float4 plane (channels 0..3) + float plane (channel 4); hand back 5 planes
    const size_t npx = (size_t)L.w * L.h, npad = align_up(npx, 64);
    std::vector<float> tmp(5 * npad);
    GD_CUDA(cudaMemcpy(tmp.data(), R.as<float>() + L.r_off, 5 * npad * sizeof(float), cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < npx; ++i) {
        for (int c = 0; c < 4; ++c) out[c * npx + i] = tmp[4 * i + c];
        out[4 * npx + i] = tmp[4 * npad + i];
    }
    return GD_OK;
}

int gd_stage_farneback(int device, const uint8_t* prev, const uint8_t* next, int w, int h, float* flow)
{
    GD_REQUIRE(prev && next && flow, "null argument");
    GD_TRY(select_device(device));
    FbPlan plan;
    GD_TRY(fb_make_plan(w, h, 0.5, 3, 3, 5, 1.2, 15, &plan));
    const size_t n = (size_t)w * h;
    DevBuf g0, g1, I, R0, R1, fa, fb, Mb, Mb2;
    GD_TRY(g0.alloc(n));
    GD_TRY(g1.alloc(n));
    GD_TRY(I.alloc(plan.i_floats * 4));
    GD_TRY(R0.alloc(plan.r_floats * 4));
    GD_TRY(R1.alloc(plan.r_floats * 4));
    GD_TRY(fa.alloc(plan.f_float2 * 8));
    GD_TRY(fb.alloc(plan.f_float2 * 8));
    GD_TRY(Mb.alloc(plan.m_floats * 4));
    GD_TRY(Mb2.alloc(plan.m_floats * 4));
    FbFlowBuffers fbuf;
    GD_TRY(fb_prepare_flow_buffers(plan, 1, Mb.as<float>(), Mb2.as<float>(), Mb.bytes, &fbuf));
    GD_CUDA(cudaMemcpy(g0.p, prev, n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(g1.p, next, n, cudaMemcpyHostToDevice));
    GD_TRY(fb_launch_pyramid_polyexp(plan, g0.as<uint8_t>(), 0, 1, I.as<float>(), 0, R0.as<float>(), 0, 0, nullptr));
    GD_TRY(fb_launch_pyramid_polyexp(plan, g1.as<uint8_t>(), 0, 1, I.as<float>(), 0, R1.as<float>(), 0, 0, nullptr));
    const float2* fin = nullptr;
    GD_TRY(fb_launch_flow(plan, R0.as<float>(), R1.as<float>(), 0, 1, fa.as<float2>(), fb.as<float2>(), 0, &fbuf, &fin, 0, nullptr));
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(flow, fin, n * 8, cudaMemcpyDeviceToHost));
    return GD_OK;
}

}  // extern "C"
