// api_orb.cu — C ABI: gd_orb_* (ORBextractor drop-in) and the gd_stage_* parity harness for its kernels.
#include "orb.cuh"

struct gd_orb {
    gd::OrbCore core;
    gd::LaunchStats stats;
};

using namespace gd;

namespace gd {
// copies results of the last extract_resident() to host buffers; n_out[b] always set
int orb_fetch_results(OrbCore& c, gd_keypoint* const* kps, uint8_t* const* desc, int capacity, int* n_out)
{
    int* hn = c.h_n.as<int>();
    GD_CUDA(cudaMemcpyAsync(hn, c.out_n.p, sizeof(int) * c.batch, cudaMemcpyDeviceToHost, c.stream));
    GD_CUDA(cudaMemcpyAsync(hn + c.batch, c.err.p, sizeof(int), cudaMemcpyDeviceToHost, c.stream));
    GD_CUDA(cudaStreamSynchronize(c.stream));
    int rc = GD_OK;
    if (hn[c.batch] != 0) {
        // reported once, then cleared: the flag must not poison later extractions.  The clamped results are still delivered.
        set_error("ORB kernel reported an internal capacity overflow (flags %d): results are truncated", hn[c.batch]);
        GD_CUDA(cudaMemsetAsync(c.err.p, 0, sizeof(int), c.stream));
        rc = GD_EINTERNAL;
    }
    for (int b = 0; b < c.batch; ++b) {
        const int n = hn[b];
        if (n_out) n_out[b] = n;
        int m = n;
        if (n > capacity) {
            m = capacity;
            if (rc == GD_OK) {
                rc = GD_ECAPACITY;
                set_error("keypoint capacity %d too small for %d keypoints", capacity, n);
            }
        }
        if (kps && kps[b] && m > 0)
            GD_CUDA(cudaMemcpyAsync(kps[b], c.out_kp.as<gd_keypoint>() + (size_t)b * c.plan.kp_capacity, sizeof(gd_keypoint) * m,
                                    cudaMemcpyDeviceToHost, c.stream));
        if (desc && desc[b] && m > 0)
            GD_CUDA(cudaMemcpyAsync(desc[b], c.out_desc.as<uint8_t>() + (size_t)b * c.plan.kp_capacity * 32, (size_t)32 * m,
                                    cudaMemcpyDeviceToHost, c.stream));
    }
    GD_CUDA(cudaStreamSynchronize(c.stream));
    return rc;
}
}  // namespace gd

extern "C" {

int gd_orb_create(gd_orb_t** out, int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast,
                  int max_width, int max_height, int device, int batch)
{
    GD_REQUIRE(out, "null argument");
    *out = nullptr;
    gd_orb* h = new (std::nothrow) gd_orb();
    if (!h) return GD_ENOMEM;
    int r = h->core.init(nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_width, max_height, device, batch,
                         nullptr, &h->stats);
    if (r != GD_OK) {
        delete h;
        return r;
    }
    h->core.graphs.enabled = GraphCache::env_default(batch);
    *out = h;
    return GD_OK;
}

void gd_orb_destroy(gd_orb_t* h)
{
    if (!h) return;
    cudaSetDevice(h->core.device);
    if (h->core.stream) cudaStreamSynchronize(h->core.stream);
    delete h;
}

int gd_orb_extract(gd_orb_t* h, const uint8_t* const* gray, size_t gray_step, int w, int h_, gd_keypoint* const* kps,
                   uint8_t* const* desc, int capacity, int* n_out)
{
    GD_REQUIRE(h && gray && n_out, "null argument");
    OrbCore& c = h->core;
    GD_TRY(select_device(c.device));
    GD_REQUIRE(gray_step >= (size_t)w, "gray_step smaller than a row");
    GD_TRY(c.set_size(w, h_));
    for (int b = 0; b < c.batch; ++b) {
        GD_REQUIRE(gray[b], "null image pointer");
        GD_CUDA(cudaMemcpy2DAsync(c.level0(b), c.plan.lv[0].pitch, gray[b], gray_step, (size_t)w, h_, cudaMemcpyHostToDevice,
                                  c.stream));
    }
    GD_TRY(c.extract_resident());
    return orb_fetch_results(c, kps, desc, capacity, n_out);
}

int gd_orb_fetch_level(gd_orb_t* h, int stream, int level, uint8_t* dst, size_t dst_step, int* w, int* h_)
{
    GD_REQUIRE(h && dst, "null argument");
    OrbCore& c = h->core;
    GD_TRY(select_device(c.device));
    GD_REQUIRE(stream >= 0 && stream < c.batch && level >= 0 && level < c.plan.nlevels, "index out of range");
    const OrbLevel& L = c.plan.lv[level];
    GD_REQUIRE(dst_step >= (size_t)L.w, "dst_step smaller than a row");
    GD_CUDA(cudaStreamSynchronize(c.stream));
    GD_CUDA(cudaMemcpy2D(dst, dst_step, c.pyr.as<uint8_t>() + (size_t)stream * c.plan.pyr_bytes + L.off, L.pitch, (size_t)L.w, L.h,
                         cudaMemcpyDeviceToHost));
    if (w) *w = L.w;
    if (h_) *h_ = L.h;
    return GD_OK;
}

int gd_orb_level_size(const gd_orb_t* h, int level, int* w, int* h_)
{
    GD_REQUIRE(h && level >= 0 && level < h->core.plan.nlevels, "bad argument");
    if (w) *w = h->core.plan.lv[level].w;
    if (h_) *h_ = h->core.plan.lv[level].h;
    return GD_OK;
}

int gd_orb_features_per_level(const gd_orb_t* h, int* n_per_level)
{
    GD_REQUIRE(h && n_per_level, "null argument");
    for (int l = 0; l < h->core.plan.nlevels; ++l) n_per_level[l] = h->core.plan.lv[l].N;
    return GD_OK;
}

// ----------------------------------------------------------------------------------------------- stages
// pyramid: `out` receives the levels tightly packed one after another; level_sizes = nlevels x (w, h)
int gd_stage_orb_pyramid(int device, const uint8_t* gray, int w, int h, int nlevels, float scale, uint8_t* out, int* level_sizes)
{
    GD_REQUIRE(gray && out, "null argument");
    OrbCore c;
    GD_TRY(c.init(1000, scale, nlevels, 20, 7, w, h, device, 1, nullptr, nullptr));
    GD_CUDA(cudaMemcpy2DAsync(c.level0(0), c.plan.lv[0].pitch, gray, (size_t)w, (size_t)w, h, cudaMemcpyHostToDevice, c.stream));
    GD_TRY(c.extract_resident());
    GD_CUDA(cudaStreamSynchronize(c.stream));
    size_t off = 0;
    for (int l = 0; l < nlevels; ++l) {
        const OrbLevel& L = c.plan.lv[l];
        GD_CUDA(cudaMemcpy2D(out + off, (size_t)L.w, c.pyr.as<uint8_t>() + L.off, L.pitch, (size_t)L.w, L.h, cudaMemcpyDeviceToHost));
        off += (size_t)L.w * L.h;
        if (level_sizes) {
            level_sizes[2 * l] = L.w;
            level_sizes[2 * l + 1] = L.h;
        }
    }
    return GD_OK;
}

// candidate list of the cell loop on `gray` taken as ONE level: out = (x, y, response) float triples, border-relative
int gd_stage_fast_cells(int device, const uint8_t* gray, int w, int h, int ini_th, int min_th, float* out, int capacity, int* n)
{
    GD_REQUIRE(gray && out && n, "null argument");
    OrbCore c;
    GD_TRY(c.init(1000, 1.2f, 1, ini_th, min_th, w, h, device, 1, nullptr, nullptr));
    GD_CUDA(cudaMemcpy2DAsync(c.level0(0), c.plan.lv[0].pitch, gray, (size_t)w, (size_t)w, h, cudaMemcpyHostToDevice, c.stream));
    GD_TRY(c.extract_resident());
    GD_CUDA(cudaStreamSynchronize(c.stream));
    int cnt = 0;
    GD_CUDA(cudaMemcpy(&cnt, c.cand_cnt.p, sizeof(int), cudaMemcpyDeviceToHost));
    *n = cnt;
    std::vector<ushort4> tmp((size_t)cnt);
    if (cnt) GD_CUDA(cudaMemcpy(tmp.data(), c.cand.p, sizeof(ushort4) * cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt && i < capacity; ++i) {
        out[3 * i] = (float)tmp[i].x;
        out[3 * i + 1] = (float)tmp[i].y;
        out[3 * i + 2] = (float)tmp[i].z;
    }
    return cnt > capacity ? GD_ECAPACITY : GD_OK;
}

int gd_stage_gaussian7(int device, const uint8_t* gray, int w, int h, uint8_t* out)
{
    GD_REQUIRE(gray && out, "null argument");
    OrbCore c;
    GD_TRY(c.init(1000, 1.2f, 1, 20, 7, w, h, device, 1, nullptr, nullptr));
    GD_CUDA(cudaMemcpy2DAsync(c.level0(0), c.plan.lv[0].pitch, gray, (size_t)w, (size_t)w, h, cudaMemcpyHostToDevice, c.stream));
    GD_TRY(c.extract_resident());
    GD_CUDA(cudaStreamSynchronize(c.stream));
    GD_CUDA(cudaMemcpy2D(out, (size_t)w, c.blur.as<uint8_t>(), c.plan.lv[0].pitch, (size_t)w, h, cudaMemcpyDeviceToHost));
    return GD_OK;
}


int gd_stage_fast_whole(int device, const uint8_t* gray, int w, int h, int threshold, uint8_t* kept)
{
    GD_REQUIRE(gray && kept && w >= 7 && h >= 7, "bad argument");
    GD_TRY(select_device(device));
    DevBuf im, sc, kp;
    const size_t n = (size_t)w * h;
    GD_TRY(im.alloc(n));
    GD_TRY(sc.alloc(n));
    GD_TRY(kp.alloc(n));
    GD_CUDA(cudaMemcpy(im.p, gray, n, cudaMemcpyHostToDevice));
    GD_TRY(orb_fast_whole(im.as<uint8_t>(), w, h, w, threshold, sc.as<uint8_t>(), kp.as<uint8_t>(), 0));
    GD_CUDA(cudaDeviceSynchronize());
    GD_CUDA(cudaMemcpy(kept, kp.p, n, cudaMemcpyDeviceToHost));
    return GD_OK;
}

}  // extern "C"
