// getrt.cu — building blocks of GeoMaskMaker::GetRt (GD-SLAM src/GeoMaskMaker.cc:77-156) for sm_100a, SURVEY 8(f)-1.
// Built with --fmad=false (integer stages and individually rounded f32).  Only the single-kernel gd_stage_* entry points
// exist so far; each one is bit-exact against oracle/getrt_proto.py, which is pinned against cv2 4.13:
//   gd_stage_resize_linear_exact   cv::resize(INTER_LINEAR_EXACT) between cv::ORB pyramid levels
//   gd_stage_gaussian7_float       cv::GaussianBlur(7x7, sigma 2) as cv::ORB gets it (float separable path: it blurs a submatrix)
//   gd_stage_harris                HarrisResponses (blockSize 7, k 0.04)
//   gd_stage_hamming_crosscheck    BFMatcher(NORM_HAMMING, crossCheck = true)::match
#include "gd_internal.h"
#include "geomask.cuh"
#include "orb.cuh"

#include <algorithm>
#include <cmath>
#include <vector>

namespace gd {

// ------------------------------------------------------------------------------------------------ INTER_LINEAR_EXACT
// 8.8 fixed-point weights round(frac * 256) per axis, horizontal then vertical, one rounding (v + 2^15) >> 16.
static void linear_exact_axis_table(int dn, int sn, ushort4* t)
{
    const double scale = (double)sn / dn;
    for (int d = 0; d < dn; ++d) {
        double f = (d + 0.5) * scale - 0.5;
        int i = (int)std::floor(f);
        double fr = f - i;
        if (i < 0) { i = 0; fr = 0; }
        if (i >= sn - 1) { i = sn - 1; fr = 0; }
        const int a1 = (int)std::floor(fr * 256 + 0.5);
        t[d] = make_ushort4((unsigned short)i, (unsigned short)std::min(i + 1, sn - 1), (unsigned short)(256 - a1), (unsigned short)a1);
    }
}

__global__ void __launch_bounds__(256) k_resize_linear_exact(const uint8_t* __restrict__ src, int sw, uint8_t* __restrict__ dst, int dw,
                                                             int dh, const ushort4* __restrict__ xt, const ushort4* __restrict__ yt)
{
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const ushort4 X = __ldg(xt + dx), Y = __ldg(yt + dy);
    const uint8_t* r0 = src + (size_t)Y.x * sw;
    const uint8_t* r1 = src + (size_t)Y.y * sw;
    const int h0 = r0[X.x] * (int)X.z + r0[X.y] * (int)X.w;
    const int h1 = r1[X.x] * (int)X.z + r1[X.y] * (int)X.w;
    const int v = (h0 * (int)Y.z + h1 * (int)Y.w + 32768) >> 16;
    dst[(size_t)dy * dw + dx] = (uint8_t)max(0, min(255, v));
}

// ------------------------------------------------------------------------------------------------ float Gaussian 7x7
// row = g3 p[x] + sum_k g[3+k] (p[x+k] + p[x-k]), column the same on the row results, cvRound at the end; every product and
// sum individually rounded to f32 in exactly this order (REFLECT_101 at the borders).
struct Gauss7 {
    float g[4];  // g[0] = centre tap
};
constexpr int GF_W = 32, GF_H = 8;

__device__ __forceinline__ int reflect101_dev(int p, int len)
{
    p = abs(p);
    return p >= len ? 2 * len - 2 - p : p;
}

__global__ void __launch_bounds__(GF_W* GF_H) k_gaussian7_float(const uint8_t* __restrict__ src, int w, int h, Gauss7 G,
                                                                uint8_t* __restrict__ dst)
{
    __shared__ float s_in[GF_H + 6][GF_W + 6];
    __shared__ float s_row[GF_H + 6][GF_W];
    const int x0 = blockIdx.x * GF_W, y0 = blockIdx.y * GF_H;
    const int tid = threadIdx.y * GF_W + threadIdx.x;
    for (int i = tid; i < (GF_H + 6) * (GF_W + 6); i += GF_W * GF_H) {
        const int ly = i / (GF_W + 6), lx = i - ly * (GF_W + 6);
        const int x = reflect101_dev(x0 + lx - 3, w), y = reflect101_dev(y0 + ly - 3, h);
        s_in[ly][lx] = (float)src[(size_t)y * w + x];
    }
    __syncthreads();
    for (int i = tid; i < (GF_H + 6) * GF_W; i += GF_W * GF_H) {
        const int ly = i / GF_W, lx = i - ly * GF_W;
        const float* p = &s_in[ly][lx + 3];
        float r = G.g[0] * p[0];
#pragma unroll
        for (int k = 1; k <= 3; ++k) r = r + G.g[k] * (p[k] + p[-k]);
        s_row[ly][lx] = r;
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    const int ly = threadIdx.y + 3, lx = threadIdx.x;
    float c = G.g[0] * s_row[ly][lx];
#pragma unroll
    for (int k = 1; k <= 3; ++k) c = c + G.g[k] * (s_row[ly + k][lx] + s_row[ly - k][lx]);
    const int v = __float2int_rn(c);  // cvRound: round half to even
    dst[(size_t)y * w + x] = (uint8_t)max(0, min(255, v));
}

// ------------------------------------------------------------------------------------------------ Harris responses
__global__ void __launch_bounds__(128) k_harris(const uint8_t* __restrict__ img, int w, const int* __restrict__ xs,
                                                const int* __restrict__ ys, int n, float s4, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t* c = img + (size_t)ys[i] * w + xs[i];
    int a = 0, b = 0, cc = 0;
    for (int dy = -3; dy <= 3; ++dy)
        for (int dx = -3; dx <= 3; ++dx) {
            const uint8_t* p = c + dy * w + dx;
            const int Ix = (p[1] - p[-1]) * 2 + (p[-w + 1] - p[-w - 1]) + (p[w + 1] - p[w - 1]);
            const int Iy = (p[w] - p[-w]) * 2 + (p[w - 1] - p[-w - 1]) + (p[w + 1] - p[-w + 1]);
            a += Ix * Ix;
            b += Iy * Iy;
            cc += Ix * Iy;
        }
    const float fa = (float)a, fb = (float)b, fc = (float)cc;
    float t = fa * fb - fc * fc;  // two products and one difference, each rounded (no FMA in this file)
    const float sab = fa + fb;
    t = t - 0.04f * sab * sab;
    out[i] = t * s4;
}

// ------------------------------------------------------------------------------------------------ Hamming nearest neighbour
// one thread per query descriptor (eight 32-bit words in registers); every thread walks the train set in the same order,
// so the train words are broadcast loads.  Strict '<' keeps the smallest index among equal distances, like cv::batchDistance.
__global__ void __launch_bounds__(128) k_hamming_nn(const uint4* __restrict__ q, int nq, const uint4* __restrict__ t, int nt,
                                                    int* __restrict__ nn, int* __restrict__ dist)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nq) return;
    const uint4 a0 = q[2 * i], a1 = q[2 * i + 1];
    int best = 1 << 30, bi = -1;
    for (int j = 0; j < nt; ++j) {
        const uint4 b0 = __ldg(t + 2 * j), b1 = __ldg(t + 2 * j + 1);
        const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                      __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
        if (d < best) { best = d; bi = j; }
    }
    nn[i] = bi;
    dist[i] = best;
}


// cv::KeyPointsFilter::retainBest as OpenCV writes it: std::nth_element + std::partition on the responses.  The order it
// leaves behind IS the order of cv::ORB's output, so the same libstdc++ routines run here on the same sequence.
static void retain_best_order(const std::vector<float>& resp, int n_points, std::vector<int>* perm)
{
    struct E {
        float r;
        int i;
    };
    std::vector<E> v(resp.size());
    for (size_t i = 0; i < resp.size(); ++i) v[i] = {resp[i], (int)i};
    if (n_points >= 0 && (int)v.size() > n_points) {
        if (n_points == 0) {
            perm->clear();
            return;
        }
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const E& a, const E& b) { return a.r > b.r; });
        const float amb = v[n_points - 1].r;
        auto e = std::partition(v.begin() + n_points, v.end(), [amb](const E& a) { return a.r >= amb; });
        v.resize((size_t)(e - v.begin()));
    }
    perm->resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) (*perm)[i] = v[i].i;
}

static Gauss7 gauss7_taps()
{
    Gauss7 G;  // getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, then rounded to float
    double g[7], sum = 0;
    for (int i = 0; i < 7; ++i) {
        const double x = i - 3;
        g[i] = std::exp(-(x * x) / (2.0 * 2.0 * 2.0));
        sum += g[i];
    }
    for (int k = 0; k <= 3; ++k) G.g[k] = (float)(g[3 + k] / sum);
    return G;
}

static float harris_scale4()
{
    const float scale = 1.0f / (float)((1 << 2) * 7 * 255.0f);
    volatile float s4 = scale * scale;  // (s * s * s * s) as three individually rounded products
    s4 = s4 * scale;
    s4 = s4 * scale;
    return s4;
}

// cv::ORB::create(nfeatures, 1.2f, 8, 31, 0, 2, HARRIS_SCORE, 31, 20)->detectAndCompute on one 8-bit image: pyramid, FAST,
// Harris, blur, orientation and descriptors on the device; the two retainBest orderings on the host (tiny arrays).
static int cvorb_detect_and_compute(const uint8_t* gray, int w, int h, int nfeatures, std::vector<gd_keypoint>* kps,
                                    std::vector<uint8_t>* desc)
{
    constexpr int NL = 8, EDGE = 31;
    const float sf = 1.2f;
    int nper[NL];
    {
        const double factor = 1.0 / (double)sf;
        double nd = nfeatures * (1 - factor) / (1 - std::pow(factor, (double)NL));
        int sum = 0;
        for (int l = 0; l < NL - 1; ++l) {
            nper[l] = (int)std::nearbyint(nd);
            sum += nper[l];
            nd *= factor;
        }
        nper[NL - 1] = std::max(nfeatures - sum, 0);
    }
    const Gauss7 G = gauss7_taps();
    const float s4 = harris_scale4();
    DevBuf lvl[NL], blur, score, kept, dx, dy, dresp, dang, ddesc, tab;
    int lw[NL], lh[NL];
    float scale[NL];
    lw[0] = w; lh[0] = h; scale[0] = 1.0f;
    GD_TRY(lvl[0].alloc((size_t)w * h));
    GD_TRY(blur.alloc((size_t)w * h));
    GD_TRY(score.alloc((size_t)w * h));
    GD_TRY(kept.alloc((size_t)w * h));
    GD_CUDA(cudaMemcpy(lvl[0].p, gray, (size_t)w * h, cudaMemcpyHostToDevice));
    std::vector<uint8_t> hk((size_t)w * h);
    kps->clear();
    desc->clear();
    for (int l = 0; l < NL; ++l) {
        if (l > 0) {
            scale[l] = (float)std::pow((double)sf, (double)l);
            lw[l] = (int)std::nearbyint((float)w / scale[l]);
            lh[l] = (int)std::nearbyint((float)h / scale[l]);
            if (lw[l] < 2 || lh[l] < 2) break;
            GD_TRY(lvl[l].alloc((size_t)lw[l] * lh[l]));
            std::vector<ushort4> t((size_t)lw[l] + lh[l]);
            linear_exact_axis_table(lw[l], lw[l - 1], t.data());
            linear_exact_axis_table(lh[l], lh[l - 1], t.data() + lw[l]);
            GD_TRY(tab.alloc(t.size() * sizeof(ushort4)));
            GD_CUDA(cudaMemcpy(tab.p, t.data(), t.size() * sizeof(ushort4), cudaMemcpyHostToDevice));
            k_resize_linear_exact<<<dim3(cdiv(lw[l], 32), cdiv(lh[l], 8)), dim3(32, 8)>>>(lvl[l - 1].as<uint8_t>(), lw[l - 1], lvl[l].as<uint8_t>(),
                                                                                    lw[l], lh[l], tab.as<ushort4>(), tab.as<ushort4>() + lw[l]);
            GD_CUDA(cudaGetLastError());
        }
        const int W = lw[l], H = lh[l];
        if (std::min(W, H) <= 2 * EDGE) continue;
        const uint8_t* img = lvl[l].as<uint8_t>();
        GD_TRY(orb_fast_whole(img, W, H, W, 20, score.as<uint8_t>(), kept.as<uint8_t>(), 0));
        GD_CUDA(cudaMemcpy(hk.data(), kept.p, (size_t)W * H, cudaMemcpyDeviceToHost));
        // raster-ordered FAST output inside the 31-px border (KeyPointsFilter::runByImageBorder), response = S' - 1
        std::vector<int> xs, ys;
        std::vector<float> resp;
        for (int y = EDGE; y < H - EDGE; ++y)
            for (int x = EDGE; x < W - EDGE; ++x)
                if (hk[(size_t)y * W + x]) {
                    xs.push_back(x);
                    ys.push_back(y);
                    resp.push_back((float)(hk[(size_t)y * W + x] - 1));
                }
        std::vector<int> perm;
        retain_best_order(resp, 2 * nper[l], &perm);
        std::vector<int> x2(perm.size()), y2(perm.size());
        for (size_t i = 0; i < perm.size(); ++i) { x2[i] = xs[perm[i]]; y2[i] = ys[perm[i]]; }
        const int n2 = (int)perm.size();
        if (n2 == 0) continue;
        GD_TRY(dx.alloc(sizeof(int) * n2));
        GD_TRY(dy.alloc(sizeof(int) * n2));
        GD_TRY(dresp.alloc(sizeof(float) * n2));
        GD_CUDA(cudaMemcpy(dx.p, x2.data(), sizeof(int) * n2, cudaMemcpyHostToDevice));
        GD_CUDA(cudaMemcpy(dy.p, y2.data(), sizeof(int) * n2, cudaMemcpyHostToDevice));
        k_harris<<<cdiv(n2, 128), 128>>>(img, W, dx.as<int>(), dy.as<int>(), n2, s4, dresp.as<float>());
        GD_CUDA(cudaGetLastError());
        std::vector<float> hr((size_t)n2);
        GD_CUDA(cudaMemcpy(hr.data(), dresp.p, sizeof(float) * n2, cudaMemcpyDeviceToHost));
        retain_best_order(hr, nper[l], &perm);
        const int n3 = (int)perm.size();
        if (n3 == 0) continue;
        std::vector<int> x3(n3), y3(n3);
        for (int i = 0; i < n3; ++i) { x3[i] = x2[perm[i]]; y3[i] = y2[perm[i]]; }
        GD_CUDA(cudaMemcpy(dx.p, x3.data(), sizeof(int) * n3, cudaMemcpyHostToDevice));
        GD_CUDA(cudaMemcpy(dy.p, y3.data(), sizeof(int) * n3, cudaMemcpyHostToDevice));
        k_gaussian7_float<<<dim3(cdiv(W, GF_W), cdiv(H, GF_H)), dim3(GF_W, GF_H)>>>(img, W, H, G, blur.as<uint8_t>());
        GD_CUDA(cudaGetLastError());
        GD_TRY(dang.alloc(sizeof(float) * n3));
        GD_TRY(ddesc.alloc((size_t)32 * n3));
        GD_TRY(orb_cv_describe(img, blur.as<uint8_t>(), W, dx.as<int>(), dy.as<int>(), n3, dang.as<float>(), ddesc.as<uint8_t>(), 0));
        std::vector<float> ang((size_t)n3);
        const size_t d0 = desc->size();
        desc->resize(d0 + (size_t)32 * n3);
        GD_CUDA(cudaMemcpy(ang.data(), dang.p, sizeof(float) * n3, cudaMemcpyDeviceToHost));
        GD_CUDA(cudaMemcpy(desc->data() + d0, ddesc.p, (size_t)32 * n3, cudaMemcpyDeviceToHost));
        for (int i = 0; i < n3; ++i) {
            gd_keypoint k;
            k.x = (float)x3[i] * scale[l];
            k.y = (float)y3[i] * scale[l];
            k.size = 31.0f * scale[l];
            k.angle = ang[i];
            k.response = hr[perm[i]];
            k.octave = l;
            k.class_id = -1;
            kps->push_back(k);
        }
    }
    return GD_OK;
}

}  // namespace gd

using namespace gd;

extern "C" {

int gd_stage_resize_linear_exact(int device, const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    GD_REQUIRE(src && dst && sw > 1 && sh > 1 && dw > 0 && dh > 0 && sw < 65536 && sh < 65536, "bad argument");
    GD_TRY(select_device(device));
    std::vector<ushort4> tab((size_t)dw + dh);
    linear_exact_axis_table(dw, sw, tab.data());
    linear_exact_axis_table(dh, sh, tab.data() + dw);
    DevBuf s, d, t;
    GD_TRY(s.alloc((size_t)sw * sh));
    GD_TRY(d.alloc((size_t)dw * dh));
    GD_TRY(t.alloc(tab.size() * sizeof(ushort4)));
    GD_CUDA(cudaMemcpy(s.p, src, (size_t)sw * sh, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(t.p, tab.data(), tab.size() * sizeof(ushort4), cudaMemcpyHostToDevice));
    k_resize_linear_exact<<<dim3(cdiv(dw, 32), cdiv(dh, 8)), dim3(32, 8)>>>(s.as<uint8_t>(), sw, d.as<uint8_t>(), dw, dh, t.as<ushort4>(),
                                                                          t.as<ushort4>() + dw);
    GD_CUDA(cudaGetLastError());
    GD_CUDA(cudaMemcpy(dst, d.p, (size_t)dw * dh, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_gaussian7_float(int device, const uint8_t* src, int w, int h, uint8_t* dst)
{
    GD_REQUIRE(src && dst && w > 3 && h > 3, "bad argument");
    GD_TRY(select_device(device));
    Gauss7 G;
    {  // getGaussianKernel(7, 2, CV_32F): exp(-x^2 / (2 sigma^2)) normalised in double, then rounded to float
        double g[7], sum = 0;
        for (int i = 0; i < 7; ++i) {
            const double x = i - 3;
            g[i] = std::exp(-(x * x) / (2.0 * 2.0 * 2.0));
            sum += g[i];
        }
        for (int k = 0; k <= 3; ++k) G.g[k] = (float)(g[3 + k] / sum);
    }
    DevBuf s, d;
    GD_TRY(s.alloc((size_t)w * h));
    GD_TRY(d.alloc((size_t)w * h));
    GD_CUDA(cudaMemcpy(s.p, src, (size_t)w * h, cudaMemcpyHostToDevice));
    k_gaussian7_float<<<dim3(cdiv(w, GF_W), cdiv(h, GF_H)), dim3(GF_W, GF_H)>>>(s.as<uint8_t>(), w, h, G, d.as<uint8_t>());
    GD_CUDA(cudaGetLastError());
    GD_CUDA(cudaMemcpy(dst, d.p, (size_t)w * h, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_harris(int device, const uint8_t* img, int w, int h, const int* xs, const int* ys, int n, float* out)
{
    GD_REQUIRE(img && xs && ys && out && n >= 0, "bad argument");
    GD_TRY(select_device(device));
    if (n == 0) return GD_OK;
    for (int i = 0; i < n; ++i) GD_REQUIRE(xs[i] >= 4 && ys[i] >= 4 && xs[i] < w - 4 && ys[i] < h - 4, "keypoint too close to the border");
    float scale = 1.0f / (float)((1 << 2) * 7 * 255.0f);
    volatile float s4 = scale * scale;  // (s * s * s * s) as three individually rounded products
    s4 = s4 * scale;
    s4 = s4 * scale;
    DevBuf im, dx, dy, o;
    GD_TRY(im.alloc((size_t)w * h));
    GD_TRY(dx.alloc(sizeof(int) * n));
    GD_TRY(dy.alloc(sizeof(int) * n));
    GD_TRY(o.alloc(sizeof(float) * n));
    GD_CUDA(cudaMemcpy(im.p, img, (size_t)w * h, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dx.p, xs, sizeof(int) * n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dy.p, ys, sizeof(int) * n, cudaMemcpyHostToDevice));
    k_harris<<<cdiv(n, 128), 128>>>(im.as<uint8_t>(), w, dx.as<int>(), dy.as<int>(), n, s4, o.as<float>());
    GD_CUDA(cudaGetLastError());
    GD_CUDA(cudaMemcpy(out, o.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_hamming_crosscheck(int device, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* query_idx, int* train_idx,
                                int* distance, int capacity, int* n_matches)
{
    GD_REQUIRE(d1 && d2 && query_idx && train_idx && distance && n_matches && n1 > 0 && n2 > 0, "bad argument");
    GD_TRY(select_device(device));
    DevBuf a, b, nn12, nn21, dd12, dd21;
    GD_TRY(a.alloc((size_t)n1 * 32));
    GD_TRY(b.alloc((size_t)n2 * 32));
    GD_TRY(nn12.alloc(sizeof(int) * n1));
    GD_TRY(dd12.alloc(sizeof(int) * n1));
    GD_TRY(nn21.alloc(sizeof(int) * n2));
    GD_TRY(dd21.alloc(sizeof(int) * n2));
    GD_CUDA(cudaMemcpy(a.p, d1, (size_t)n1 * 32, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(b.p, d2, (size_t)n2 * 32, cudaMemcpyHostToDevice));
    k_hamming_nn<<<cdiv(n1, 128), 128>>>(a.as<uint4>(), n1, b.as<uint4>(), n2, nn12.as<int>(), dd12.as<int>());
    k_hamming_nn<<<cdiv(n2, 128), 128>>>(b.as<uint4>(), n2, a.as<uint4>(), n1, nn21.as<int>(), dd21.as<int>());
    GD_CUDA(cudaGetLastError());
    std::vector<int> h12(n1), hd(n1), h21(n2);
    GD_CUDA(cudaMemcpy(h12.data(), nn12.p, sizeof(int) * n1, cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(hd.data(), dd12.p, sizeof(int) * n1, cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(h21.data(), nn21.p, sizeof(int) * n2, cudaMemcpyDeviceToHost));
    int m = 0;  // cross check, ordered by query index like BFMatcher returns them
    for (int q = 0; q < n1; ++q) {
        const int t = h12[q];
        if (t >= 0 && h21[t] == q) {
            if (m < capacity) {
                query_idx[m] = q;
                train_idx[m] = t;
                distance[m] = hd[q];
            }
            ++m;
        }
    }
    *n_matches = m;
    GD_REQUIRE(m <= capacity, "match capacity too small");
    return GD_OK;
}


int gd_stage_cvorb_detect_and_compute(int device, const uint8_t* gray, int w, int h, int nfeatures, gd_keypoint* kps, uint8_t* desc,
                                      int capacity, int* n)
{
    GD_REQUIRE(gray && kps && desc && n && w > 0 && h > 0 && nfeatures > 0, "bad argument");
    GD_TRY(select_device(device));
    std::vector<gd_keypoint> k;
    std::vector<uint8_t> d;
    GD_TRY(cvorb_detect_and_compute(gray, w, h, nfeatures, &k, &d));
    *n = (int)k.size();
    GD_REQUIRE(*n <= capacity, "keypoint capacity too small");
    std::memcpy(kps, k.data(), sizeof(gd_keypoint) * k.size());
    std::memcpy(desc, d.data(), d.size());
    return GD_OK;
}


int gd_getrt_points(int device, const uint8_t* gray_first, const uint8_t* gray_second, int w, int h, const float* depth_first_m,
                    const float K[9], const float* dist, int ndist, float* object_points, float* image_pixels, int* n_points)
{
    GD_REQUIRE(gray_first && gray_second && depth_first_m && K && object_points && image_pixels && n_points, "null argument");
    for (int i = 0; i < ndist; ++i) GD_REQUIRE(!dist || dist[i] == 0.f, "distorted cameras are not built for GetRt yet (undistortPoints on the matches)");
    GD_TRY(select_device(device));
    *n_points = 0;
    std::vector<gd_keypoint> k1, k2;
    std::vector<uint8_t> d1, d2;
    GD_TRY(cvorb_detect_and_compute(gray_first, w, h, 2000, &k1, &d1));    // GeoMaskMaker.cc:82-90
    GD_TRY(cvorb_detect_and_compute(gray_second, w, h, 2000, &k2, &d2));
    if (k1.empty() || k2.empty()) return GD_OK;
    const int n1 = (int)k1.size(), n2 = (int)k2.size();
    std::vector<int> mq(n1), mt(n1), md(n1);
    int nm = 0;
    GD_TRY(gd_stage_hamming_crosscheck(device, d1.data(), n1, d2.data(), n2, mq.data(), mt.data(), md.data(), n1, &nm));  // :92-94
    // sort(matches.begin(), matches.end()) (:95): DMatch::operator< looks at the distance only; the same libstdc++ introsort
    // on the same sequence leaves equal distances in the same order as the reference
    struct M {
        float d;
        int i;
        bool operator<(const M& o) const { return d < o.d; }
    };
    std::vector<M> ms((size_t)nm);
    for (int i = 0; i < nm; ++i) ms[i] = {(float)md[i], i};
    std::sort(ms.begin(), ms.end());
    const int ntop = std::min(nm, 100);  // the reference takes begin()+100 unconditionally (:97); fewer matches are undefined there
    CamConst cam;
    make_cam_const(K, &cam);
    int n = 0;
    for (int r = 0; r < ntop; ++r) {
        const int i = ms[r].i;
        const float x = k1[mq[i]].x, y = k1[mq[i]].y;      // undistortPoints with D = 0, P = K is the identity in f32
        const int dx = (int)x, dy = (int)y;                 // :122-123
        if (dx < 0 || dy < 0 || dx >= w || dy >= h) continue;
        const float depth = depth_first_m[(size_t)dy * w + dx];
        if (depth == 0.f) continue;                         // :125-128
        for (int c = 0; c < 3; ++c) {                       // (inv(K) * [x y 1]^T) * depth, f32 gemm of inner length 3 (:133)
            volatile float t = cam.Ki[3 * c] * x;
            t = t + cam.Ki[3 * c + 1] * y;
            t = t + cam.Ki[3 * c + 2] * 1.0f;
            object_points[3 * n + c] = t * depth;
        }
        image_pixels[2 * n] = k2[mt[i]].x;                  // :139
        image_pixels[2 * n + 1] = k2[mt[i]].y;
        ++n;
    }
    *n_points = n;
    return GD_OK;
}

}  // extern "C"
