// getrt.cu — building blocks of GeoMaskMaker::GetRt (GD-SLAM src/GeoMaskMaker.cc:77-156) for sm_100a, SURVEY 8(f)-1.
// Built with --fmad=false (integer stages and individually rounded f32).  Only the single-kernel gd_stage_* entry points
// exist so far; each one is bit-exact against oracle/getrt_proto.py, which is pinned against cv2 4.13:
//   gd_stage_resize_linear_exact   cv::resize(INTER_LINEAR_EXACT) between cv::ORB pyramid levels
//   gd_stage_gaussian7_float       cv::GaussianBlur(7x7, sigma 2) as cv::ORB gets it (float separable path: it blurs a submatrix)
//   gd_stage_harris                HarrisResponses (blockSize 7, k 0.04)
//   gd_stage_hamming_crosscheck    BFMatcher(NORM_HAMMING, crossCheck = true)::match
#include "gd_internal.h"

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

}  // extern "C"
