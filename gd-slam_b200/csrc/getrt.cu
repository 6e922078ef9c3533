// getrt.cu — GeoMaskMaker::GetRt (GD-SLAM src/GeoMaskMaker.cc:77-156) up to solvePnPRansac, resident and batched.
// Built with --fmad=false (integer stages and individually rounded f32 / f64, like the CPU code it reproduces).
//
//   per frame (enqueue_features)   cv::ORB::create(2000, 1.2f, 8, 31, 0, 2)->detectAndCompute(gray)            :82-90
//     k_resize_linear_exact          pyramid: every level from the previous one, INTER_LINEAR_EXACT (8.8 fixed point)
//     k_cv_fast_kept_levels          cv::FAST(20, nonmax) on every level + per-row corner counts          (orb.cu)
//     k_getrt_select                 raster-ordered corner list, retainBest(2 N_l) on the FAST response, Harris responses,
//                                    retainBest(N_l): std::nth_element / std::partition run as libstdc++ runs them, by one
//                                    thread per (level, stream) on lists in shared memory (stdalgo.cuh)
//     k_gaussian7_float_levels       the FLOAT separable 7x7 Gaussian cv::ORB's pyramid sub-matrices get
//     k_cv_describe_sel              IC_Angle + rBRIEF, cv::KeyPoint records in cv::ORB's order         (orb.cu)
//   per pair (enqueue_match)       features of ring slot t-5 against slot t
//     k_getrt_nn                     Hamming nearest neighbour in both directions (BFMatcher, crossCheck = true)  :92-94
//     k_getrt_points                 cross check in query order, std::sort -> first 100 (stdalgo::sort_prefix), undistortPoints,
//                                    depth look-up, back-projection                                               :95-141
// The caller gets object points / image pixels and hands them to cv::solvePnPRansac + cv::Rodrigues (:143-150).
#include "getrt.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <vector>

#include "stdalgo.cuh"

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

// blockIdx.z = stream; src / dst are dense (pitch = width) images at a per-stream stride
__global__ void __launch_bounds__(256) k_resize_linear_exact(const uint8_t* __restrict__ src, size_t sstride_b, int sw,
                                                             uint8_t* __restrict__ dst, size_t dstride_b, int dw, int dh,
                                                             const ushort4* __restrict__ xt, const ushort4* __restrict__ yt)
{
    const int dx = blockIdx.x * 32 + threadIdx.x, dy = blockIdx.y * 8 + threadIdx.y;
    if (dx >= dw || dy >= dh) return;
    const ushort4 X = __ldg(xt + dx), Y = __ldg(yt + dy);
    const uint8_t* s = src + (size_t)blockIdx.z * sstride_b;
    const uint8_t* r0 = s + (size_t)Y.x * sw;
    const uint8_t* r1 = s + (size_t)Y.y * sw;
    const int h0 = r0[X.x] * (int)X.z + r0[X.y] * (int)X.w;
    const int h1 = r1[X.x] * (int)X.z + r1[X.y] * (int)X.w;
    const int v = (h0 * (int)Y.z + h1 * (int)Y.w + 32768) >> 16;
    dst[(size_t)blockIdx.z * dstride_b + (size_t)dy * dw + dx] = (uint8_t)max(0, min(255, v));
}

// ------------------------------------------------------------------------------------------------ float Gaussian 7x7
// row = g3 p[x] + sum_k g[3+k] (p[x+k] + p[x-k]), column the same on the row results, cvRound at the end; every product and
// sum individually rounded to f32 in exactly this order (REFLECT_101 at the borders).
struct Gauss7 {
    float g[4];  // g[0] = centre tap
};

__device__ __forceinline__ int reflect101_dev(int p, int len)
{
    p = abs(p);
    return p >= len ? 2 * len - 2 - p : p;
}

// blockIdx = (tile, level, stream); levels dense at lv[l].off of a stream's pyramid.  One warp filters a strip of GF_COLS
// columns x GF_ROWS rows marching down: lane i owns column x0 - 3 + i (three halo lanes on either side), the row pass takes
// its six neighbours from the adjacent lanes, the seven row results the column pass needs stay in registers.  Same products
// and sums in the same order as the two-pass form (this file is built without FMA contraction).
constexpr int GF_COLS = 26, GF_ROWS = 16, GF_WARPS = 8, GF_PF = 4;

__global__ void __launch_bounds__(32 * GF_WARPS) k_gaussian7_float_levels(const uint8_t* __restrict__ pyr, size_t stride_b, CvPyrArgs a,
                                                                          Gauss7 G, uint8_t* __restrict__ out)
{
    const CvLevelDev L = a.lv[blockIdx.y];
    const int w = L.w, h = L.h;
    const int tiles_x = (w + GF_COLS - 1) / GF_COLS, tiles_y = (h + GF_ROWS * GF_WARPS - 1) / (GF_ROWS * GF_WARPS);
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int ty = (int)blockIdx.x / tiles_x, tx = (int)blockIdx.x - ty * tiles_x;
    const int lane = threadIdx.x, x = tx * GF_COLS - 3 + lane, y0 = (ty * GF_WARPS + threadIdx.y) * GF_ROWS;
    if (y0 >= h) return;  // warp-uniform
    const uint8_t* src = pyr + (size_t)blockIdx.z * stride_b + L.off;
    uint8_t* dst = out + (size_t)blockIdx.z * stride_b + L.off;
    const int xs = reflect101_dev(min(x, w + 2), w);        // lanes further right feed no owned column
    const bool owner = lane >= 3 && lane < 3 + GF_COLS && x < w;
    auto load_row = [&](int r) { return (float)__ldg(src + (size_t)reflect101_dev(min(y0 - 3 + r, h + 2), h) * w + xs); };
    float pre[GF_PF];
#pragma unroll
    for (int r = 0; r < GF_PF; ++r) pre[r] = load_row(r);
    float ring[7];
#pragma unroll
    for (int r = 0; r < GF_ROWS + 6; ++r) {
        const float v = pre[r % GF_PF];
        if (r + GF_PF < GF_ROWS + 6) pre[r % GF_PF] = load_row(r + GF_PF);
        float rr = G.g[0] * v;
#pragma unroll
        for (int k = 1; k <= 3; ++k)
            rr = rr + G.g[k] * (__shfl_down_sync(0xffffffffu, v, k) + __shfl_up_sync(0xffffffffu, v, k));
        ring[r % 7] = rr;
        if (r >= 6) {
            const int y = y0 + r - 6;
            float c = G.g[0] * ring[(r - 3) % 7];
#pragma unroll
            for (int k = 1; k <= 3; ++k) c = c + G.g[k] * (ring[(r - 3 + k) % 7] + ring[(r - 3 - k) % 7]);
            const int vi = __float2int_rn(c);  // cvRound: round half to even
            if (owner && y < h) dst[(size_t)y * w + x] = (uint8_t)max(0, min(255, vi));
        }
    }
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

// ------------------------------------------------------------------------------------------------ Harris responses
// HarrisResponses(blockSize 7, k 0.04) of cv::ORB at an integer level position (at least 4 px from the border)
__device__ __forceinline__ float harris_at(const uint8_t* __restrict__ img, int w, int x, int y, float s4)
{
    const uint8_t* c = img + (size_t)y * w + x;
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
    return t * s4;
}

__global__ void __launch_bounds__(128) k_harris(const uint8_t* __restrict__ img, int w, const int* __restrict__ xs,
                                                const int* __restrict__ ys, int n, float s4, float* __restrict__ out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = harris_at(img, w, xs[i], ys[i], s4);
}

static float harris_scale4()
{
    const float scale = 1.0f / (float)((1 << 2) * 7 * 255.0f);
    volatile float s4 = scale * scale;  // (s * s * s * s) as three individually rounded products
    s4 = s4 * scale;
    s4 = s4 * scale;
    return s4;
}

// ------------------------------------------------------------------------------------------------ selection per level
struct SelLevel {
    int w, h;
    unsigned long long off;
    int row_off;
    int N;       // features wanted at this level
    int n1_cap;  // capacity of the raster corner list
};
struct SelArgs {
    int nlevels, edge, n2_cap, sel_cap, n1_cap_max;
    float s4;
    SelLevel lv[GETRT_LEVELS];
};
struct HarrisEntry {
    float r;
    unsigned idx;
};
constexpr int SEL_THREADS = 512;

// exclusive scan of data[0..n) by warp 0 (all threads call); returns the total
__device__ int sel_block_scan(int* data, int n, int* s_total)
{
    __syncthreads();
    if (threadIdx.x < 32) {
        int carry = 0;
        for (int base = 0; base < n; base += 32) {
            const int idx = base + threadIdx.x;
            const int v = idx < n ? data[idx] : 0;
            int x = v;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int y = __shfl_up_sync(0xffffffffu, x, o);
                if ((int)threadIdx.x >= o) x += y;
            }
            if (idx < n) data[idx] = carry + x - v;
            carry += __shfl_sync(0xffffffffu, x, 31);
        }
        if (threadIdx.x == 0) *s_total = carry;
    }
    __syncthreads();
    return *s_total;
}

// ---- CTA-wide Hoare partition: the rank rule of stdalgo::rank_pair_swap evaluated with prefix sums ---------------------
// Every thread owns a contiguous chunk of [lo, hi) (at most 128 elements: n <= 128 * SEL_THREADS) and keeps the left-stop /
// right-stop flags of its elements, taken from the ORIGINAL values, in registers.  tmp: 2 * W elements of scratch (the values
// of the swapped pairs travel through it in batches of W ranks).  s_cnt: 2 * SEL_THREADS ints; s_red: 8 ints.  All threads
// call; returns (to every thread) the position std::__unguarded_partition would return; for std::partition the caller
// derives its own return value from the predicate count.
constexpr int PAR_MAX_CHUNK = 128;

template <class T, class IsL, class IsR>
__device__ int par_pair_swap(T* v, int lo, int hi, IsL is_l, IsR is_r, T* tmp, int W, int* s_cnt, int* s_red)
{
    const int tid = threadIdx.x, lane = tid & 31;
    const int m = hi - lo, chunk = (m + SEL_THREADS - 1) / SEL_THREADS;
    const int c0 = min(hi, lo + tid * chunk), c1 = min(hi, c0 + chunk);
    unsigned long long fl0 = 0, fl1 = 0, fr0 = 0, fr1 = 0;  // flags of chunk element j: bit j of (j < 64 ? x0 : x1)
    int cl = 0, cr = 0;
    for (int i = c0; i < c1; ++i) {
        const T e = v[i];
        const int j = i - c0;
        const unsigned long long bit = 1ull << (j & 63);
        if (is_l(e)) {
            cl += 1;
            if (j < 64) fl0 |= bit; else fl1 |= bit;
        }
        if (is_r(e)) {
            cr += 1;
            if (j < 64) fr0 |= bit; else fr1 |= bit;
        }
    }
    s_cnt[tid] = cl;
    s_cnt[SEL_THREADS + tid] = cr;
    if (tid < 8) s_red[tid] = tid == 0 ? 0 : 0x7FFFFFFF;  // [0] K, [1] l_0, [2] l_K, [3] r_{K-1}
    __syncthreads();
    if (tid < 32) {  // exclusive prefix sums of both count arrays by warp 0; totals in s_red[4], s_red[5]
        int carryL = 0, carryR = 0;
        for (int base = 0; base < SEL_THREADS; base += 32) {
            const int a = s_cnt[base + lane], b2 = s_cnt[SEL_THREADS + base + lane];
            int xa = a, xb = b2;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int ya = __shfl_up_sync(0xffffffffu, xa, o), yb = __shfl_up_sync(0xffffffffu, xb, o);
                if (lane >= o) {
                    xa += ya;
                    xb += yb;
                }
            }
            s_cnt[base + lane] = carryL + xa - a;
            s_cnt[SEL_THREADS + base + lane] = carryR + xb - b2;
            carryL += __shfl_sync(0xffffffffu, xa, 31);
            carryR += __shfl_sync(0xffffffffu, xb, 31);
        }
        if (lane == 0) {
            s_red[4] = carryL;
            s_red[5] = carryR;
        }
    }
    __syncthreads();
    const int preL = s_cnt[tid];                                   // left stops left of my chunk
    const int sufR = s_red[5] - s_cnt[SEL_THREADS + tid] - cr;     // right stops right of my chunk
    auto flag = [](unsigned long long x0, unsigned long long x1, int j) -> bool { return ((j < 64 ? x0 : x1) >> (j & 63)) & 1ull; };
    // decide: a left stop of rank L is swapped iff more than L right stops lie to its right (and symmetrically)
    int myK = 0;
    {
        int L = preL, Rr = sufR + cr;
        for (int i = c0; i < c1; ++i) {
            const bool isl = flag(fl0, fl1, i - c0), isr = flag(fr0, fr1, i - c0);
            if (isr) Rr -= 1;
            if (isl) {
                if (L == 0) atomicMin(&s_red[1], i);
                if (Rr > L)
                    myK += 1;
                else
                    atomicMin(&s_red[2], i);
            }
            if (isr && L > Rr) atomicMin(&s_red[3], i);
            if (isl) L += 1;
        }
    }
    if (myK) atomicAdd(&s_red[0], myK);
    __syncthreads();
    const int K = s_red[0];
    for (int base = 0; base < K; base += W) {  // values of the pairs with rank in [base, base + W) through the scratch
        for (int pass = 0; pass < 2; ++pass) {
            int L = preL, Rr = sufR + cr;
            for (int i = c0; i < c1; ++i) {
                const bool isl = flag(fl0, fl1, i - c0), isr = flag(fr0, fr1, i - c0);
                if (isr) Rr -= 1;
                if (isl && Rr > L && L >= base && L < base + W) {
                    if (pass == 0)
                        tmp[L - base] = v[i];
                    else
                        v[i] = tmp[W + L - base];
                } else if (isr && L > Rr && Rr >= base && Rr < base + W) {
                    if (pass == 0)
                        tmp[W + Rr - base] = v[i];
                    else
                        v[i] = tmp[Rr - base];
                }
                if (isl) L += 1;
            }
            __syncthreads();
        }
    }
    const int l0 = s_red[1], lK = s_red[2], rK1 = s_red[3];
    __syncthreads();
    if (K == 0) return l0;
    return (lK != 0x7FFFFFFF && lK < rK1) ? lK : rK1;
}

// std::nth_element by the whole CTA: libstdc++'s introselect loop with the partition above; ranges of at most SEQ_CUT
// elements (and the heap fallback) are finished by thread 0 with the sequential restatement.  Less(a, b).
constexpr int PAR_SEQ_CUT = 256;
template <class T, class Less>
__device__ void par_nth_element(T* v, int n, int nth, Less less, T* tmp, int W, int* s_cnt, int* s_red)
{
    int first = 0, last = n;
    int depth = stdalgo::lg_((long)n) * 2;
    while (last - first > 3) {
        if (last - first <= PAR_SEQ_CUT || depth == 0) break;
        --depth;
        if (threadIdx.x == 0) stdalgo::move_median_to_first(v + first, v + first + 1, v + first + (last - first) / 2, v + last - 1, less);
        __syncthreads();
        const T pivot = v[first];
        const int cut = par_pair_swap(v, first + 1, last, [pivot, less](const T& e) { return !less(e, pivot); },
                                      [pivot, less](const T& e) { return !less(pivot, e); }, tmp, W, s_cnt, s_red);
        if (cut <= nth)
            first = cut;
        else
            last = cut;
    }
    if (threadIdx.x == 0) stdalgo::introselect(v + first, v + nth, v + last, depth, less);  // also covers last - first <= 3
    __syncthreads();
}

// cv::KeyPointsFilter::retainBest by the whole CTA (see stdalgo::retain_best for the sequential statement); key(e) is the
// response, larger first.  Returns the new size to every thread.
template <class T, class Key>
__device__ int par_retain_best(T* v, int n, int n_points, Key key, T* tmp, int W, int* s_cnt, int* s_red)
{
    if (n_points < 0 || n <= n_points) return n;
    if (n_points == 0) return 0;
    par_nth_element(v, n, n_points - 1, [key](const T& a, const T& b) { return key(a) > key(b); }, tmp, W, s_cnt, s_red);
    const auto amb = key(v[n_points - 1]);
    // std::partition(v + n_points, v + n, key >= amb): left pointer stops at !pred, right pointer at pred
    const int m = n - n_points;
    if (m <= PAR_SEQ_CUT) {
        __shared__ int s_end;
        if (threadIdx.x == 0) s_end = (int)(stdalgo::partition(v + n_points, v + n, [key, amb](const T& e) { return key(e) >= amb; }) - v);
        __syncthreads();
        const int r = s_end;
        __syncthreads();
        return r;
    }
    par_pair_swap(v, n_points, n, [key, amb](const T& e) { return !(key(e) >= amb); }, [key, amb](const T& e) { return key(e) >= amb; }, tmp, W,
                  s_cnt, s_red);
    const int npred = s_red[5];  // total number of right stops = elements with pred true (left there by par_pair_swap)
    __syncthreads();
    return n_points + npred;
}

// One CTA per (level, stream): cv::ORB's computeKeyPoints for that level after cv::FAST.
//   list1 (u32: response << 24 | pixel index)  = the FAST output inside the 31-px border in raster order
//   retainBest(list1, 2 N) by one thread, Harris responses by all, retainBest(list2, N) by one thread
__global__ void __launch_bounds__(SEL_THREADS) k_getrt_select(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ kept,
                                                              size_t stride_b, const int* __restrict__ rowcnt, size_t rowcnt_stride,
                                                              SelArgs a, uint2* __restrict__ sel, int* __restrict__ sel_n,
                                                              int* __restrict__ err)
{
    extern __shared__ __align__(16) unsigned char sel_sm[];
    unsigned* list1 = reinterpret_cast<unsigned*>(sel_sm);
    HarrisEntry* list2 = reinterpret_cast<HarrisEntry*>(sel_sm + (size_t)a.n1_cap_max * 4);
    int* rowoff = reinterpret_cast<int*>(list2);  // scratch until list2 is filled (h <= 2 * n2_cap ints)
    __shared__ int s_total, s_cnt[2 * SEL_THREADS], s_red[8];
    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const SelLevel L = a.lv[l];
    const int edge = a.edge;
    int* out_n = sel_n + b * a.nlevels + l;
    if (min(L.w, L.h) <= 2 * edge) {  // cv::ORB skips levels that small
        if (tid == 0) *out_n = 0;
        return;
    }
    const int rows = L.h - 2 * edge;
    const int* rc = rowcnt + (size_t)b * rowcnt_stride + L.row_off + edge;
    for (int r = tid; r < rows; r += SEL_THREADS) rowoff[r] = rc[r];
    int n1 = sel_block_scan(rowoff, rows, &s_total);
    if (n1 > L.n1_cap) {
        if (tid == 0) atomicOr(err, 4);
        n1 = L.n1_cap;
    }
    const uint8_t* kp = kept + (size_t)b * stride_b + L.off;
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int r = warp; r < rows; r += SEL_THREADS / 32) {
            int pos = rowoff[r];
            const int y = edge + r;
            // six segments of 32 pixels per pass: their loads are in flight together (one dependent load per segment made
            // this scan latency bound: 18 round trips to L2 / HBM per row)
            constexpr int SEG = 6;
            const uint8_t* row = kp + (size_t)y * L.w;
            for (int x0 = edge; x0 < L.w - edge; x0 += 32 * SEG) {
                int v[SEG];
#pragma unroll
                for (int u = 0; u < SEG; ++u) {
                    const int x = x0 + 32 * u + lane;
                    v[u] = x < L.w - edge ? (int)__ldg(row + x) : 0;
                }
#pragma unroll
                for (int u = 0; u < SEG; ++u) {
                    const int x = x0 + 32 * u + lane;
                    const unsigned bal = __ballot_sync(0xffffffffu, v[u] > 0);
                    if (v[u] > 0) {
                        const int p = pos + __popc(bal & ((1u << lane) - 1));
                        if (p < L.n1_cap) list1[p] = ((unsigned)(v[u] - 1) << 24) | (unsigned)(y * L.w + x);
                    }
                    pos += __popc(bal);
                }
            }
        }
    }
    __syncthreads();
    // stage 1: the scratch of the partition lives in the (still unused) list2 region
    int n2 = par_retain_best(list1, n1, 2 * L.N, [](const unsigned& e) { return e >> 24; }, reinterpret_cast<unsigned*>(list2), a.n2_cap, s_cnt,
                             s_red);
    if (n2 > a.n2_cap) {
        if (tid == 0) atomicOr(err, 8);
        n2 = a.n2_cap;
    }
    const uint8_t* img = pyr + (size_t)b * stride_b + L.off;
    for (int i = tid; i < n2; i += SEL_THREADS) {
        const unsigned idx = list1[i] & 0xFFFFFFu;
        const int y = (int)(idx / (unsigned)L.w), x = (int)(idx - (unsigned)y * L.w);
        HarrisEntry e;
        e.r = harris_at(img, L.w, x, y, a.s4);
        e.idx = idx;
        list2[i] = e;
    }
    __syncthreads();
    // stage 2: list1 is no longer needed — its region is the scratch (n1_cap_max * 4 bytes >= 2 * W * 8)
    int n3 = par_retain_best(list2, n2, L.N, [](const HarrisEntry& e) { return e.r; }, reinterpret_cast<HarrisEntry*>(list1), a.n1_cap_max / 4,
                             s_cnt, s_red);
    if (n3 > a.sel_cap) {
        if (tid == 0) atomicOr(err, 16);
        n3 = a.sel_cap;
    }
    uint2* so = sel + ((size_t)b * a.nlevels + l) * a.sel_cap;
    for (int i = tid; i < n3; i += SEL_THREADS) so[i] = make_uint2(list2[i].idx, __float_as_uint(list2[i].r));
    if (tid == 0) *out_n = n3;
}

// ------------------------------------------------------------------------------------------------ Hamming nearest neighbour
// Both directions of BFMatcher(crossCheck = true) from ONE pass over the distance matrix (ncu of the two-pass form: the
// POPC pipe 97 % busy — the Hamming distances themselves are the bound, so each is computed once).
// blockIdx = (tile of 128 ref features, stream).  One thread per ref descriptor (eight words in registers); the cur
// descriptors go through shared memory in rounds of 512 (broadcast reads).  Per pair:
//   direction 0 (query = ref i, train = cur j): running minimum in the thread, strict '<' in ascending j
//   direction 1 (query = cur j, train = ref i): key = distance << 16 | i, minimum over the warp by one REDUX, over the CTA by a
//     shared-memory atomicMin per warp, over the CTAs by a global atomicMin per cur feature -> smallest distance, then the
//     smallest ref index: the element cv::batchDistance keeps.  The warp only reduces when one of its keys beats the bound
//     already known for the column (seeded from the global keys).  k_getrt_nn_finish unpacks the keys.
constexpr int NN_THREADS = 128, NN_TILE = 512;
__global__ void __launch_bounds__(NN_THREADS) k_getrt_nn(const uint4* __restrict__ desc_ref, const uint4* __restrict__ desc_cur,
                                                         const int* __restrict__ n_ref, const int* __restrict__ n_cur, int feat_cap,
                                                         int* __restrict__ nn, int* __restrict__ dd, unsigned* __restrict__ colkey)
{
    __shared__ uint4 t0[NN_TILE], t1[NN_TILE];
    __shared__ unsigned s_col[NN_TILE];
    const int b = blockIdx.y;
    const int nq = min(n_ref[b], feat_cap), nt = min(n_cur[b], feat_cap);
    if ((int)blockIdx.x * NN_THREADS >= nq) return;
    const uint4* q = desc_ref + (size_t)b * feat_cap * 2;
    const uint4* t = desc_cur + (size_t)b * feat_cap * 2;
    const int i = blockIdx.x * NN_THREADS + threadIdx.x, lane = threadIdx.x & 31;
    const bool valid = i < nq;
    uint4 a0 = make_uint4(0, 0, 0, 0), a1 = a0;
    if (valid) {
        a0 = q[2 * i];
        a1 = q[2 * i + 1];
    }
    const unsigned kinv = valid ? 0u : 0xFFFFFFFFu;  // rows past the end never win a column
    int best = 1 << 30, bi = -1;
    for (int base = 0; base < nt; base += NN_TILE) {
        const int m = min(NN_TILE, nt - base);
        unsigned* ck = colkey + (size_t)b * feat_cap + base;
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += NN_THREADS) {
            t0[k] = t[2 * (base + k)];
            t1[k] = t[2 * (base + k) + 1];
            s_col[k] = ck[k];  // what the other CTAs have found so far: an upper bound that makes most updates unnecessary
        }
        __syncthreads();
#pragma unroll 2
        for (int j = 0; j < m; ++j) {
            const uint4 b0 = t0[j], b1 = t1[j];
            const int d = __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) +
                          __popc(a1.y ^ b1.y) + __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
            if (d < best) {
                best = d;
                bi = base + j;
            }
            const unsigned key = (((unsigned)d << 16) | (unsigned)i) | kinv;
            if (__any_sync(0xffffffffu, key < s_col[j])) {  // (a stale bound only costs an unnecessary update)
                const unsigned kmin = __reduce_min_sync(0xffffffffu, key);
                if (lane == 0) atomicMin(&s_col[j], kmin);
            }
        }
        __syncthreads();
        for (int k = threadIdx.x; k < m; k += NN_THREADS) atomicMin(ck + k, s_col[k]);
    }
    if (valid) {
        nn[((size_t)b * 2 + 0) * feat_cap + i] = bi;
        dd[((size_t)b * 2 + 0) * feat_cap + i] = best;
    }
}

// ---- the same on the tensor cores.  Hamming(a, b) = popc(a) + popc(b) - 2 <a, b> with the descriptors expanded to 256
// bytes of 0 / 1: the inner products of a 32 x 8 block of (ref, cur) pairs are two m16n8k32 u8 MMAs per 32 bits of
// descriptor (IMMA.16832.U8 in the SASS), the POPC pipe — the bound of the scalar form — is only used once per descriptor.
// A CTA owns 128 ref features (one warp = 32 of them, their A fragments stay in registers: 2 x 8 x 4 words) and walks the
// cur features in rounds of 128 that are expanded into shared memory.  The order of the k index inside an MMA is free as long
// as A and B agree (a distance does not depend on the order of the bits): thread `tig` of a quad takes bytes 8 tig .. 8 tig + 7
// of every 32-byte chunk, (a0, a2) / (a1, a3) / (b0, b1) are one 64-bit load each; the row pitch of 288 bytes makes those
// loads conflict free.  Minima as in k_getrt_nn: rows in the thread (columns ascending, strict '<', quad merge by
// (distance, index)), columns through packed keys with a rarely taken shared atomicMin.
constexpr int HM_TILE = 128, HM_PITCH = 288, HM_THREADS = 128;

__device__ __forceinline__ int hm_expand(uint8_t* dst, const uint4 lo, const uint4 hi)
{
    const unsigned w[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
    int pc = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        pc += __popc(w[k]);
        unsigned o[8];
#pragma unroll
        for (int n = 0; n < 8; ++n) o[n] = (((w[k] >> (4 * n)) & 15u) * 0x00204081u) & 0x01010101u;  // bit q of the nibble -> byte q
        uint4* d4 = reinterpret_cast<uint4*>(dst + 32 * k);
        d4[0] = make_uint4(o[0], o[1], o[2], o[3]);
        d4[1] = make_uint4(o[4], o[5], o[6], o[7]);
    }
    return pc;
}

__device__ __forceinline__ void hm_mma(int (&c)[4], const unsigned (&a)[4], const uint2 b)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

__global__ void __launch_bounds__(HM_THREADS) k_getrt_nn_mma(const uint4* __restrict__ desc_ref, const uint4* __restrict__ desc_cur,
                                                             const int* __restrict__ n_ref, const int* __restrict__ n_cur, int feat_cap,
                                                             int* __restrict__ nn, int* __restrict__ dd, unsigned* __restrict__ colkey)
{
    __shared__ __align__(16) uint8_t hm[HM_TILE * HM_PITCH];
    __shared__ int s_pc[HM_TILE];
    __shared__ unsigned s_col[HM_TILE];
    const int b = blockIdx.y;
    const int nq = min(n_ref[b], feat_cap), nt = min(n_cur[b], feat_cap);
    const int row0 = blockIdx.x * HM_TILE;
    if (row0 >= nq) return;
    const uint4* q = desc_ref + (size_t)b * feat_cap * 2;
    const uint4* t = desc_cur + (size_t)b * feat_cap * 2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, g = lane >> 2, tig = lane & 3;
    const uint4 zero4 = make_uint4(0, 0, 0, 0);
    {   // the CTA's ref descriptors: expanded once, then only their fragments (registers) and popcounts are kept
        const bool v = row0 + tid < nq;
        s_pc[tid] = hm_expand(hm + tid * HM_PITCH, v ? q[2 * (row0 + tid)] : zero4, v ? q[2 * (row0 + tid) + 1] : zero4);
    }
    __syncthreads();
    unsigned A[2][8][4];
    int pa[4], rowi[4];
    unsigned kinv[4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
            const uint2 lo = *reinterpret_cast<const uint2*>(hm + (warp * 32 + 16 * mt + g) * HM_PITCH + 32 * ks + 8 * tig);
            const uint2 hi = *reinterpret_cast<const uint2*>(hm + (warp * 32 + 16 * mt + g + 8) * HM_PITCH + 32 * ks + 8 * tig);
            A[mt][ks][0] = lo.x; A[mt][ks][2] = lo.y;
            A[mt][ks][1] = hi.x; A[mt][ks][3] = hi.y;
        }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int lr = warp * 32 + g + 8 * r;  // r = 2 mt + half: rows g, g + 8 of m-tile 0, then of m-tile 1
        pa[r] = s_pc[lr];
        rowi[r] = row0 + lr;
        kinv[r] = rowi[r] < nq ? 0u : 0xFFFFFFFFu;
    }
    int best[4], bj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        best[r] = 1 << 30;
        bj[r] = -1;
    }
    for (int base = 0; base < nt; base += HM_TILE) {
        const int m = min(HM_TILE, nt - base);
        unsigned* ck = colkey + (size_t)b * feat_cap + base;
        __syncthreads();  // fragments loaded / previous round done with the tile
        {
            const bool v = tid < m;
            const int pc = hm_expand(hm + tid * HM_PITCH, v ? t[2 * (base + tid)] : zero4, v ? t[2 * (base + tid) + 1] : zero4);
            s_pc[tid] = v ? pc : (1 << 30);  // a column past the end never beats a row minimum ...
            s_col[tid] = v ? ck[tid] : 0u;   // ... and no key is below its bound
        }
        __syncthreads();
        // two column tiles per pass: four independent accumulator chains per thread (the tile always holds 16 column tiles;
        // columns past the end are zero descriptors with an unreachable popcount)
        const int ntile = (m + 7) >> 3;
        for (int n8 = 0; n8 < ntile; n8 += 2) {
            int c[2][2][4] = {{{0, 0, 0, 0}, {0, 0, 0, 0}}, {{0, 0, 0, 0}, {0, 0, 0, 0}}};  // [column tile][m-tile][4]
            const uint8_t* bp = hm + (8 * n8 + g) * HM_PITCH + 8 * tig;
#pragma unroll
            for (int ks = 0; ks < 8; ++ks) {
                const uint2 b0 = *reinterpret_cast<const uint2*>(bp + 32 * ks);
                const uint2 b1 = *reinterpret_cast<const uint2*>(bp + 8 * HM_PITCH + 32 * ks);
                hm_mma(c[0][0], A[0][ks], b0);
                hm_mma(c[0][1], A[1][ks], b0);
                hm_mma(c[1][0], A[0][ks], b1);
                hm_mma(c[1][1], A[1][ks], b1);
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int col = 8 * (n8 + u) + 2 * tig;
                const int pb0 = s_pc[col], pb1 = s_pc[col + 1];
                const unsigned bound0 = s_col[col], bound1 = s_col[col + 1];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int d0 = pa[r] + pb0 - 2 * c[u][r >> 1][2 * (r & 1)], d1 = pa[r] + pb1 - 2 * c[u][r >> 1][2 * (r & 1) + 1];
                    if (d0 < best[r]) {
                        best[r] = d0;
                        bj[r] = base + col;
                    }
                    if (d1 < best[r]) {
                        best[r] = d1;
                        bj[r] = base + col + 1;
                    }
                    const unsigned key0 = (((unsigned)d0 << 16) | (unsigned)rowi[r]) | kinv[r];
                    const unsigned key1 = (((unsigned)d1 << 16) | (unsigned)rowi[r]) | kinv[r];
                    if (key0 < bound0) atomicMin(&s_col[col], key0);
                    if (key1 < bound1) atomicMin(&s_col[col + 1], key1);
                }
            }
        }
        __syncthreads();
        if (tid < m) atomicMin(ck + tid, s_col[tid]);
    }
    // merge the four threads of a quad (they hold disjoint columns of the same rows)
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int off = 1; off < 4; off <<= 1) {
            const int ob = __shfl_xor_sync(0xffffffffu, best[r], off), oi = __shfl_xor_sync(0xffffffffu, bj[r], off);
            if (ob < best[r] || (ob == best[r] && oi < bj[r])) {
                best[r] = ob;
                bj[r] = oi;
            }
        }
        if (tig == 0 && rowi[r] < nq) {
            nn[((size_t)b * 2 + 0) * feat_cap + rowi[r]] = bj[r];
            dd[((size_t)b * 2 + 0) * feat_cap + rowi[r]] = best[r];
        }
    }
}

// direction 1 from the column keys (0xFFFFFFFF: no ref feature at all -> index -1, distance 1 << 30 like an empty scan)
__global__ void __launch_bounds__(256) k_getrt_nn_finish(const unsigned* __restrict__ colkey, const int* __restrict__ n_cur, int feat_cap,
                                                         int* __restrict__ nn, int* __restrict__ dd)
{
    const int b = blockIdx.y, j = blockIdx.x * 256 + threadIdx.x;
    if (j >= min(n_cur[b], feat_cap)) return;
    const unsigned key = colkey[(size_t)b * feat_cap + j];
    const bool none = key == 0xFFFFFFFFu;
    nn[((size_t)b * 2 + 1) * feat_cap + j] = none ? -1 : (int)(key & 0xFFFFu);
    dd[((size_t)b * 2 + 1) * feat_cap + j] = none ? (1 << 30) : (int)(key >> 16);
}

static int launch_hamming_both(const uint4* desc_ref, const uint4* desc_cur, const int* n_ref, const int* n_cur, int feat_cap, int batch,
                               int* nn, int* dd, unsigned* colkey, cudaStream_t s)
{
    GD_REQUIRE(feat_cap <= 65536, "ref feature index does not fit the column key");
    GD_CUDA(cudaMemsetAsync(colkey, 0xFF, (size_t)batch * feat_cap * sizeof(unsigned), s));
    static const bool scalar = [] {
        const char* e = std::getenv("GD_GETRT_NN_SCALAR");  // the POPC form, kept for A/B runs and as the in-tree cross-check
        return e && std::atoi(e) != 0;
    }();
    if (scalar)
        k_getrt_nn<<<dim3(cdiv(feat_cap, NN_THREADS), batch), NN_THREADS, 0, s>>>(desc_ref, desc_cur, n_ref, n_cur, feat_cap, nn, dd, colkey);
    else
        k_getrt_nn_mma<<<dim3(cdiv(feat_cap, HM_TILE), batch), HM_THREADS, 0, s>>>(desc_ref, desc_cur, n_ref, n_cur, feat_cap, nn, dd, colkey);
    GD_CUDA(cudaGetLastError());
    k_getrt_nn_finish<<<dim3(cdiv(feat_cap, 256), batch), 256, 0, s>>>(colkey, n_cur, feat_cap, nn, dd);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------ matches -> points
struct PointArgs {
    int w, h, feat_cap;
    UndistortArgs und;      // (double) of the f32 K and of k1 k2 p1 p2 k3
    float Ki[9];            // inv(K) in f32
};
struct MatchEntry {
    float d;
    int i;
};
constexpr int PT_THREADS = 256;

// One CTA per stream.  Matches come out of BFMatcher ordered by query index; GetRt sorts them (std::sort on the distance,
// unstable) and keeps the first 100; points whose depth is zero are dropped (GeoMaskMaker.cc:95-141).
__global__ void __launch_bounds__(PT_THREADS) k_getrt_points(const gd_keypoint* __restrict__ kp_ref, const gd_keypoint* __restrict__ kp_cur,
                                                             const int* __restrict__ n_ref, const int* __restrict__ n_cur,
                                                             const int* __restrict__ nn, const int* __restrict__ dd,
                                                             const float* __restrict__ depth_ref, size_t depth_stride_b, PointArgs a,
                                                             float* __restrict__ out_obj, float* __restrict__ out_pix,
                                                             int* __restrict__ out_cnt)
{
    extern __shared__ __align__(16) unsigned char pt_sm[];
    MatchEntry* ms = reinterpret_cast<MatchEntry*>(pt_sm);            // [feat_cap]
    int* mq = reinterpret_cast<int*>(pt_sm + (size_t)a.feat_cap * 8);  // [feat_cap] query index of match i
    __shared__ int s_warp[PT_THREADS / 32], s_base, s_valid[GETRT_TOP];
    __shared__ float s_obj[GETRT_TOP][3], s_pix[GETRT_TOP][2];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n1 = min(n_ref[b], a.feat_cap), n2 = min(n_cur[b], a.feat_cap);
    const int* nn0 = nn + ((size_t)b * 2 + 0) * a.feat_cap;
    const int* nn1 = nn + ((size_t)b * 2 + 1) * a.feat_cap;
    const int* dd0 = dd + ((size_t)b * 2 + 0) * a.feat_cap;
    if (tid == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n1; base += PT_THREADS) {  // cross check, order preserving compaction
        const int q = base + tid;
        bool ok = false;
        int t = -1;
        if (q < n1 && n2 > 0) {
            t = nn0[q];
            ok = t >= 0 && t < n2 && nn1[t] == q;
        }
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < PT_THREADS / 32; ++w2) {
            if (w2 < warp) woff += s_warp[w2];
            tot += s_warp[w2];
        }
        if (ok) {
            const int m = s_base + woff + __popc(bal & ((1u << lane) - 1));
            ms[m].d = (float)dd0[q];
            ms[m].i = m;
            mq[m] = q;
        }
        __syncthreads();
        if (tid == 0) s_base += tot;
        __syncthreads();
    }
    const int nm = s_base;
    if (tid == 0 && nm > 0)
        stdalgo::sort_prefix(ms, ms + nm, (long)GETRT_TOP, [](const MatchEntry& x, const MatchEntry& y) { return x.d < y.d; });
    __syncthreads();
    // the reference takes begin() + 100 unconditionally (:97); with fewer matches that is undefined there, all of them here
    const int ntop = min(nm, GETRT_TOP);
    if (tid < GETRT_TOP) {
        int valid = 0;
        if (tid < ntop) {
            const int q = mq[ms[tid].i], t = nn0[q];
            const gd_keypoint k1 = kp_ref[(size_t)b * a.feat_cap + q];
            const gd_keypoint k2 = kp_cur[(size_t)b * a.feat_cap + t];
            float ux, uy;
            undistort_point_cv(a.und, k1.x, k1.y, &ux, &uy);  // :104-110 (D = 0: returns the point itself)
            const int dx = (int)ux, dy = (int)uy;      // :122-123
            if (dx >= 0 && dy >= 0 && dx < a.w && dy < a.h) {
                const float depth = depth_ref[(size_t)b * depth_stride_b + (size_t)dy * a.w + dx];
                if (depth != 0.f) {                    // :125-128
                    valid = 1;
#pragma unroll
                    for (int c = 0; c < 3; ++c) {      // (inv(K) * [x y 1]^T) * depth, f32 gemm of inner length 3 (:133)
                        float v = a.Ki[3 * c] * ux;
                        v = v + a.Ki[3 * c + 1] * uy;
                        v = v + a.Ki[3 * c + 2] * 1.0f;
                        s_obj[tid][c] = v * depth;
                    }
                    s_pix[tid][0] = k2.x;              // :139
                    s_pix[tid][1] = k2.y;
                }
            }
        }
        s_valid[tid] = valid;
    }
    __syncthreads();
    if (tid == 0) {  // 100 flags: a serial prefix is cheaper than another scan
        int n = 0;
        for (int r = 0; r < GETRT_TOP; ++r) {
            const int v = s_valid[r];
            s_valid[r] = v ? n : -1;
            n += v;
        }
        out_cnt[b] = n;
    }
    __syncthreads();
    if (tid < GETRT_TOP && s_valid[tid] >= 0) {
        const int o = s_valid[tid];
        for (int c = 0; c < 3; ++c) out_obj[((size_t)b * GETRT_TOP + o) * 3 + c] = s_obj[tid][c];
        out_pix[((size_t)b * GETRT_TOP + o) * 2] = s_pix[tid][0];
        out_pix[((size_t)b * GETRT_TOP + o) * 2 + 1] = s_pix[tid][1];
    }
}

// ------------------------------------------------------------------------------------------------ GetRtCore
int GetRtCore::init(const float K[9], const float* dist_coef, int ndist, int width, int height, int device_, int batch_, int ring_slots,
                    int nfeatures_)
{
    GD_REQUIRE(nfeatures_ >= 8 && nfeatures_ <= 8192, "nfeatures out of range");
    nfeatures = nfeatures_;
    GD_REQUIRE(width >= 2 * GETRT_EDGE + 8 && height >= 2 * GETRT_EDGE + 8 && batch_ >= 1 && ring_slots >= 1, "bad size / batch");
    GD_REQUIRE((long long)width * height < (1ll << 24), "image too large for the 24-bit pixel index of the corner lists");
    GD_TRY(select_device(device_));
    device = device_;
    batch = batch_;
    w = width;
    h = height;
    ring = ring_slots;
    make_cam_const(K, &cam);
    Kd[0] = (double)K[0]; Kd[1] = (double)K[4]; Kd[2] = (double)K[2]; Kd[3] = (double)K[5];
    for (int i = 0; i < 5; ++i) dist[i] = 0.0;
    distorted = false;
    for (int i = 0; i < ndist && i < 5 && dist_coef; ++i) {
        dist[i] = (double)dist_coef[i];
        distorted = distorted || dist_coef[i] != 0.f;
    }
    // cv::ORB: features per level (geometric series, last level takes the remainder), level sizes from the float scale
    const float sf = 1.2f;
    {
        const double factor = 1.0 / (double)sf;
        double nd = nfeatures * (1 - factor) / (1 - std::pow(factor, (double)GETRT_LEVELS));
        int sum = 0;
        for (int l = 0; l < GETRT_LEVELS - 1; ++l) {
            nper[l] = (int)std::nearbyint(nd);
            sum += nper[l];
            nd *= factor;
        }
        nper[GETRT_LEVELS - 1] = std::max(nfeatures - sum, 0);
    }
    pyr_args.nlevels = GETRT_LEVELS;
    size_t off = 0;
    int rows = 0, max_n = 0;
    std::vector<ushort4> tab;
    for (int l = 0; l < GETRT_LEVELS; ++l) {
        CvLevelDev& L = pyr_args.lv[l];
        L.scale = l == 0 ? 1.0f : (float)std::pow((double)sf, (double)l);
        L.w = l == 0 ? w : (int)std::nearbyint((float)w / L.scale);
        L.h = l == 0 ? h : (int)std::nearbyint((float)h / L.scale);
        GD_REQUIRE(L.w >= 8 && L.h >= 8, "image too small for the 8-level cv::ORB pyramid");
        L.off = off;
        off += align_up((size_t)L.w * L.h, 256);
        L.row_off = rows;
        rows += L.h;
        max_n = std::max(max_n, nper[l]);
        if (l > 0) {
            tab_x[l] = (int)tab.size();
            tab.resize(tab.size() + L.w);
            linear_exact_axis_table(L.w, pyr_args.lv[l - 1].w, tab.data() + tab_x[l]);
            tab_y[l] = (int)tab.size();
            tab.resize(tab.size() + L.h);
            linear_exact_axis_table(L.h, pyr_args.lv[l - 1].h, tab.data() + tab_y[l]);
        }
    }
    pyr_bytes = off;
    rows_total = rows;
    // capacities: the NMS map has no two 8-adjacent corners -> at most a quarter of the interior pixels; bounded by what one
    // CTA's shared memory holds (an overflow is reported through the error flag, never silently)
    n1_cap = 0;
    for (int l = 0; l < GETRT_LEVELS; ++l) {
        const CvLevelDev& L = pyr_args.lv[l];
        const long long interior = (long long)std::max(0, L.w - 2 * GETRT_EDGE) * std::max(0, L.h - 2 * GETRT_EDGE);
        // (34816 entries would let one CTA of the flow's box kernel share the SM with a selection CTA: measured, no gain)
        n1_cap = std::max<long long>(n1_cap, std::min<long long>((interior + 3) / 4 + 32, 36864));
    }
    n2_cap = 2 * max_n + 4096;
    sel_cap = max_n + 64;
    feat_cap = (int)align_up((size_t)nfeatures + 256, 128);
    select_smem = (size_t)n1_cap * 4 + (size_t)n2_cap * 8;
    GD_REQUIRE(select_smem <= 200 * 1024 && 2 * n2_cap >= h, "selection kernel's shared memory plan does not fit");
    GD_CUDA(cudaFuncSetAttribute(k_getrt_select, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)select_smem));
    GD_REQUIRE((size_t)feat_cap * 12 <= 48 * 1024, "match list larger than the default shared memory");
    const size_t B = (size_t)batch;
    GD_TRY(pyr.alloc(B * pyr_bytes));
    GD_TRY(kept.alloc(B * pyr_bytes));
    GD_TRY(blur.alloc(B * pyr_bytes));
    GD_TRY(rowcnt.alloc(B * rows_total * sizeof(int)));
    GD_TRY(tabs.alloc(std::max<size_t>(1, tab.size()) * sizeof(ushort4)));
    GD_TRY(sel.alloc(B * GETRT_LEVELS * sel_cap * sizeof(uint2)));
    GD_TRY(sel_n.alloc(B * GETRT_LEVELS * sizeof(int)));
    GD_TRY(feat_kp.alloc((size_t)ring * B * feat_cap * sizeof(gd_keypoint)));
    GD_TRY(feat_desc.alloc((size_t)ring * B * feat_cap * 32));
    GD_TRY(feat_n.alloc((size_t)ring * B * sizeof(int)));
    GD_TRY(nn.alloc(B * 2 * feat_cap * sizeof(int)));
    GD_TRY(dd.alloc(B * 2 * feat_cap * sizeof(int)));
    GD_TRY(colkey.alloc(B * feat_cap * sizeof(unsigned)));
    GD_TRY(out_obj.alloc(B * GETRT_TOP * 3 * sizeof(float)));
    GD_TRY(out_pix.alloc(B * GETRT_TOP * 2 * sizeof(float)));
    GD_TRY(out_cnt.alloc((B + 1) * sizeof(int)));
    GD_TRY(err.alloc(sizeof(int)));
    GD_TRY(h_out.alloc(B * GETRT_TOP * 5 * sizeof(float) + (B + 1) * sizeof(int)));
    GD_CUDA(cudaMemcpy(tabs.p, tab.data(), tab.size() * sizeof(ushort4), cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemset(feat_n.p, 0, feat_n.bytes));
    GD_CUDA(cudaMemset(err.p, 0, sizeof(int)));
    GD_CUDA(cudaMemset(out_cnt.p, 0, out_cnt.bytes));
    return GD_OK;
}

int GetRtCore::enqueue_features(const uint8_t* gray, size_t gray_stride_b, int slot)
{
    GD_REQUIRE(slot >= 0 && slot < ring, "ring slot out of range");
    const cudaStream_t s = stream;
    uint8_t* py = pyr.as<uint8_t>();
    {
        LaunchScope ls(stats, s, "G1_cvorb_pyramid", GETRT_LEVELS - 1);
        GD_CUDA(cudaMemcpy2DAsync(py, pyr_bytes, gray, gray_stride_b, (size_t)w * h, batch, cudaMemcpyDeviceToDevice, s));
        const ushort4* t = tabs.as<ushort4>();
        for (int l = 1; l < GETRT_LEVELS; ++l) {
            const CvLevelDev& S = pyr_args.lv[l - 1];
            const CvLevelDev& D = pyr_args.lv[l];
            k_resize_linear_exact<<<dim3(cdiv(D.w, 32), cdiv(D.h, 8), batch), dim3(32, 8), 0, s>>>(py + S.off, pyr_bytes, S.w, py + D.off, pyr_bytes,
                                                                                              D.w, D.h, t + tab_x[l], t + tab_y[l]);
            GD_CUDA(cudaGetLastError());
        }
    }
    {
        LaunchScope ls(stats, s, "G2_cvorb_fast", 1);
        GD_CUDA(cudaMemsetAsync(rowcnt.p, 0, rowcnt.bytes, s));
        GD_TRY(orb_cv_fast_levels(py, pyr_bytes, pyr_args, batch, 20, GETRT_EDGE, nullptr, kept.as<uint8_t>(), rowcnt.as<int>(),
                                  (size_t)rows_total, s));
    }
    {
        LaunchScope ls(stats, s, "G3_cvorb_select", 1);
        SelArgs a;
        a.nlevels = GETRT_LEVELS;
        a.edge = GETRT_EDGE;
        a.n2_cap = n2_cap;
        a.sel_cap = sel_cap;
        a.n1_cap_max = n1_cap;
        a.s4 = harris_scale4();
        for (int l = 0; l < GETRT_LEVELS; ++l) {
            const CvLevelDev& L = pyr_args.lv[l];
            const long long interior = (long long)std::max(0, L.w - 2 * GETRT_EDGE) * std::max(0, L.h - 2 * GETRT_EDGE);
            a.lv[l] = {L.w, L.h, L.off, L.row_off, nper[l], (int)std::min<long long>((interior + 3) / 4 + 32, n1_cap)};
        }
        k_getrt_select<<<dim3(GETRT_LEVELS, batch), SEL_THREADS, select_smem, s>>>(py, kept.as<uint8_t>(), pyr_bytes, rowcnt.as<int>(),
                                                                               (size_t)rows_total, a, sel.as<uint2>(), sel_n.as<int>(),
                                                                               err.as<int>());
        GD_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls(stats, s, "G4_cvorb_blur", 1);
        int tiles = 0;
        for (int l = 0; l < GETRT_LEVELS; ++l)
            tiles = std::max(tiles, cdiv(pyr_args.lv[l].w, GF_COLS) * cdiv(pyr_args.lv[l].h, GF_ROWS * GF_WARPS));
        k_gaussian7_float_levels<<<dim3(tiles, GETRT_LEVELS, batch), dim3(32, GF_WARPS), 0, s>>>(py, pyr_bytes, pyr_args, gauss7_taps(),
                                                                                          blur.as<uint8_t>());
        GD_CUDA(cudaGetLastError());
    }
    {
        LaunchScope ls(stats, s, "G5_cvorb_describe", 1);
        GD_TRY(orb_cv_describe_sel(py, blur.as<uint8_t>(), pyr_bytes, pyr_args, batch, sel.as<uint2>(), sel_cap, sel_n.as<int>(), slot_kp(slot),
                                   slot_desc(slot), slot_n(slot), feat_cap, s));
    }
    return GD_OK;
}

int GetRtCore::enqueue_match(int ref_slot, int cur_slot, const float* depth_ref, size_t depth_stride_b)
{
    GD_REQUIRE(ref_slot >= 0 && ref_slot < ring && cur_slot >= 0 && cur_slot < ring && depth_ref, "bad argument");
    const cudaStream_t s = stream;
    {
        LaunchScope ls(stats, s, "G6_match", 2);
        GD_TRY(launch_hamming_both(reinterpret_cast<const uint4*>(slot_desc(ref_slot)), reinterpret_cast<const uint4*>(slot_desc(cur_slot)),
                                   slot_n(ref_slot), slot_n(cur_slot), feat_cap, batch, nn.as<int>(), dd.as<int>(), colkey.as<unsigned>(), s));
    }
    {
        LaunchScope ls(stats, s, "G7_points", 1);
        PointArgs a;
        a.w = w; a.h = h; a.feat_cap = feat_cap;
        a.und.fx = Kd[0]; a.und.fy = Kd[1]; a.und.cx = Kd[2]; a.und.cy = Kd[3];
        for (int i = 0; i < 5; ++i) a.und.k[i] = dist[i];
        for (int i = 0; i < 9; ++i) a.Ki[i] = cam.Ki[i];
        k_getrt_points<<<batch, PT_THREADS, (size_t)feat_cap * 12, s>>>(slot_kp(ref_slot), slot_kp(cur_slot), slot_n(ref_slot), slot_n(cur_slot),
                                                                      nn.as<int>(), dd.as<int>(), depth_ref, depth_stride_b, a,
                                                                      out_obj.as<float>(), out_pix.as<float>(), out_cnt.as<int>());
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

int GetRtCore::enqueue_fetch()
{
    const cudaStream_t s = stream;
    const size_t B = (size_t)batch;
    float* hp = h_out.as<float>();
    GD_CUDA(cudaMemcpyAsync(hp, out_obj.p, B * GETRT_TOP * 3 * sizeof(float), cudaMemcpyDeviceToHost, s));
    GD_CUDA(cudaMemcpyAsync(hp + B * GETRT_TOP * 3, out_pix.p, B * GETRT_TOP * 2 * sizeof(float), cudaMemcpyDeviceToHost, s));
    GD_CUDA(cudaMemcpyAsync(hp + B * GETRT_TOP * 5, out_cnt.p, B * sizeof(int), cudaMemcpyDeviceToHost, s));
    GD_CUDA(cudaMemcpyAsync(reinterpret_cast<int*>(hp + B * GETRT_TOP * 5) + B, err.p, sizeof(int), cudaMemcpyDeviceToHost, s));
    return GD_OK;
}

}  // namespace gd

using namespace gd;

// ------------------------------------------------------------------------------------------------ stage entry points
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
    k_resize_linear_exact<<<dim3(cdiv(dw, 32), cdiv(dh, 8), 1), dim3(32, 8)>>>(s.as<uint8_t>(), 0, sw, d.as<uint8_t>(), 0, dw, dh, t.as<ushort4>(),
                                                                             t.as<ushort4>() + dw);
    GD_CUDA(cudaGetLastError());
    GD_CUDA(cudaMemcpy(dst, d.p, (size_t)dw * dh, cudaMemcpyDeviceToHost));
    return GD_OK;
}

int gd_stage_gaussian7_float(int device, const uint8_t* src, int w, int h, uint8_t* dst)
{
    GD_REQUIRE(src && dst && w > 3 && h > 3, "bad argument");
    GD_TRY(select_device(device));
    DevBuf s, d;
    GD_TRY(s.alloc((size_t)w * h));
    GD_TRY(d.alloc((size_t)w * h));
    GD_CUDA(cudaMemcpy(s.p, src, (size_t)w * h, cudaMemcpyHostToDevice));
    CvPyrArgs a;
    a.nlevels = 1;
    a.lv[0] = {w, h, 0ull, 0, 1.0f};
    k_gaussian7_float_levels<<<dim3(cdiv(w, GF_COLS) * cdiv(h, GF_ROWS * GF_WARPS), 1, 1), dim3(32, GF_WARPS)>>>(s.as<uint8_t>(), 0, a, gauss7_taps(), d.as<uint8_t>());
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
    DevBuf im, dx, dy, o;
    GD_TRY(im.alloc((size_t)w * h));
    GD_TRY(dx.alloc(sizeof(int) * n));
    GD_TRY(dy.alloc(sizeof(int) * n));
    GD_TRY(o.alloc(sizeof(float) * n));
    GD_CUDA(cudaMemcpy(im.p, img, (size_t)w * h, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dx.p, xs, sizeof(int) * n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(dy.p, ys, sizeof(int) * n, cudaMemcpyHostToDevice));
    k_harris<<<cdiv(n, 128), 128>>>(im.as<uint8_t>(), w, dx.as<int>(), dy.as<int>(), n, harris_scale4(), o.as<float>());
    GD_CUDA(cudaGetLastError());
    GD_CUDA(cudaMemcpy(out, o.p, sizeof(float) * n, cudaMemcpyDeviceToHost));
    return GD_OK;
}

// BFMatcher(NORM_HAMMING, crossCheck = true)->match on two host descriptor sets: the matching kernel of the resident stage
// (both directions) + the cross check in query order on the host
int gd_stage_hamming_crosscheck(int device, const uint8_t* d1, int n1, const uint8_t* d2, int n2, int* query_idx, int* train_idx,
                                int* distance, int capacity, int* n_matches)
{
    GD_REQUIRE(d1 && d2 && query_idx && train_idx && distance && n_matches && n1 > 0 && n2 > 0, "bad argument");
    GD_TRY(select_device(device));
    const int cap = std::max(n1, n2);
    DevBuf a, b, cnt, dnn, ddd;
    GD_TRY(a.alloc((size_t)cap * 32));
    GD_TRY(b.alloc((size_t)cap * 32));
    GD_TRY(cnt.alloc(2 * sizeof(int)));
    GD_TRY(dnn.alloc(sizeof(int) * 2 * cap));
    GD_TRY(ddd.alloc(sizeof(int) * 2 * cap));
    GD_CUDA(cudaMemcpy(a.p, d1, (size_t)n1 * 32, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(b.p, d2, (size_t)n2 * 32, cudaMemcpyHostToDevice));
    const int hn[2] = {n1, n2};
    GD_CUDA(cudaMemcpy(cnt.p, hn, sizeof(hn), cudaMemcpyHostToDevice));
    DevBuf ck;
    GD_TRY(ck.alloc(sizeof(unsigned) * cap));
    GD_TRY(launch_hamming_both(a.as<uint4>(), b.as<uint4>(), cnt.as<int>(), cnt.as<int>() + 1, cap, 1, dnn.as<int>(), ddd.as<int>(),
                               ck.as<unsigned>(), 0));
    GD_CUDA(cudaDeviceSynchronize());
    std::vector<int> h12(n1), hd(n1), h21(n2);
    GD_CUDA(cudaMemcpy(h12.data(), dnn.p, sizeof(int) * n1, cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(hd.data(), ddd.p, sizeof(int) * n1, cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(h21.data(), dnn.as<int>() + cap, sizeof(int) * n2, cudaMemcpyDeviceToHost));
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

// cv::ORB(2000, 1.2, 8, 31, 0, 2)->detectAndCompute of one host image through the resident stage (ring of one slot)
int gd_stage_cvorb_detect_and_compute(int device, const uint8_t* gray, int w, int h, int nfeatures, gd_keypoint* kps, uint8_t* desc,
                                      int capacity, int* n)
{
    GD_REQUIRE(gray && kps && desc && n && w > 0 && h > 0, "bad argument");
    GD_TRY(select_device(device));
    const float K[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    GetRtCore c;
    GD_TRY(c.init(K, nullptr, 0, w, h, device, 1, 1, nfeatures));
    DevBuf g;
    GD_TRY(g.alloc((size_t)w * h));
    GD_CUDA(cudaMemcpy(g.p, gray, (size_t)w * h, cudaMemcpyHostToDevice));
    GD_TRY(c.enqueue_features(g.as<uint8_t>(), (size_t)w * h, 0));
    GD_CUDA(cudaDeviceSynchronize());
    int cnt = 0, e = 0;
    GD_CUDA(cudaMemcpy(&cnt, c.slot_n(0), sizeof(int), cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(&e, c.err.p, sizeof(int), cudaMemcpyDeviceToHost));
    if (e != 0) {
        set_error("cv::ORB stage: capacity overflow of the selection lists (flags %d)", e);
        return GD_EINTERNAL;
    }
    *n = cnt;
    GD_REQUIRE(cnt <= capacity, "keypoint capacity too small");
    GD_CUDA(cudaMemcpy(kps, c.slot_kp(0), sizeof(gd_keypoint) * cnt, cudaMemcpyDeviceToHost));
    GD_CUDA(cudaMemcpy(desc, c.slot_desc(0), (size_t)32 * cnt, cudaMemcpyDeviceToHost));
    return GD_OK;
}

// GetRt up to solvePnPRansac for one host image pair (ring of two slots)
int gd_getrt_points(int device, const uint8_t* gray_first, const uint8_t* gray_second, int w, int h, const float* depth_first_m,
                    const float K[9], const float* dist, int ndist, float* object_points, float* image_pixels, int* n_points)
{
    GD_REQUIRE(gray_first && gray_second && depth_first_m && K && object_points && image_pixels && n_points, "null argument");
    GD_TRY(select_device(device));
    *n_points = 0;
    GetRtCore c;
    GD_TRY(c.init(K, dist, ndist, w, h, device, 1, 2));
    DevBuf g, d;
    const size_t n = (size_t)w * h;
    GD_TRY(g.alloc(2 * n));
    GD_TRY(d.alloc(n * sizeof(float)));
    GD_CUDA(cudaMemcpy(g.p, gray_first, n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(g.as<uint8_t>() + n, gray_second, n, cudaMemcpyHostToDevice));
    GD_CUDA(cudaMemcpy(d.p, depth_first_m, n * sizeof(float), cudaMemcpyHostToDevice));
    GD_TRY(c.enqueue_features(g.as<uint8_t>(), n, 0));
    GD_TRY(c.enqueue_features(g.as<uint8_t>() + n, n, 1));
    GD_TRY(c.enqueue_match(0, 1, d.as<float>(), n));
    GD_TRY(c.enqueue_fetch());
    GD_CUDA(cudaDeviceSynchronize());
    if (c.host_err() != 0) {
        set_error("GetRt stage: capacity overflow of the selection lists (flags %d)", c.host_err());
        return GD_EINTERNAL;
    }
    const int cnt = c.host_cnt()[0];
    std::memcpy(object_points, c.host_obj(0), sizeof(float) * 3 * cnt);
    std::memcpy(image_pixels, c.host_pix(0), sizeof(float) * 2 * cnt);
    *n_points = cnt;
    return GD_OK;
}

}  // extern "C"
