// farneback.cu — Farnebäck dense optical flow for sm_100a, batched over independent RGB-D streams.
//
// Replaces cv::calcOpticalFlowFarneback(prvs, next, flow, 0.5, 3, 15, 3, 5, 1.2, 0) called by
// GeoMaskMaker::GetFlow (GD-SLAM src/GeoMaskMaker.cc:158-166).  OpenCV-4.13 semantics:
//   per image  : GaussianBlur(f32, ksize/sigma of the level, on the FULL-RES image) -> resize(INTER_LINEAR)
//                -> FarnebackPolyExp(n=5, sigma=1.2)  ==> R: 5 f32 planes per level           [k_fb_pyr, k_fb_polyexp]
//   per pair   : per level, coarse to fine: flow = 2*resize(prev level flow) (or 0), then 3 x
//                { M = UpdateMatrices(R0, R1, flow); flow = solve(box15x15(M)) }              [k_fb_upsample, k_fb_flow_iter]
// The per-image half is computed ONCE per frame and kept in the stream's device ring (the reference recomputes
// both pyramids on every call); M never touches HBM: each CTA recomputes it for its tile + 7-px halo in shared
// memory, runs the 15x15 box sums in FP64 (like OpenCV's double vsum) and solves the 2x2 system.
#include "farneback.cuh"

#include <cuda.h>  // CUtensorMap + the cuTensorMapEncodeTiled prototype (resolved at run time through the CUDA runtime, no libcuda link)

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <type_traits>

namespace gd {

// ------------------------------------------------------------------------------------------------ plan
static int cv_round_d(double v) { return (int)lrint(v); }

static void gaussian_taps(int n, double sigma, float* k)
{
    if (sigma <= 0 && n == 3) {
        k[0] = 0.25f; k[1] = 0.5f; k[2] = 0.25f;
        return;
    }
    const double sx = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double t[FB_MAX_KSIZE], sum = 0;
    for (int i = 0; i < n; ++i) {
        const double x = i - (n - 1) * 0.5;
        t[i] = std::exp(scale2x * x * x);
        sum += t[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < n; ++i) k[i] = (float)(t[i] * sum);
}

// FarnebackPrepareGaussian: taps g, x*g, x*x*g and the four entries of inv(G) that are used
static void poly_constants(int n, double sigma, FbPlan* p)
{
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    float* g = p->g + n;
    float* xg = p->xg + n;
    float* xxg = p->xxg + n;
    double s = 0.;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)std::exp(-x * x / (2 * sigma * sigma));
        s += g[x];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)(g[x] * s);
        xg[x] = (float)(x * g[x]);
        xxg[x] = (float)(x * x * g[x]);
    }
    double G[6][6] = {};
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            G[0][0] += g[y] * g[x];
            G[1][1] += g[y] * g[x] * x * x;
            G[3][3] += g[y] * g[x] * x * x * x * x;
            G[5][5] += g[y] * g[x] * x * x * y * y;
        }
    G[2][2] = G[0][3] = G[0][4] = G[3][0] = G[4][0] = G[1][1];
    G[4][4] = G[3][3];
    G[3][4] = G[4][3] = G[5][5];
    // G is block structured: {1, x^2, y^2} couple, x, y, xy are diagonal.  Invert the 3x3 block in closed form.
    const double a = G[0][0], b = G[0][3], c = G[3][3], d = G[3][4];
    // block [[a,b,b],[b,c,d],[b,d,c]]
    const double det = a * (c * c - d * d) - 2.0 * b * b * (c - d);
    p->ig03 = -b * (c - d) / det;            // inv(0,3)
    p->ig33 = (a * c - b * b) / det;         // inv(3,3)
    p->ig11 = 1.0 / G[1][1];
    p->ig55 = 1.0 / G[5][5];
}

int fb_prepare_device();

int fb_make_plan(int w, int h, double pyr_scale, int levels, int iterations, int poly_n, double poly_sigma, int winsize,
                 FbPlan* plan)
{
    GD_REQUIRE(poly_n == FB_POLY_N && winsize == FB_WIN, "only poly_n=5 / winsize=15 (the reference's parameters) are built");
    GD_REQUIRE(w >= 48 && h >= 48, "image too small (the pyramid blurs reflect at most once at a border)");
    GD_TRY(fb_prepare_device());
    *plan = FbPlan();
    plan->w = w;
    plan->h = h;
    plan->iterations = iterations;
    const int min_size = 32;
    int k;
    double scale = 1;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (w * scale < min_size || h * scale < min_size) break;
    }
    levels = k;
    GD_REQUIRE(levels + 1 <= FB_MAX_LEVELS, "too many pyramid levels");
    plan->nlevels = levels + 1;
    size_t r_off = 0, i_off = 0, f_off = 0;
    for (k = 0; k <= levels; ++k) {
        FbLevel& L = plan->lv[k];
        double sc = 1;
        for (int i = 0; i < k; i++) sc *= pyr_scale;
        const double sigma = (1. / sc - 1) * 0.5;
        int sz = cv_round_d(sigma * 5) | 1;
        L.ksize = sz > 3 ? sz : 3;
        GD_REQUIRE(L.ksize <= FB_MAX_KSIZE, "pyramid blur kernel too large");
        L.w = cv_round_d(w * sc);
        L.h = cv_round_d(h * sc);
        gaussian_taps(L.ksize, sigma, L.taps);
        L.scale_x = (double)w / L.w;
        L.scale_y = (double)h / L.h;
        L.r_off = r_off;
        L.i_off = i_off;
        L.f_off = f_off;
        const size_t n = align_up((size_t)L.w * L.h, 64);
        r_off += 5 * n;
        i_off += n;
        f_off += n;
    }
    plan->r_floats = r_off;
    plan->hrow_off = align_up(i_off, 64);
    {
        size_t mx = 0, sum = 0;  // the row-pass scratches of the resampled levels lie side by side (one launch for all levels)
        for (int q = 0; q < plan->nlevels; ++q) {
            mx = std::max(mx, (size_t)plan->lv[q].w);
            if (q > 0) sum += (size_t)plan->lv[q].w;
        }
        plan->i_floats = align_up(plan->hrow_off + 2 * (size_t)h * std::max(mx, sum), 64);  // + float2 row-pass scratch [h][widths]
    }
    plan->f_float2 = f_off;
    plan->m_floats = 5 * align_up((size_t)plan->lv[0].w * plan->lv[0].h, 64);
    poly_constants(poly_n, poly_sigma, plan);
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------ K1a-1
__device__ __forceinline__ int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

// BORDER_REFLECT_101 for -len < p < 2 len - 1 (blur radius smaller than the image: checked by the plan): no loop, no branch
__device__ __forceinline__ int reflect101_once(int p, int len)
{
    p = abs(p);
    return p >= len ? 2 * len - 2 - p : p;
}

// cv::resize source column of level column dx (and its interpolation weight)
__device__ __forceinline__ int fb_src_col(int dx, double scale_x, int W, float* frac)
{
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= W - 1) { fx = 0; sx = W - 1; }
    *frac = fx;
    return sx;
}

// column pass for an interior level row: rows sy - r .. sy + 1 + r of the row-pass scratch, both windows from one set of loads
template <int KS>
__device__ __forceinline__ void fb_colblur_window(const float2* __restrict__ hp, int lw, int sy, const float* taps, float& B00,
                                                  float& B01, float& B10, float& B11)
{
    constexpr int r = KS >> 1;
    float2 wv[KS + 1];
    const float2* p = hp + (size_t)(sy - r) * lw;
#pragma unroll
    for (int j = 0; j <= KS; ++j) wv[j] = __ldg(p + (size_t)j * lw);
    B00 = taps[r] * wv[r].x;
    B01 = taps[r] * wv[r].y;
    B10 = taps[r] * wv[r + 1].x;
    B11 = taps[r] * wv[r + 1].y;
#pragma unroll
    for (int i = 1; i <= r; ++i) {
        const float t = taps[r + i];
        B00 += t * (wv[r + i].x + wv[r - i].x);
        B01 += t * (wv[r + i].y + wv[r - i].y);
        B10 += t * (wv[r + 1 + i].x + wv[r + 1 - i].x);
        B11 += t * (wv[r + 1 + i].y + wv[r + 1 - i].y);
    }
}

// Blur + resample of all resampled levels in ONE launch per pass (the levels are independent: every level is blurred from
// the full-res image).  blockIdx.x walks the 32 x 8 tiles of all levels (tile_start), so the narrow levels do not cost a
// launch of their own — at small batches the pyramid was 7 dependent launches of a few microseconds of work.
// Pass A (k_fb_rowblur_levels): the row pass of the separable Gaussian on the FULL-RES rows, evaluated only at the two source
// columns (sx, sx + 1) every level column interpolates between; one thread = one (full-res row, level column); one code path
// for interior and border columns, the first tap peeled so that the loop body is two FMAs per load.
// Pass B (k_fb_colblur_resize_levels): column pass (symmetric form, REFLECT_101) at the two source rows of every level
// pixel, then cv::resize's bilinear combination (horizontal first, then vertical); interior rows load the ksize + 1 rows of
// the two overlapping windows once.
struct PyrLevelDev {
    int lw, lh, ksize, tiles_x;
    int tile_start_row, tile_start_col;  // first tile of this level in the row-pass / column-pass launch
    double scale_x, scale_y;
    unsigned long long hrow_off2, i_off;  // float2 offset of the level's row-pass scratch, float offset of its I plane
    float taps[FB_MAX_KSIZE];
};
struct PyrMultiArgs {
    int W, H, nl;
    PyrLevelDev lv[FB_MAX_LEVELS];
};

constexpr int RB_ROWS = 8;  // full-res rows per thread of the row pass: the per-column set-up (source column in f64, level
                            // look-up) is paid once for eight rows — it was two thirds of the instructions with one row per thread
__global__ void __launch_bounds__(256) k_fb_rowblur_levels(const uint8_t* __restrict__ gray, size_t gstride_b, PyrMultiArgs a,
                                                           float* __restrict__ scratch, size_t istride_b)
{
    pdl_wait();
    int l = 0;
    while (l + 1 < a.nl && (int)blockIdx.x >= a.lv[l + 1].tile_start_row) ++l;
    const PyrLevelDev& L = a.lv[l];
    // the level is only known at run time: the taps go through shared memory instead of indexed constant-bank reads
    __shared__ float taps[FB_MAX_KSIZE];
    {
        const int tl = threadIdx.y * 32 + threadIdx.x;
        if (tl < L.ksize) taps[tl] = L.taps[tl];
    }
    __syncthreads();
    const int lw = L.lw, ksize = L.ksize, W = a.W, H = a.H;
    const int t = (int)blockIdx.x - L.tile_start_row, ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
    const int dx = tx * 32 + threadIdx.x, ybase = (ty * 8 + threadIdx.y) * RB_ROWS, b = blockIdx.y;
    if (dx >= lw || ybase >= H) return;
    float fx;
    const int sx = fb_src_col(dx, L.scale_x, W, &fx);
    const int x0 = sx - (ksize >> 1);
    const uint8_t* g = gray + (size_t)b * gstride_b;
    int rowoff[RB_ROWS];  // rows past the image end read the last row (results discarded)
#pragma unroll
    for (int j = 0; j < RB_ROWS; ++j) rowoff[j] = min(ybase + j, H - 1) * W;
    // source columns x0 .. x0 + ksize ascending; column c feeds tap c of the window at sx and tap c - 1 of the one at sx + 1:
    // per accumulator the same products and sums in the same order as a per-row loop
    float acc0[RB_ROWS], acc1[RB_ROWS];
    {
        const int c0 = reflect101_once(x0, W), c1 = reflect101_once(x0 + 1, W);
        const float t0 = taps[0], t1 = taps[1];
#pragma unroll
        for (int j = 0; j < RB_ROWS; ++j) {
            const float v0 = (float)__ldg(g + rowoff[j] + c0), v1 = (float)__ldg(g + rowoff[j] + c1);
            acc0[j] = t0 * v0;
            acc0[j] = acc0[j] + t1 * v1;
            acc1[j] = t0 * v1;
        }
    }
    for (int c = 2; c < ksize; ++c) {
        const int col = reflect101_once(x0 + c, W);
        const float tA = taps[c], tB = taps[c - 1];
#pragma unroll
        for (int j = 0; j < RB_ROWS; ++j) {
            const float v = (float)__ldg(g + rowoff[j] + col);
            acc0[j] = acc0[j] + tA * v;
            acc1[j] = acc1[j] + tB * v;
        }
    }
    {
        const int col = reflect101_once(x0 + ksize, W);
        const float tB = taps[ksize - 1];
#pragma unroll
        for (int j = 0; j < RB_ROWS; ++j) acc1[j] = acc1[j] + tB * (float)__ldg(g + rowoff[j] + col);
    }
    float2* Hrow = reinterpret_cast<float2*>(scratch + (size_t)b * istride_b) + L.hrow_off2 + (size_t)ybase * lw + dx;
#pragma unroll
    for (int j = 0; j < RB_ROWS; ++j)
        if (ybase + j < H) Hrow[(size_t)j * lw] = make_float2(acc0[j], acc1[j]);
}

constexpr int CB_ROWS = 4;  // level rows per thread of the column pass (the per-column set-up is paid once)
__global__ void __launch_bounds__(256) k_fb_colblur_resize_levels(PyrMultiArgs a, float* __restrict__ scratch, size_t istride_b)
{
    pdl_wait();
    int l = 0;
    while (l + 1 < a.nl && (int)blockIdx.x >= a.lv[l + 1].tile_start_col) ++l;
    const PyrLevelDev& L = a.lv[l];
    __shared__ float taps[FB_MAX_KSIZE];
    {
        const int tl = threadIdx.y * 32 + threadIdx.x;
        if (tl < L.ksize) taps[tl] = L.taps[tl];
    }
    __syncthreads();
    const int lw = L.lw, lh = L.lh, ksize = L.ksize, H = a.H;
    const int t = (int)blockIdx.x - L.tile_start_col, ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
    const int dx = tx * 32 + threadIdx.x, dy0 = (ty * 8 + threadIdx.y) * CB_ROWS, b = blockIdx.y;
    if (dx >= lw || dy0 >= lh) return;
    float fx;
    (void)fb_src_col(dx, L.scale_x, a.W, &fx);
    const double scale_y = L.scale_y;
    const int r = ksize >> 1;
    float* sb = scratch + (size_t)b * istride_b;
    const float2* hp = reinterpret_cast<const float2*>(sb) + L.hrow_off2 + dx;
    float* out = sb + L.i_off + dx;
    const float a0 = 1.f - fx, a1 = fx;
#pragma unroll 1
    for (int j = 0; j < CB_ROWS; ++j) {
        const int dy = dy0 + j;
        if (dy >= lh) break;
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        int sy1 = sy + 1;
        sy = max(0, min(H - 1, sy));
        sy1 = max(0, min(H - 1, sy1));
        float B00, B01, B10, B11;
        // interior rows (all but the first / last few level rows): the two windows centred on sy and sy + 1 share ksize - 1 of
        // their rows; load the ksize + 1 rows once.  Same products and sums as the general path below.
        const bool interior = sy1 == sy + 1 && sy - r >= 0 && sy1 + r < H;
        if (interior && ksize == 19) {
            fb_colblur_window<19>(hp, lw, sy, taps, B00, B01, B10, B11);
        } else if (interior && ksize == 9) {
            fb_colblur_window<9>(hp, lw, sy, taps, B00, B01, B10, B11);
        } else if (interior && ksize == 3) {
            fb_colblur_window<3>(hp, lw, sy, taps, B00, B01, B10, B11);
        } else {
            const float2 c0 = __ldg(hp + (size_t)sy * lw);
            B00 = taps[r] * c0.x;
            B01 = taps[r] * c0.y;
#pragma unroll 4
            for (int i = 1; i <= r; ++i) {
                const float tp = taps[r + i];
                const float2 u = __ldg(hp + (size_t)reflect101(sy + i, H) * lw), d = __ldg(hp + (size_t)reflect101(sy - i, H) * lw);
                B00 += tp * (u.x + d.x);
                B01 += tp * (u.y + d.y);
            }
            B10 = B00;
            B11 = B01;
            if (sy1 != sy) {
                const float2 c1 = __ldg(hp + (size_t)sy1 * lw);
                B10 = taps[r] * c1.x;
                B11 = taps[r] * c1.y;
#pragma unroll 4
                for (int i = 1; i <= r; ++i) {
                    const float tp = taps[r + i];
                    const float2 u = __ldg(hp + (size_t)reflect101(sy1 + i, H) * lw), d = __ldg(hp + (size_t)reflect101(sy1 - i, H) * lw);
                    B10 += tp * (u.x + d.x);
                    B11 += tp * (u.y + d.y);
                }
            }
        }
        const float b0 = 1.f - fy, b1 = fy;
        const float r0 = B00 * a0 + B01 * a1;
        const float r1 = B10 * a0 + B11 * a1;
        out[(size_t)dy * lw] = r0 * b0 + r1 * b1;
    }
}

// level 0: the level has the size of the image, cv::resize is a copy -> one 3x3 separable blur per pixel
__global__ void __launch_bounds__(256) k_fb_blur3_same(const uint8_t* __restrict__ gray, size_t gstride_b, int W, int H, float t0,
                                                       float t1, float t2, float* __restrict__ I, size_t istride_b)
{
    pdl_wait();
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= W) return;
    const uint8_t* g = gray + (size_t)b * gstride_b;
    const int xm = reflect101_once(x - 1, W), xp = reflect101_once(x + 1, W);
    float h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint8_t* row = g + reflect101_once(y - 1 + j, H) * W;
        float acc = t0 * (float)__ldg(row + xm);
        acc = acc + t1 * (float)__ldg(row + x);
        acc = acc + t2 * (float)__ldg(row + xp);
        h[j] = acc;
    }
    I[(size_t)b * istride_b + (size_t)y * W + x] = t1 * h[1] + t2 * (h[2] + h[0]);
}

// the same for row strides that are a multiple of four: one thread = 4 consecutive columns x RY rows, marching down with the
// horizontal sums of three rows in registers (one word + two bytes loaded per row instead of 36 byte loads per 4 pixels);
// expressions and operation order as above, so the values are identical
template <int RY>
__global__ void __launch_bounds__(128) k_fb_blur3_same_v4(const uint8_t* __restrict__ gray, size_t gstride_b, int W, int H, float t0,
                                                          float t1, float t2, float* __restrict__ I, size_t istride_b)
{
    pdl_wait();
    const int x4 = (blockIdx.x * 32 + threadIdx.x) * 4, yb = (blockIdx.y * blockDim.y + threadIdx.y) * RY, b = blockIdx.z;
    if (x4 >= W || yb >= H) return;
    const uint8_t* g = gray + (size_t)b * gstride_b;
    const int xl = reflect101_once(x4 - 1, W), xr = reflect101_once(x4 + 4, W);
    auto hrow = [&](int y, float* h) {
        const uint8_t* row = g + (size_t)reflect101_once(y, H) * W;
        const unsigned wd = __ldg(reinterpret_cast<const unsigned*>(row + x4));
        const float p[6] = {(float)__ldg(row + xl), (float)(wd & 0xffu), (float)((wd >> 8) & 0xffu), (float)((wd >> 16) & 0xffu),
                            (float)(wd >> 24), (float)__ldg(row + xr)};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float acc = t0 * p[j];
            acc = acc + t1 * p[j + 1];
            acc = acc + t2 * p[j + 2];
            h[j] = acc;
        }
    };
    float h[3][4];
    hrow(yb - 1, h[0]);
    hrow(yb, h[1]);
    float* out = I + (size_t)b * istride_b + x4;
#pragma unroll
    for (int r = 0; r < RY; ++r) {
        const int y = yb + r;
        if (y >= H) break;
        hrow(y + 1, h[(r + 2) % 3]);
        const float* hm = h[r % 3];
        const float* hc = h[(r + 1) % 3];
        const float* hp = h[(r + 2) % 3];
        *reinterpret_cast<float4*>(out + (size_t)y * W) =
            make_float4(t1 * hc[0] + t2 * (hp[0] + hm[0]), t1 * hc[1] + t2 * (hp[1] + hm[1]), t1 * hc[2] + t2 * (hp[2] + hm[2]),
                        t1 * hc[3] + t2 * (hp[3] + hm[3]));
    }
}

// ------------------------------------------------------------------------------------------------ K1a-2
__device__ __forceinline__ size_t align_up_dev(size_t v, size_t a);
struct PolyLevelDev {
    int w, h, tiles_x, tile_start;
    unsigned long long i_off, r_off;  // float offsets of the level's I plane / R planes inside a stream's scratch / R pyramid
};
struct PolyArgs {
    int nl;
    PolyLevelDev lv[FB_MAX_LEVELS];
    float g[2 * FB_POLY_N + 1], xg[2 * FB_POLY_N + 1], xxg[2 * FB_POLY_N + 1];
    double gd[FB_POLY_N + 1], xxgd[FB_POLY_N + 1];  // (double)g[k], (double)xxg[k] for k >= 0: the two taps multiplied in f64
    double ig11, ig03, ig33, ig55;
};

constexpr int PT_W = 32, PT_H = 16, PT_TY = 8, PN = FB_POLY_N;  // 32x16 tile, 32x8 threads (2 rows per thread)
constexpr int PT_NT = PT_W * PT_TY;

// blockIdx.x walks the 32 x 16 tiles of ALL levels (independent of each other): one launch for the whole pyramid
__global__ void __launch_bounds__(PT_NT) k_fb_polyexp(const float* __restrict__ I, size_t istride_b, PolyArgs pa,
                                                            float* __restrict__ R, size_t rstride_b)
{
    pdl_wait();
    constexpr int TW = PT_W + 2 * PN;  // 42
    constexpr int TH = PT_H + 2 * PN;  // 26
    __shared__ float sI[TH][TW];
    __shared__ float4 sR[PT_H][TW];    // vertical sums (t0, t1, t2, -) of a column position: one vector load per tap
    int lvl = 0;
    while (lvl + 1 < pa.nl && (int)blockIdx.x >= pa.lv[lvl + 1].tile_start) ++lvl;
    const PolyLevelDev& LV = pa.lv[lvl];
    struct { int w, h; const float *g_, *xg_, *xxg_; const double *gd, *xxgd; double ig11, ig03, ig33, ig55; } a = {
        LV.w, LV.h, pa.g, pa.xg, pa.xxg, pa.gd, pa.xxgd, pa.ig11, pa.ig03, pa.ig33, pa.ig55};
    const int tile = (int)blockIdx.x - LV.tile_start, tyi = tile / LV.tiles_x, txi = tile - tyi * LV.tiles_x;
    const int b = blockIdx.y;
    const float* Ip = I + (size_t)b * istride_b + LV.i_off;
    const int x0 = txi * PT_W, y0 = tyi * PT_H;
    const int tid = threadIdx.y * PT_W + threadIdx.x;
    const float* g = a.g_ + PN;
    const float* xg = a.xg_ + PN;
    const float* xxg = a.xxg_ + PN;
    // tile + 5-px halo (replicate border).  Columns 0..31: warp = tile rows (stride 8), lane = column.  The 10 halo columns
    // 32..41 are a flat list (260 elements for the tile, 160 for the vertical pass) so that whole warps work on them.
    {
        const int lx = threadIdx.x;
        const int x = min(max(x0 + lx - PN, 0), a.w - 1);
#pragma unroll
        for (int ly = threadIdx.y; ly < TH; ly += PT_TY) {
            const int y = min(max(y0 + ly - PN, 0), a.h - 1);
            sI[ly][lx] = __ldg(Ip + (size_t)y * a.w + x);
        }
        for (int e = tid; e < 2 * PN * TH; e += PT_NT) {
            const int ly = (e * 6554) >> 16, hx = PT_W + e - 10 * ly;  // e / 10 for e < 16384
            static_assert(2 * PN == 10, "flat halo index split assumes 10 halo columns");
            const int xx = min(max(x0 + hx - PN, 0), a.w - 1), y = min(max(y0 + ly - PN, 0), a.h - 1);
            sI[ly][hx] = __ldg(Ip + (size_t)y * a.w + xx);
        }
    }
    __syncthreads();
    // vertical pass for PT_H rows x (PT_W + 10) columns
    auto vertical = [&](int ly, int lx) {
        const int cy = ly + PN;
        float t0 = sI[cy][lx] * g[0], t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int k = 1; k <= PN; ++k) {
            const float s0 = sI[cy - k][lx], s1 = sI[cy + k][lx];
            const float p = s0 + s1;
            t0 = t0 + g[k] * p;
            t1 = t1 + xg[k] * (s1 - s0);
            t2 = t2 + xxg[k] * p;
        }
        sR[ly][lx] = make_float4(t0, t1, t2, 0.f);
    };
#pragma unroll
    for (int ly = threadIdx.y; ly < PT_H; ly += PT_TY) vertical(ly, threadIdx.x);
    if (tid < 2 * PN * PT_H) {
        const int ly = (tid * 6554) >> 16;
        vertical(ly, PT_W + tid - 10 * ly);
    }
    __syncthreads();
    const int x = x0 + threadIdx.x;
    if (x >= a.w) return;
    const int lx = threadIdx.x + PN;
    const size_t npad = align_up_dev((size_t)a.w * a.h, 64);
    float* Rb = R + (size_t)b * rstride_b + LV.r_off;
#pragma unroll
    for (int rr = 0; rr < PT_H / PT_TY; ++rr) {
        const int ly = threadIdx.y + PT_TY * rr, y = y0 + ly;
        if (y >= a.h) break;
        const float4* r = sR[ly];
        const float4 c = r[lx];
        double b1 = c.x * g[0], b2 = 0, b3 = c.y * g[0], b4 = 0, b5 = c.z * g[0], b6 = 0;
#pragma unroll
        for (int k = 1; k <= PN; ++k) {
            const float4 p = r[lx + k], m = r[lx - k];
            const double tg = p.x + m.x;
            b1 += tg * a.gd[k];
            b4 += tg * a.xxgd[k];
            b2 += (p.x - m.x) * xg[k];
            b3 += (p.y + m.y) * g[k];
            b6 += (p.y - m.y) * xg[k];
            b5 += (p.z + m.z) * g[k];
        }
        // R layout per level: float4 plane (channels 0..3) followed by a float plane (channel 4) -> the flow kernels'
        // bilinear gather needs 2 loads per tap instead of 5
        const size_t o = (size_t)y * a.w + x;
        reinterpret_cast<float4*>(Rb)[o] = make_float4((float)(b3 * a.ig11), (float)(b2 * a.ig11), (float)(b1 * a.ig03 + b5 * a.ig33),
                                                       (float)(b1 * a.ig03 + b4 * a.ig33));
        Rb[4 * npad + o] = (float)(b6 * a.ig55);
    }
}

int fb_launch_pyramid_polyexp(const FbPlan& plan, const uint8_t* gray, size_t gray_stride_b, int batch, float* scratch_I,
                              size_t i_stride_b, float* R, size_t r_stride_b, cudaStream_t s, LaunchStats* st)
{
    // levels of the image size with a 3-tap blur (level 0): one 3x3 kernel; every other level: row pass + column pass /
    // resize, ALL of them in one launch each
    PyrMultiArgs ma;
    ma.W = plan.w; ma.H = plan.h; ma.nl = 0;
    int tiles_row = 0, tiles_col = 0, n_same = 0;
    size_t hrow2 = plan.hrow_off / 2;  // float2 units (hrow_off is a multiple of 64 floats)
    for (int k = 0; k < plan.nlevels; ++k) {
        const FbLevel& L = plan.lv[k];
        if (L.w == plan.w && L.h == plan.h && L.ksize == 3) {
            ++n_same;
            continue;
        }
        PyrLevelDev& D = ma.lv[ma.nl++];
        D.lw = L.w; D.lh = L.h; D.ksize = L.ksize; D.tiles_x = cdiv(L.w, 32);
        D.tile_start_row = tiles_row; D.tile_start_col = tiles_col;
        tiles_row += D.tiles_x * cdiv(plan.h, 8 * RB_ROWS);
        tiles_col += D.tiles_x * cdiv(L.h, 8 * CB_ROWS);
        D.scale_x = L.scale_x; D.scale_y = L.scale_y;
        D.hrow_off2 = hrow2;
        hrow2 += (size_t)plan.h * L.w;
        D.i_off = L.i_off;
        std::memcpy(D.taps, L.taps, sizeof(D.taps));
    }
    GD_REQUIRE(2 * hrow2 <= plan.i_floats, "row-pass scratch too small for the resampled levels");
    {
        LaunchScope ls(st, s, "K1a_blur_resample", n_same + (ma.nl ? 2 : 0));
        for (int k = 0; k < plan.nlevels; ++k) {
            const FbLevel& L = plan.lv[k];
            if (!(L.w == plan.w && L.h == plan.h && L.ksize == 3)) continue;
            float* Ik = scratch_I + L.i_off;
            const bool v4 = plan.w % 4 == 0 && plan.h >= 2 && gray_stride_b % 4 == 0 && i_stride_b % 4 == 0 &&
                            ((uintptr_t)gray & 3) == 0 && ((uintptr_t)Ik & 15) == 0;
            if (v4) {
                constexpr int RY = 8;
                dim3 b4(32, 4), g4(cdiv(plan.w / 4, 32), cdiv(plan.h, 4 * RY), batch);
                GD_CUDA(launch_pdl(k_fb_blur3_same_v4<RY>, g4, b4, 0, s, gray, gray_stride_b, plan.w, plan.h, L.taps[0], L.taps[1], L.taps[2], Ik, i_stride_b));
            } else {
                dim3 block(128), grid(cdiv(L.w, 128), L.h, batch);
                GD_CUDA(launch_pdl(k_fb_blur3_same, grid, block, 0, s, gray, gray_stride_b, plan.w, plan.h, L.taps[0], L.taps[1], L.taps[2], Ik, i_stride_b));
            }
            GD_CUDA(cudaGetLastError());
        }
        if (ma.nl) {
            GD_CUDA(launch_pdl(k_fb_rowblur_levels, dim3(tiles_row, batch), dim3(32, 8), 0, s, gray, gray_stride_b, ma, scratch_I, i_stride_b));
            GD_CUDA(cudaGetLastError());
            GD_CUDA(launch_pdl(k_fb_colblur_resize_levels, dim3(tiles_col, batch), dim3(32, 8), 0, s, ma, scratch_I, i_stride_b));
            GD_CUDA(cudaGetLastError());
        }
    }
    {
        LaunchScope ls(st, s, "K1a_polyexp", 1);
        PolyArgs po;
        po.nl = plan.nlevels;
        int tiles = 0;
        for (int k = 0; k < plan.nlevels; ++k) {
            const FbLevel& L = plan.lv[k];
            po.lv[k] = {L.w, L.h, cdiv(L.w, PT_W), tiles, (unsigned long long)L.i_off, (unsigned long long)L.r_off};
            tiles += cdiv(L.w, PT_W) * cdiv(L.h, PT_H);
        }
        std::memcpy(po.g, plan.g, sizeof(po.g));
        std::memcpy(po.xg, plan.xg, sizeof(po.xg));
        std::memcpy(po.xxg, plan.xxg, sizeof(po.xxg));
        po.ig11 = plan.ig11; po.ig03 = plan.ig03; po.ig33 = plan.ig33; po.ig55 = plan.ig55;
        for (int q = 0; q <= FB_POLY_N; ++q) {
            po.gd[q] = (double)plan.g[FB_POLY_N + q];
            po.xxgd[q] = (double)plan.xxg[FB_POLY_N + q];
        }
        GD_CUDA(launch_pdl(k_fb_polyexp, dim3(tiles, batch), dim3(PT_W, PT_TY), 0, s, scratch_I, i_stride_b, po, R, r_stride_b));
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------ K1b
// flow upsample: cv::resize(prevFlow, flow, INTER_LINEAR) then flow *= 1/pyr_scale
__global__ void __launch_bounds__(128) k_fb_upsample(const float2* __restrict__ src, int sw, int sh, float2* __restrict__ dst,
                                                     int dw, int dh, size_t fstride_b, float mul)
{
    const int dx = blockIdx.x * blockDim.x + threadIdx.x, dy = blockIdx.y, b = blockIdx.z;
    if (dx >= dw) return;
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    float fx = (float)((dx + 0.5) * scale_x - 0.5);
    int sx = (int)floorf(fx);
    fx -= sx;
    if (sx < 0) { fx = 0; sx = 0; }
    if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
    const int sx1 = min(sx + 1, sw - 1);
    float fy = (float)((dy + 0.5) * scale_y - 0.5);
    int sy = (int)floorf(fy);
    fy -= sy;
    int sy1 = sy + 1;
    sy = max(0, min(sh - 1, sy));
    sy1 = max(0, min(sh - 1, sy1));
    const float2* sp = src + (size_t)b * fstride_b;
    const float2 p00 = __ldg(sp + (size_t)sy * sw + sx), p01 = __ldg(sp + (size_t)sy * sw + sx1);
    const float2 p10 = __ldg(sp + (size_t)sy1 * sw + sx), p11 = __ldg(sp + (size_t)sy1 * sw + sx1);
    const float a0 = 1.f - fx, a1 = fx, b0 = 1.f - fy, b1 = fy;
    float2 o;
    o.x = ((p00.x * a0 + p01.x * a1) * b0 + (p10.x * a0 + p11.x * a1) * b1) * mul;
    o.y = ((p00.y * a0 + p01.y * a1) * b0 + (p10.y * a0 + p11.y * a1) * b1) * mul;
    dst[(size_t)b * fstride_b + (size_t)dy * dw + dx] = o;
}

constexpr int FT_W = 64, FT_H = 32, FHALO = FB_WIN / 2;            // output tile, 7-px halo
constexpr int FH_W = FT_W + 2 * FHALO, FH_H = FT_H + 2 * FHALO;    // 78 x 46 cells of M per tile
constexpr int FM_P = FH_W + 1;                                     // odd pitch: row-strided reads are conflict free
constexpr int FHS_P = FT_W + 1;                                    // pitch (doubles) of the horizontal-sum buffer
constexpr int FT_THREADS = 256;
constexpr int FT_SEG = 16;                                         // columns per horizontal running-sum task
constexpr int FT_ROWS_PER_THREAD = FT_H / (FT_THREADS / FT_W);     // 8 output rows per thread in the vertical pass
constexpr size_t FT_SMEM = sizeof(float) * 5 * FH_H * FM_P + sizeof(double) * FH_H * FHS_P;

__device__ __forceinline__ size_t align_up_dev(size_t v, size_t a) { return (v + a - 1) / a * a; }

// FarnebackUpdateMatrices for one pixel.  RA = float4 plane (channels 0..3), RB = float plane (channel 4).
__device__ __forceinline__ void fb_um_compute(const float4 a0, const float a04, const float4 p00, const float4 p01,
                                              const float4 p10, const float4 p11, const float e00, const float e01,
                                              const float e10, const float e11, const bool inside, const float fx, const float fy,
                                              const float2 fl, const int x, const int y, const int w, const int h, float M[5])
{
    const float dx = fl.x, dy = fl.y;
    float r2, r3, r4, r5, r6;
    if (inside) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        r2 = a00 * p00.x + a01 * p01.x + a10 * p10.x + a11 * p11.x;
        r3 = a00 * p00.y + a01 * p01.y + a10 * p10.y + a11 * p11.y;
        r4 = a00 * p00.z + a01 * p01.z + a10 * p10.z + a11 * p11.z;
        r5 = a00 * p00.w + a01 * p01.w + a10 * p10.w + a11 * p11.w;
        r6 = a00 * e00 + a01 * e01 + a10 * e10 + a11 * e11;
        r4 = (a0.z + r4) * 0.5f;
        r5 = (a0.w + r5) * 0.5f;
        r6 = (a04 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f;
        r4 = a0.z;
        r5 = a0.w;
        r6 = a04 * 0.5f;
    }
    r2 = (a0.x - r2) * 0.5f;
    r3 = (a0.y - r3) * 0.5f;
    r2 += r4 * dy + r6 * dx;
    r3 += r6 * dy + r5 * dx;
    constexpr int BORDER = 5;
    if ((unsigned)(x - BORDER) >= (unsigned)(w - BORDER * 2) || (unsigned)(y - BORDER) >= (unsigned)(h - BORDER * 2)) {
        const float border[BORDER] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        const float scale = (x < BORDER ? border[x] : 1.f) * (x >= w - BORDER ? border[w - x - 1] : 1.f) *
                            (y < BORDER ? border[y] : 1.f) * (y >= h - BORDER ? border[h - y - 1] : 1.f);
        r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
    }
    M[0] = r4 * r4 + r6 * r6;
    M[1] = (r4 + r5) * r6;
    M[2] = r5 * r5 + r6 * r6;
    M[3] = r4 * r2 + r6 * r3;
    M[4] = r6 * r2 + r5 * r3;
}

// One iteration of FarnebackUpdateFlow_Blur with the UpdateMatrices that precedes it fused in.
//   phase 1: M for the 64x32 tile + 7-px halo (clamped = replicate border) -> shared memory (f32, never in HBM)
//   phase 2: per channel, 15-wide horizontal window sums as FP64 RUNNING sums (one task = one halo row x 16 columns),
//            then 15-tall vertical running sums (one thread = one column x 8 rows) accumulated in registers
//   phase 3: 2x2 solve in FP64, coalesced float2 store
__global__ void __launch_bounds__(FT_THREADS, 2) k_fb_flow_iter(const float* __restrict__ R0, const float* __restrict__ R1,
                                                                size_t rstride_b, const float2* __restrict__ fin,
                                                                float2* __restrict__ fout, size_t fstride_b, int w, int h)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* sM = reinterpret_cast<float*>(smem_raw);                                     // [5][FH_H][FM_P]
    double* sH = reinterpret_cast<double*>(smem_raw + sizeof(float) * 5 * FH_H * FM_P);  // [FH_H][FHS_P]
    const int b = blockIdx.z;
    const size_t npad = align_up_dev((size_t)w * h, 64);
    const float* r0 = R0 + (size_t)b * rstride_b;
    const float* r1 = R1 + (size_t)b * rstride_b;
    const float4* R0A = reinterpret_cast<const float4*>(r0);
    const float4* R1A = reinterpret_cast<const float4*>(r1);
    const float* R0B = r0 + 4 * npad;
    const float* R1B = r1 + 4 * npad;
    const float2* fi = fin + (size_t)b * fstride_b;
    const int x0 = blockIdx.x * FT_W, y0 = blockIdx.y * FT_H;
    const int tid = threadIdx.x;
    // phase 1, software pipelined: U cells per thread in flight (flow loads, then all gathers, then the arithmetic) so
    // that the dependent global loads of several cells overlap (the kernel was long-scoreboard bound at 16 warps/SM)
    constexpr int U = 3;
    for (int i0 = tid; i0 < FH_H * FH_W; i0 += FT_THREADS * U) {
        int lxs[U], lys[U], xs[U], ys[U];
        float2 fl[U];
        bool live[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * FT_THREADS;
            live[u] = i < FH_H * FH_W;
            const int ii = live[u] ? i : 0;
            lys[u] = ii / FH_W;
            lxs[u] = ii - lys[u] * FH_W;
            xs[u] = min(max(x0 + lxs[u] - FHALO, 0), w - 1);
            ys[u] = min(max(y0 + lys[u] - FHALO, 0), h - 1);
            fl[u] = __ldg(fi + (size_t)ys[u] * w + xs[u]);
        }
        float4 a0[U], p00[U], p01[U], p10[U], p11[U];
        float a04[U], e00[U], e01[U], e10[U], e11[U], fxs[U], fys[U];
        bool inside[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const size_t o = (size_t)ys[u] * w + xs[u];
            a0[u] = __ldg(R0A + o);
            a04[u] = __ldg(R0B + o);
            float fx = xs[u] + fl[u].x, fy = ys[u] + fl[u].y;
            const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
            fxs[u] = fx - x1;
            fys[u] = fy - y1;
            inside[u] = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
            const size_t q = inside[u] ? (size_t)y1 * w + x1 : 0;
            p00[u] = __ldg(R1A + q);
            p01[u] = __ldg(R1A + q + 1);
            p10[u] = __ldg(R1A + q + w);
            p11[u] = __ldg(R1A + q + w + 1);
            e00[u] = __ldg(R1B + q);
            e01[u] = __ldg(R1B + q + 1);
            e10[u] = __ldg(R1B + q + w);
            e11[u] = __ldg(R1B + q + w + 1);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            float M[5];
            fb_um_compute(a0[u], a04[u], p00[u], p01[u], p10[u], p11[u], e00[u], e01[u], e10[u], e11[u], inside[u], fxs[u], fys[u],
                          fl[u], xs[u], ys[u], w, h, M);
            if (live[u]) {
#pragma unroll
                for (int c = 0; c < 5; ++c) sM[(c * FH_H + lys[u]) * FM_P + lxs[u]] = M[c];
            }
        }
    }
    __syncthreads();
    // horizontal task: seg-major so that the lanes of a warp walk different rows (odd pitch -> no bank conflicts)
    constexpr int NSEG = FT_W / FT_SEG;          // 4
    constexpr int NTASK = NSEG * FH_H;           // 184
    const int hseg = tid / FH_H, hrow = tid - hseg * FH_H;
    // vertical task: column cx, rows vr0 .. vr0+7
    const int cx = tid & (FT_W - 1), vr0 = (tid / FT_W) * FT_ROWS_PER_THREAD;
    double acc[5][FT_ROWS_PER_THREAD];
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        if (tid < NTASK) {
            const float* m = sM + (c * FH_H + hrow) * FM_P + hseg * FT_SEG;
            double* ho = sH + hrow * FHS_P + hseg * FT_SEG;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < FB_WIN; ++i) s += (double)m[i];
            ho[0] = s;
#pragma unroll
            for (int j = 1; j < FT_SEG; ++j) {
                s += (double)m[j + FB_WIN - 1] - (double)m[j - 1];
                ho[j] = s;
            }
        }
        __syncthreads();
        {
            const double* hp = sH + vr0 * FHS_P + cx;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < FB_WIN; ++i) s += hp[i * FHS_P];
            acc[c][0] = s;
#pragma unroll
            for (int j = 1; j < FT_ROWS_PER_THREAD; ++j) {
                s += hp[(j + FB_WIN - 1) * FHS_P] - hp[(j - 1) * FHS_P];
                acc[c][j] = s;
            }
        }
        __syncthreads();
    }
    const int x = x0 + cx;
    if (x >= w) return;
    float2* fo = fout + (size_t)b * fstride_b;
#pragma unroll
    for (int j = 0; j < FT_ROWS_PER_THREAD; ++j) {
        const int y = y0 + vr0 + j;
        if (y >= h) break;
        const double scale = 1. / (FB_WIN * FB_WIN);
        const double g11 = acc[0][j] * scale, g12 = acc[1][j] * scale, g22 = acc[2][j] * scale, h1 = acc[3][j] * scale,
                     h2 = acc[4][j] * scale;
        const double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
        float2 o;
        o.x = (float)((g11 * h2 - g12 * h1) * idet);
        o.y = (float)((g22 * h1 - g12 * h2) * idet);
        fo[(size_t)y * w + x] = o;
    }
}

// FarnebackUpdateMatrices of pixel (x, y) for the flow fl: loads of R0 at the pixel and the bilinear gather of R1 at the
// displaced position, then fb_um_compute.
__device__ __forceinline__ void fb_matrices_pixel(const float4* __restrict__ R0A, const float* __restrict__ R0B,
                                                  const float4* __restrict__ R1A, const float* __restrict__ R1B, int w, int h, int x,
                                                  int y, float2 fl, float M[5])
{
    const int o = y * w + x;
    const float4 a0 = __ldg(R0A + o);
    const float a04 = __ldg(R0B + o);
    float fx = x + fl.x, fy = y + fl.y;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= x1;
    fy -= y1;
    const bool inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    const int q = inside ? y1 * w + x1 : 0;
    const float4 p00 = __ldg(R1A + q), p01 = __ldg(R1A + q + 1), p10 = __ldg(R1A + q + w), p11 = __ldg(R1A + q + w + 1);
    const float e00 = __ldg(R1B + q), e01 = __ldg(R1B + q + 1), e10 = __ldg(R1B + q + w), e11 = __ldg(R1B + q + w + 1);
    fb_um_compute(a0, a04, p00, p01, p10, p11, e00, e01, e10, e11, inside, fx, fy, fl, x, y, w, h, M);
}

// the same in two steps, so that a thread can have the loads of several pixels in flight before it needs any of them
struct FbMatPix {
    float4 a0, p00, p01, p10, p11;
    float a04, e00, e01, e10, e11, fx, fy;
    bool inside;
};
__device__ __forceinline__ void fb_matrices_load(const float4* __restrict__ R0A, const float* __restrict__ R0B,
                                                 const float4* __restrict__ R1A, const float* __restrict__ R1B, int w, int h, int x, int y,
                                                 float2 fl, FbMatPix& m)
{
    const int o = y * w + x;
    m.a0 = __ldg(R0A + o);
    m.a04 = __ldg(R0B + o);
    float fx = x + fl.x, fy = y + fl.y;
    const int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    m.fx = fx - x1;
    m.fy = fy - y1;
    m.inside = (unsigned)x1 < (unsigned)(w - 1) && (unsigned)y1 < (unsigned)(h - 1);
    const int q = m.inside ? y1 * w + x1 : 0;
    m.p00 = __ldg(R1A + q); m.p01 = __ldg(R1A + q + 1); m.p10 = __ldg(R1A + q + w); m.p11 = __ldg(R1A + q + w + 1);
    m.e00 = __ldg(R1B + q); m.e01 = __ldg(R1B + q + 1); m.e10 = __ldg(R1B + q + w); m.e11 = __ldg(R1B + q + w + 1);
}

// ------------------------------------------------------------------------------------------------ K1b, split form
// (a) k_fb_matrices: FarnebackUpdateMatrices for every pixel exactly once -> M (5 f32 planes) in HBM.  Embarrassingly
//     parallel, no shared memory, full occupancy: the dependent bilinear gathers are hidden by ~40 resident warps.
//     Algorithmic bytes: R0 20 + R1 20 + flow 8 + M 20 = 68 B/px.
// (b) k_fb_box_solve: 15x15 box sums of M (FP64 running sums) + 2x2 solve -> flow.  28 B/px (M 20 + flow 8).
// MODE 0: flow of this level read from `fin`; MODE 1: first iteration of a level, flow = 2 * cv::resize(coarser flow)
// evaluated on the fly from `fin` (pw x ph), same arithmetic as k_fb_upsample; MODE 2: coarsest level, flow = 0.
template <int MODE>
__global__ void __launch_bounds__(256, 8) k_fb_matrices(const float* __restrict__ R0, const float* __restrict__ R1, size_t rstride_b,
                                                     const float2* __restrict__ fin, size_t fstride_b, int pw, int ph,
                                                     float* __restrict__ Mout, size_t mstride_b, int w, int h)
{
    pdl_trigger();
    pdl_wait();
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z;
    if (x >= w || y >= h) return;
    const size_t npad = align_up_dev((size_t)w * h, 64);
    const float* r0 = R0 + (size_t)b * rstride_b;
    const float* r1 = R1 + (size_t)b * rstride_b;
    const float4* R0A = reinterpret_cast<const float4*>(r0);
    const float4* R1A = reinterpret_cast<const float4*>(r1);
    const float* R0B = r0 + 4 * npad;
    const float* R1B = r1 + 4 * npad;
    const int o = y * w + x;
    float2 fl;
    if (MODE == 0) {
        fl = __ldg(fin + (size_t)b * fstride_b + o);
    } else if (MODE == 1) {
        const double scale_x = (double)pw / w, scale_y = (double)ph / h;
        float ux = (float)((x + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(ux);
        ux -= sx;
        if (sx < 0) { ux = 0; sx = 0; }
        if (sx >= pw - 1) { ux = 0; sx = pw - 1; }
        const int sx1 = min(sx + 1, pw - 1);
        float uy = (float)((y + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(uy);
        uy -= sy;
        int sy1 = sy + 1;
        sy = max(0, min(ph - 1, sy));
        sy1 = max(0, min(ph - 1, sy1));
        const float2* sp = fin + (size_t)b * fstride_b;
        const float2 q00 = __ldg(sp + sy * pw + sx), q01 = __ldg(sp + sy * pw + sx1);
        const float2 q10 = __ldg(sp + sy1 * pw + sx), q11 = __ldg(sp + sy1 * pw + sx1);
        const float c0 = 1.f - ux, c1 = ux, d0 = 1.f - uy, d1 = uy;
        fl.x = ((q00.x * c0 + q01.x * c1) * d0 + (q10.x * c0 + q11.x * c1) * d1) * 2.0f;
        fl.y = ((q00.y * c0 + q01.y * c1) * d0 + (q10.y * c0 + q11.y * c1) * d1) * 2.0f;
    } else {
        fl = make_float2(0.f, 0.f);
    }
    float M[5];
    fb_matrices_pixel(R0A, R0B, R1A, R1B, w, h, x, y, fl, M);
    float* mo = Mout + (size_t)b * mstride_b + o;
#pragma unroll
    for (int c = 0; c < 5; ++c) mo[(size_t)c * npad] = M[c];
}

constexpr int BX_W = 64, BX_H = 32;
constexpr int BXH_H = BX_H + 2 * FHALO;      // 46 tile rows
constexpr int BX_LEFT = 8;                   // tile starts at x0 - 8 (16-byte aligned), of which 7 columns are halo
constexpr int BXT_W = BX_W + 2 * BX_LEFT;    // 80 floats = 20 x 16-byte chunks per row
constexpr int BXM_P = 84;                    // smem row pitch (floats): 16-byte aligned rows, LDS.128 by row is conflict free
constexpr int BXS_P = BX_W + 1;              // 65 doubles
constexpr int BX_THREADS = 256;
constexpr int BX_ROWS = BX_H / (BX_THREADS / BX_W);  // 8 rows per thread in the vertical pass
constexpr int BX_TILE_BYTES = BXH_H * BXM_P * (int)sizeof(float);                    // 15 456: one channel tile (TMA box 84 x 46)
constexpr int BX_TILE_STRIDE = (BX_TILE_BYTES + 127) / 128 * 128;                    // 128-byte aligned tile buffers (TMA)
// shared memory of k_fb_box_solve<..., NBUF, F32>: NBUF channel-tile buffers, the horizontal-sum exchange buffer, mbarriers
constexpr int BXS_PF = BX_W + 4;             // 68 floats: pitch of the f32 horizontal-sum buffer (STS.128 by row conflict free)
template <int NBUF, bool F32>
constexpr size_t bx_smem_bytes()
{
    return (size_t)NBUF * BX_TILE_STRIDE + (F32 ? sizeof(float) * BXH_H * BXS_PF : sizeof(double) * BXH_H * BXS_P) + 8 * NBUF;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(sa), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

// ---- TMA (cp.async.bulk.tensor, SASS UTMALDG) + mbarrier helpers
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(bar);
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra.uni WAIT_DONE;\n"
        "bra.uni WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(a),
        "r"(parity)
        : "memory");
}
// one box of a rank-3 tensor (x, y, plane) -> shared memory; out-of-bounds elements are written as zeros
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y, int z, unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];\n" ::"r"(
            (unsigned)__cvta_generic_to_shared(smem_dst)),
        "l"(reinterpret_cast<unsigned long long>(tmap)), "r"(x), "r"(y), "r"(z), "r"((unsigned)__cvta_generic_to_shared(bar))
        : "memory");
}

// what the kernel does with the solved flow: NEXT = false -> store it (last iteration of a level); NEXT = true -> never store
// it: UpdateMatrices of the NEXT iteration only needs the flow of the pixel itself, so the same thread evaluates it right
// away and stores M for the following launch (R0 20 + R1 20 + M 20 bytes per pixel instead of flow 8 out + 8 in and a launch)
struct BoxNext {
    const float* R0;
    const float* R1;
    size_t rstride_b;
    float* Mout;  // same per-stream stride as Min
};

// sliding sums of 15 consecutive values without a running sum (no drift): pair, quad and octet partial sums, then
// out[j] = oct[j] + quad[j + 8] + pair[j + 12] + a[j + 14].  N outputs from N + 14 inputs.
template <int N>
__device__ __forceinline__ void window15_f32(const float* a, float* out)
{
    float p[N + 13], q[N + 11], o[N];
#pragma unroll
    for (int i = 0; i < N + 13; ++i) p[i] = a[i] + a[i + 1];
#pragma unroll
    for (int i = 0; i < N + 11; ++i) q[i] = p[i] + p[i + 2];
#pragma unroll
    for (int i = 0; i < N; ++i) o[i] = q[i] + q[i + 4];
#pragma unroll
    for (int j = 0; j < N; ++j) out[j] = ((o[j] + q[j + 8]) + p[j + 12]) + a[j + 14];
}

// VEC: the row stride is a multiple of 4 floats -> every 16-byte chunk is either fully inside or fully outside the image;
//      outside chunks are skipped and the replicated border columns are filled in shared memory afterwards (edge tiles only).
// TMA (needs VEC): the whole 84 x 46 channel tile is ONE cp.async.bulk.tensor issued by one thread, completion counted in
//      bytes on an mbarrier; the unit zero-fills outside the image, so edge tiles replicate the border rows as well.
// NEXT: see BoxNext.   NBUF: channel-tile buffers in flight (2 = double buffering, 5 = the whole CTA's input at once).
// F32:  true  = the 15 x 15 box sums in f32 as direct (tree) window sums, only the 2 x 2 solve in f64.  OpenCV accumulates
//               the box in f64; the difference stays below 12 % of the 1e-4 tolerance on the 640 x 480 pairs and
//               the kernel drops the f32->f64 conversions, half the exchange traffic and 40 registers;
//       false = f64 running sums like OpenCV's vsum / hsum (GD_FLOW_BOX_F64=1).
// MB:   resident CTAs per SM the register allocation aims at (2: 128 registers, 3: 85 — only the f32 form fits that).
template <bool VEC, bool TMA, bool NEXT, int NBUF, bool F32, int MB>
__global__ void __launch_bounds__(BX_THREADS, MB) k_fb_box_solve(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ Min,
                                                                size_t mstride_b, float2* __restrict__ fout, size_t fstride_b, int w,
                                                                int h, BoxNext nx)
{
    pdl_trigger();
    pdl_wait();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* sT = reinterpret_cast<float*>(smem_raw);                                   // [NBUF] channel tiles of [BXH_H][BXM_P]
    double* sH = reinterpret_cast<double*>(smem_raw + NBUF * BX_TILE_STRIDE);         // [BXH_H][BXS_P]   (f64 form)
    float* sHf = reinterpret_cast<float*>(smem_raw + NBUF * BX_TILE_STRIDE);          // [BXH_H][BXS_PF]  (f32 form)
    unsigned long long* s_bar = reinterpret_cast<unsigned long long*>(
        smem_raw + NBUF * BX_TILE_STRIDE + (F32 ? sizeof(float) * BXH_H * BXS_PF : sizeof(double) * BXH_H * BXS_P));
    constexpr int TS = BX_TILE_STRIDE / (int)sizeof(float);
    const int b = blockIdx.z;
    const size_t npad = align_up_dev((size_t)w * h, 64);
    const float* Mb = Min + (size_t)b * mstride_b;
    const int x0 = blockIdx.x * BX_W, y0 = blockIdx.y * BX_H;
    const int tid = threadIdx.x;
    // the chunk geometry is the same for the five channel planes: compute it once per thread (920 chunks / 256 threads)
    constexpr int CH = BXT_W / 4;                                       // 20 chunks per row
    constexpr int NCHUNK = (BXH_H * CH + BX_THREADS - 1) / BX_THREADS;  // 4 per thread
    int c_src[NCHUNK], c_dst[NCHUNK];  // element offsets in the plane / in the smem tile; c_src < 0: nothing to copy
    if (VEC && !TMA) {
#pragma unroll
        for (int k = 0; k < NCHUNK; ++k) {
            const int i = tid + k * BX_THREADS;
            const int ly = i / CH, ch = i - ly * CH;
            const int y = min(max(y0 + ly - FHALO, 0), h - 1);
            const int xs = x0 - BX_LEFT + 4 * ch;
            c_dst[k] = ly * BXM_P + 4 * ch;
            c_src[k] = (i < BXH_H * CH && xs >= 0 && xs + 3 < w) ? y * w + xs : -1;
        }
    }
    if (TMA) {
        if (tid == 0) {
#pragma unroll
            for (int c = 0; c < NBUF; ++c) mbar_init(&s_bar[c], 1);
            asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        }
        __syncthreads();
    }
    // tile reaches beyond the left / right (top / bottom: only the TMA path zero-fills rows, the others clamp the row index)
    const bool edge_l = x0 == 0, edge_r = x0 + BX_W + BX_LEFT > w;
    const bool edge_t = TMA && y0 == 0, edge_b = TMA && y0 + BX_H + FHALO > h;
    auto load_tile = [&](int c, int buf) {
        float* dst = sT + buf * TS;
        if (TMA) {
            if (tid == 0) {
                mbar_expect_tx(&s_bar[buf], (unsigned)BX_TILE_BYTES);
                tma_load_3d(dst, &tmap, x0 - BX_LEFT, y0 - FHALO, b * 5 + c, &s_bar[buf]);
            }
            return;
        }
        const float* plane = Mb + (size_t)c * npad;
        if (VEC) {
#pragma unroll
            for (int k = 0; k < NCHUNK; ++k)
                if (c_src[k] >= 0) cp_async16(dst + c_dst[k], plane + c_src[k]);
        } else {  // unaligned row stride (odd level widths): clamped scalar copies
            for (int i = tid; i < BXH_H * BXT_W; i += BX_THREADS) {
                const int ly = i / BXT_W, lx = i - ly * BXT_W;
                const int x = min(max(x0 - BX_LEFT + lx, 0), w - 1), y = min(max(y0 + ly - FHALO, 0), h - 1);
                cp_async4(dst + ly * BXM_P + lx, plane + (size_t)y * w + x);
            }
        }
        cp_async_commit();
    };
    // replicate the border pixel into the tile cells that lie outside the image (edge tiles only): columns first (on the
    // rows that exist), then whole rows
    auto fix_border = [&](int buf) {
        float* t = sT + buf * TS;
        if (edge_l)
            for (int i = tid; i < BXH_H * BX_LEFT; i += BX_THREADS) {
                const int ly = i / BX_LEFT, lx = i - ly * BX_LEFT;
                t[ly * BXM_P + lx] = t[ly * BXM_P + BX_LEFT];
            }
        if (edge_r) {
            const int last = w - 1 - (x0 - BX_LEFT);  // tile column of the last image pixel
            for (int i = tid; i < BXH_H * BXT_W; i += BX_THREADS) {
                const int ly = i / BXT_W, lx = i - ly * BXT_W;
                if (lx > last) t[ly * BXM_P + lx] = t[ly * BXM_P + last];
            }
        }
        if (edge_t || edge_b) {
            __syncthreads();
            const int first = edge_t ? FHALO : 0;                              // tile row of image row 0
            const int lastr = min(BXH_H - 1, h - 1 - (y0 - FHALO));            // tile row of the last image row
            for (int i = tid; i < BXH_H * BXT_W; i += BX_THREADS) {
                const int ly = i / BXT_W, lx = i - ly * BXT_W;
                if (ly < first) t[ly * BXM_P + lx] = t[first * BXM_P + lx];
                if (ly > lastr) t[ly * BXM_P + lx] = t[lastr * BXM_P + lx];
            }
        }
    };
    constexpr int NTASK = (BX_W / FT_SEG) * BXH_H;  // 184 horizontal tasks: (segment of 16 columns) x (tile row)
    const int hseg = tid / BXH_H, hrow = tid - hseg * BXH_H;
    const int cx = tid & (BX_W - 1), vr0 = (tid / BX_W) * BX_ROWS;
    typedef typename std::conditional<F32, float, double>::type acc_t;
    acc_t acc[5][BX_ROWS];
#pragma unroll
    for (int c = 0; c < NBUF && c < 5; ++c) load_tile(c, c);
#pragma unroll
    for (int c = 0; c < 5; ++c) {
        constexpr int dummy = 0;
        (void)dummy;
        const int buf = c % NBUF;
        if (TMA) {
            mbar_wait(&s_bar[buf], (unsigned)((c / NBUF) & 1));  // buffer `buf` is filled for the (c / NBUF)-th time
        } else {
            // groups complete in order; committed so far: min(5, NBUF + c); channel c has landed when at most
            // min(5, NBUF + c) - (c + 1) groups are still pending
            const int pending = (NBUF + c < 5 ? NBUF + c : 5) - (c + 1);
            switch (pending) {
                case 4: cp_async_wait<4>(); break;
                case 3: cp_async_wait<3>(); break;
                case 2: cp_async_wait<2>(); break;
                case 1: cp_async_wait<1>(); break;
                default: cp_async_wait<0>(); break;
            }
        }
        __syncthreads();  // tile c visible to everyone; the V pass of channel c - 1 is done with the exchange buffer
        if ((VEC || TMA) && (edge_l || edge_r || edge_t || edge_b)) {  // block-uniform
            fix_border(buf);
            __syncthreads();
        }
        if (tid < NTASK) {
            // window of output column j (tile-local) = tile floats [j + 1, j + 15]; the segment reads floats [16 seg, 16 seg + 32)
            const float4* m4 = reinterpret_cast<const float4*>(sT + buf * TS + hrow * BXM_P + hseg * FT_SEG);
            float m[32];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float4 v = m4[q];
                m[4 * q] = v.x; m[4 * q + 1] = v.y; m[4 * q + 2] = v.z; m[4 * q + 3] = v.w;
            }
            if (F32) {
                float hs[FT_SEG];
                window15_f32<FT_SEG>(m + 1, hs);
                float4* ho = reinterpret_cast<float4*>(sHf + hrow * BXS_PF + hseg * FT_SEG);
#pragma unroll
                for (int q = 0; q < FT_SEG / 4; ++q) ho[q] = make_float4(hs[4 * q], hs[4 * q + 1], hs[4 * q + 2], hs[4 * q + 3]);
            } else {
                double* ho = sH + hrow * BXS_P + hseg * FT_SEG;
                double s = 0.0;
#pragma unroll
                for (int i = 1; i <= FB_WIN; ++i) s += (double)m[i];
                ho[0] = s;
#pragma unroll
                for (int j = 1; j < FT_SEG; ++j) {
                    s += (double)m[j + FB_WIN] - (double)m[j];
                    ho[j] = s;
                }
            }
        }
        __syncthreads();
        if (c + NBUF < 5) load_tile(c + NBUF, buf);  // the H pass was the last reader of this tile buffer
        if (F32) {
            const float* hp = sHf + vr0 * BXS_PF + cx;
            float hv[BX_ROWS + FB_WIN - 1];
#pragma unroll
            for (int i = 0; i < BX_ROWS + FB_WIN - 1; ++i) hv[i] = hp[i * BXS_PF];
            float vs[BX_ROWS];
            window15_f32<BX_ROWS>(hv, vs);
#pragma unroll
            for (int j = 0; j < BX_ROWS; ++j) acc[c][j] = (acc_t)vs[j];
        } else {
            const double* hp = sH + vr0 * BXS_P + cx;
            double s = 0.0;
#pragma unroll
            for (int i = 0; i < FB_WIN; ++i) s += hp[i * BXS_P];
            acc[c][0] = (acc_t)s;
#pragma unroll
            for (int j = 1; j < BX_ROWS; ++j) {
                s += hp[(j + FB_WIN - 1) * BXS_P] - hp[(j - 1) * BXS_P];
                acc[c][j] = (acc_t)s;
            }
        }
        // the next iteration's first __syncthreads orders these exchange-buffer reads before the next H-pass writes
    }
    const int x = x0 + cx;
    if (x >= w) return;
    float2 o[BX_ROWS];
#pragma unroll
    for (int j = 0; j < BX_ROWS; ++j) {
        const double scale = 1. / (FB_WIN * FB_WIN);
        const double g11 = (double)acc[0][j] * scale, g12 = (double)acc[1][j] * scale, g22 = (double)acc[2][j] * scale,
                     h1 = (double)acc[3][j] * scale, h2 = (double)acc[4][j] * scale;
        const double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
        o[j].x = (float)((g11 * h2 - g12 * h1) * idet);
        o[j].y = (float)((g22 * h1 - g12 * h2) * idet);
    }
    if (!NEXT) {
        float2* fo = fout + (size_t)b * fstride_b;
#pragma unroll
        for (int j = 0; j < BX_ROWS; ++j) {
            const int y = y0 + vr0 + j;
            if (y < h) fo[(size_t)y * w + x] = o[j];
        }
    } else {
        const float* r0 = nx.R0 + (size_t)b * nx.rstride_b;
        const float* r1 = nx.R1 + (size_t)b * nx.rstride_b;
        const float4* R0A = reinterpret_cast<const float4*>(r0);
        const float4* R1A = reinterpret_cast<const float4*>(r1);
        const float* R0B = r0 + 4 * npad;
        const float* R1B = r1 + 4 * npad;
        float* mo = nx.Mout + (size_t)b * mstride_b;
        // several rows at a time: all gathers of the group are issued before the first one is consumed (the phase runs with
        // 16 warps per SM, so the memory parallelism has to come from inside the thread)
        constexpr int U = (F32 && MB == 2) ? 4 : 2;
#pragma unroll
        for (int j0 = 0; j0 < BX_ROWS; j0 += U) {
            FbMatPix mp[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = min(y0 + vr0 + j0 + u, h - 1);  // rows below the image: a valid address, result discarded
                fb_matrices_load(R0A, R0B, R1A, R1B, w, h, x, y, o[j0 + u], mp[u]);
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int y = y0 + vr0 + j0 + u;
                float M[5];
                fb_um_compute(mp[u].a0, mp[u].a04, mp[u].p00, mp[u].p01, mp[u].p10, mp[u].p11, mp[u].e00, mp[u].e01, mp[u].e10,
                              mp[u].e11, mp[u].inside, mp[u].fx, mp[u].fy, o[j0 + u], x, y, w, h, M);
                if (y < h) {
#pragma unroll
                    for (int c = 0; c < 5; ++c) mo[(size_t)c * npad + (size_t)y * w + x] = M[c];
                }
            }
        }
    }
}

// Function attributes are per device: called from fb_make_plan() with the plan's device current.
template <bool VEC, bool TMA, bool NEXT, int NBUF, bool F32, int MB>
static cudaError_t box_attr()
{
    return cudaFuncSetAttribute(k_fb_box_solve<VEC, TMA, NEXT, NBUF, F32, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)bx_smem_bytes<NBUF, F32>());
}
template <int NBUF, bool F32, int MB>
static cudaError_t box_attr_all()
{
    cudaError_t e;
    if ((e = box_attr<true, false, false, NBUF, F32, MB>()) != cudaSuccess) return e;
    if ((e = box_attr<true, false, true, NBUF, F32, MB>()) != cudaSuccess) return e;
    if ((e = box_attr<true, true, false, NBUF, F32, MB>()) != cudaSuccess) return e;
    if ((e = box_attr<true, true, true, NBUF, F32, MB>()) != cudaSuccess) return e;
    if ((e = box_attr<false, false, false, NBUF, F32, MB>()) != cudaSuccess) return e;
    return box_attr<false, false, true, NBUF, F32, MB>();
}

int fb_prepare_device()
{
    GD_CUDA(cudaFuncSetAttribute(k_fb_flow_iter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FT_SMEM));
    GD_CUDA((box_attr_all<2, true, 2>()));
    GD_CUDA((box_attr_all<5, true, 2>()));
    GD_CUDA((box_attr_all<2, true, 3>()));
    GD_CUDA((box_attr_all<5, true, 3>()));
    GD_CUDA((box_attr_all<2, false, 2>()));
    GD_CUDA((box_attr_all<5, false, 2>()));
    return GD_OK;
}

static int env_flag(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links the CUDA runtime only)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            p = nullptr;
        }
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

int fb_prepare_flow_buffers(const FbPlan& plan, int batch, float* M0, float* M1, size_t m_bytes_each, FbFlowBuffers* fb)
{
    *fb = FbFlowBuffers();
    fb->M[0] = M0;
    fb->M[1] = M1;
    fb->m_bytes = m_bytes_each;
    const bool whole_batch = m_bytes_each >= (size_t)batch * plan.m_floats * sizeof(float);
    fb->fuse_next = M1 != nullptr && whole_batch && env_flag("GD_FLOW_NEXT", 1) != 0;
    fb->box_f32 = env_flag("GD_FLOW_BOX_F64", 0) == 0;
    fb->nbuf = env_flag("GD_FLOW_NBUF", 2);
    fb->min_blocks = env_flag("GD_FLOW_MB", 3);
    fb->use_tma = false;
    // the TMA tile load is built and tested but OFF by default: measured 2-5 % slower than the cp.async chunks on B200 for this
    // tile (84-float box rows vs 80 needed, border fix-ups on all four sides, issue by a single thread) — profiles/README.md
    if (!whole_batch || env_flag("GD_FLOW_TMA", 0) == 0) return GD_OK;
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) return GD_OK;  // no driver entry point: the cp.async tile load is used
    for (int k = 0; k < plan.nlevels; ++k) {
        const FbLevel& L = plan.lv[k];
        if (L.w & 3) continue;  // rows not 16-byte aligned: that level takes the scalar path
        const size_t npad = align_up((size_t)L.w * L.h, 64);
        for (int i = 0; i < 2; ++i) {
            if (!fb->M[i]) continue;
            const cuuint64_t dims[3] = {(cuuint64_t)L.w, (cuuint64_t)L.h, (cuuint64_t)5 * batch};
            const cuuint64_t strides[2] = {(cuuint64_t)L.w * sizeof(float), (cuuint64_t)npad * sizeof(float)};
            const cuuint32_t box[3] = {(cuuint32_t)BXM_P, (cuuint32_t)BXH_H, 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            const CUresult r = enc(&fb->tmap[k][i], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, fb->M[i], dims, strides, box, estr,
                                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                set_error("cuTensorMapEncodeTiled failed (%d) for level %d (%d x %d)", (int)r, k, L.w, L.h);
                return GD_ECUDA;
            }
            fb->tmap_ok[k][i] = true;
        }
    }
    fb->use_tma = true;
    return GD_OK;
}

template <bool NEXT, int NBUF, bool F32, int MB>
static cudaError_t launch_box_v(const FbFlowBuffers& fb, int level, int src, dim3 grid, cudaStream_t s, const float* Min, size_t mstride,
                                float2* fout, size_t fstride, int w, int h, BoxNext nx)
{
    static const CUtensorMap no_map = {};
    const size_t smem = bx_smem_bytes<NBUF, F32>();
    if ((w & 3) != 0)
        return launch_pdl(k_fb_box_solve<false, false, NEXT, NBUF, F32, MB>, grid, dim3(BX_THREADS), smem, s, no_map, Min, mstride, fout, fstride, w, h, nx);
    if (fb.use_tma && fb.tmap_ok[level][src])
        return launch_pdl(k_fb_box_solve<true, true, NEXT, NBUF, F32, MB>, grid, dim3(BX_THREADS), smem, s, fb.tmap[level][src], Min, mstride, fout, fstride, w, h, nx);
    return launch_pdl(k_fb_box_solve<true, false, NEXT, NBUF, F32, MB>, grid, dim3(BX_THREADS), smem, s, no_map, Min, mstride, fout, fstride, w, h, nx);
}

template <bool NEXT>
static cudaError_t launch_box(const FbFlowBuffers& fb, int level, int src, dim3 grid, cudaStream_t s, const float* Min, size_t mstride,
                              float2* fout, size_t fstride, int w, int h, BoxNext nx)
{
    if (fb.box_f32 && fb.min_blocks >= 3) {
        if (fb.nbuf >= 5) return launch_box_v<NEXT, 5, true, 3>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
        return launch_box_v<NEXT, 2, true, 3>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
    }
    if (fb.box_f32) {
        if (fb.nbuf >= 5) return launch_box_v<NEXT, 5, true, 2>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
        return launch_box_v<NEXT, 2, true, 2>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
    }
    if (fb.nbuf >= 5) return launch_box_v<NEXT, 5, false, 2>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
    return launch_box_v<NEXT, 2, false, 2>(fb, level, src, grid, s, Min, mstride, fout, fstride, w, h, nx);
}

int fb_launch_flow(const FbPlan& plan, const float* R0, const float* R1, size_t r_stride_b, int batch, float2* flowA,
                   float2* flowB, size_t f_stride_b, const FbFlowBuffers* fbuf, const float2** final_flow, cudaStream_t s,
                   LaunchStats* st)
{
    const float2* prev = nullptr;
    int pw = 0, ph = 0;
    float* Mbuf = fbuf ? fbuf->M[0] : nullptr;
    const size_t m_bytes = fbuf ? fbuf->m_bytes : 0;
    for (int k = plan.nlevels - 1; k >= 0; --k) {
        const FbLevel& L = plan.lv[k];
        float2* A = flowA + L.f_off;
        float2* B = flowB + L.f_off;
        const size_t mstride = 5 * align_up((size_t)L.w * L.h, 64);
        if (fbuf && fbuf->fuse_next) {
            // M ping-pong: matrices(level entry) -> M[0]; box/solve + next matrices: M[0] -> M[1], M[1] -> M[0], ...;
            // the last iteration stores the flow of the level.  4 launches per level instead of 6, no intermediate flow.
            const float* r0 = R0 + L.r_off;
            const float* r1 = R1 + L.r_off;
            {
                LaunchScope ls(st, s, "K1b_matrices", 1);
                dim3 block(32, 8), grid(cdiv(L.w, 32), cdiv(L.h, 8), batch);
                if (prev)
                    GD_CUDA(launch_pdl(k_fb_matrices<1>, grid, block, 0, s, r0, r1, r_stride_b, prev, f_stride_b, pw, ph, fbuf->M[0], mstride, L.w, L.h));
                else
                    GD_CUDA(launch_pdl(k_fb_matrices<2>, grid, block, 0, s, r0, r1, r_stride_b, (const float2*)nullptr, f_stride_b, 0, 0, fbuf->M[0], mstride, L.w, L.h));
                GD_CUDA(cudaGetLastError());
            }
            dim3 grid(cdiv(L.w, BX_W), cdiv(L.h, BX_H), batch);
            int src = 0;
            for (int it = 0; it < plan.iterations; ++it) {
                const bool last = it + 1 == plan.iterations;
                BoxNext nx = {r0, r1, r_stride_b, fbuf->M[src ^ 1]};
                if (last) {
                    LaunchScope ls(st, s, "K1b_box_solve", 1);
                    GD_CUDA(launch_box<false>(*fbuf, k, src, grid, s, fbuf->M[src], mstride, A, f_stride_b, L.w, L.h, nx));
                } else {
                    LaunchScope ls(st, s, "K1b_box_matrices", 1);
                    GD_CUDA(launch_box<true>(*fbuf, k, src, grid, s, fbuf->M[src], mstride, A, f_stride_b, L.w, L.h, nx));
                }
                GD_CUDA(cudaGetLastError());
                src ^= 1;
            }
            prev = A;
            pw = L.w;
            ph = L.h;
            continue;
        }
        if (Mbuf) {
            // split form: the first matrices launch of a level takes its flow from the coarser level (or zero) on the fly
        } else if (!prev) {
            if (batch == 1)
                GD_CUDA(cudaMemsetAsync(A, 0, (size_t)L.w * L.h * sizeof(float2), s));
            else
                GD_CUDA(cudaMemset2DAsync(A, f_stride_b * sizeof(float2), 0, (size_t)L.w * L.h * sizeof(float2), batch, s));
        } else {
            LaunchScope ls(st, s, "K1b_flow_upsample", 1);
            dim3 block(128), grid(cdiv(L.w, 128), L.h, batch);
            k_fb_upsample<<<grid, block, 0, s>>>(prev, pw, ph, A, L.w, L.h, f_stride_b, 2.0f);
            GD_CUDA(cudaGetLastError());
        }
        float2* in = A;
        float2* out = B;
        for (int it = 0; it < plan.iterations; ++it) {
            if (Mbuf) {
                // split form: matrices once per pixel into M, then box filter + solve.  The streams of the batch go through
                // in groups whose M (packed at the level's own stride) stays inside the L2 window of m_bytes: written by
                // one kernel, read by the next, overwritten by the following group — M never has to reach HBM.
                const int group = (int)std::max<size_t>(1, std::min<size_t>((size_t)batch, m_bytes / (mstride * sizeof(float))));
                for (int b0 = 0; b0 < batch; b0 += group) {
                    const int nb = std::min(group, batch - b0);
                    const float* r0 = R0 + L.r_off + (size_t)b0 * r_stride_b;
                    const float* r1 = R1 + L.r_off + (size_t)b0 * r_stride_b;
                    {
                        LaunchScope ls(st, s, "K1b_matrices", 1);
                        dim3 block(32, 8), grid(cdiv(L.w, 32), cdiv(L.h, 8), nb);
                        if (it > 0)
                            GD_CUDA(launch_pdl(k_fb_matrices<0>, grid, block, 0, s, r0, r1, r_stride_b, (const float2*)(in + (size_t)b0 * f_stride_b), f_stride_b, 0, 0, Mbuf, mstride, L.w, L.h));
                        else if (prev)
                            GD_CUDA(launch_pdl(k_fb_matrices<1>, grid, block, 0, s, r0, r1, r_stride_b, prev + (size_t)b0 * f_stride_b, f_stride_b, pw, ph, Mbuf, mstride, L.w, L.h));
                        else
                            GD_CUDA(launch_pdl(k_fb_matrices<2>, grid, block, 0, s, r0, r1, r_stride_b, (const float2*)nullptr, f_stride_b, 0, 0, Mbuf, mstride, L.w, L.h));
                        GD_CUDA(cudaGetLastError());
                    }
                    LaunchScope ls(st, s, "K1b_box_solve", 1);
                    dim3 grid(cdiv(L.w, BX_W), cdiv(L.h, BX_H), nb);
                    // the tensor maps describe the whole batch from stream 0: the TMA path needs un-grouped launches
                    FbFlowBuffers one = *fbuf;
                    if (group < batch) one.use_tma = false;
                    GD_CUDA(launch_box<false>(one, k, 0, grid, s, (const float*)Mbuf, mstride, out + (size_t)b0 * f_stride_b, f_stride_b, L.w, L.h, BoxNext{}));
                    GD_CUDA(cudaGetLastError());
                }
            } else {
                LaunchScope ls(st, s, "K1b_flow_iter", 1);
                dim3 grid(cdiv(L.w, FT_W), cdiv(L.h, FT_H), batch);
                k_fb_flow_iter<<<grid, FT_THREADS, FT_SMEM, s>>>(R0 + L.r_off, R1 + L.r_off, r_stride_b, in, out, f_stride_b, L.w, L.h);
                GD_CUDA(cudaGetLastError());
            }
            float2* t = in;
            in = out;
            out = t;
        }
        prev = in;  // result of the last iteration
        pw = L.w;
        ph = L.h;
    }
    *final_flow = prev;
    return GD_OK;
}

}  // namespace gd
