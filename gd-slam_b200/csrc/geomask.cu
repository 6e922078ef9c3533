// geomask.cu — GeoMaskMaker kernels for sm_100a (compiled with --fmad=false: every f32/f64 operation below
// is the individually rounded operation the CPU oracle performs; the two fused operations OpenCV itself
// performs are explicit fmaf()).
//
// Reference being replaced (GD-SLAM tree):
//   K0  gray             src/GeoMaskMaker.cc:163-164 (BGR2GRAY), src/Tracking.cc:219-225 (RGB2GRAY on BGR bytes)
//   K2a depth edge       src/GeoMaskMaker.cc:854-964 (GetEdge)
//   K2b Mahalanobis      src/GeoMaskMaker.cc:208-272
//   K3  normalise/mask   src/GeoMaskMaker.cc:276-277, 405-407
//
// All of these are HBM-bound streaming kernels: one pass over the inputs, coalesced vector loads/stores,
// no tensor-core work (largest matrix on the path is 6x6).
#include "geomask.cuh"

#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace gd {

// ------------------------------------------------------------------------------------------------
// host helpers: OpenCV-semantics 3x3 inverse (f64 determinant/cofactors) and f32 3x3 gemm
// ------------------------------------------------------------------------------------------------
template <typename T>
__host__ __device__ inline bool inv3_cv(const T* m_, T* out)
{
    double m[9];
    for (int i = 0; i < 9; ++i) m[i] = (double)m_[i];
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0.0) {
        for (int i = 0; i < 9; ++i) out[i] = (T)0;
        return false;
    }
    d = 1.0 / d;
    out[0] = (T)((m[4] * m[8] - m[5] * m[7]) * d);
    out[1] = (T)((m[2] * m[7] - m[1] * m[8]) * d);
    out[2] = (T)((m[1] * m[5] - m[2] * m[4]) * d);
    out[3] = (T)((m[5] * m[6] - m[3] * m[8]) * d);
    out[4] = (T)((m[0] * m[8] - m[2] * m[6]) * d);
    out[5] = (T)((m[2] * m[3] - m[0] * m[5]) * d);
    out[6] = (T)((m[3] * m[7] - m[4] * m[6]) * d);
    out[7] = (T)((m[1] * m[6] - m[0] * m[7]) * d);
    out[8] = (T)((m[0] * m[4] - m[1] * m[3]) * d);
    return true;
}

// host-side strict f32 helpers (this TU is built with -ffp-contract=off on the host side as well)
static inline float h_dot3(float a0, float b0, float a1, float b1, float a2, float b2)
{
    volatile float t = a0 * b0;
    volatile float u = a1 * b1;
    t = t + u;
    u = a2 * b2;
    t = t + u;
    return t;
}

void make_cam_const(const float K[9], CamConst* c)
{
    c->fu = K[0];
    c->fv = K[4];
    c->cu = K[2];
    c->cv = K[5];
    c->rfu = 1.0f / c->fu;
    c->rfv = 1.0f / c->fv;
    {
        volatile float r = c->rfu * c->rfu;  // same individually rounded f32 products as depth2std()
        r = r * 0.5f;
        r = r * 0.5f;
        c->std_k = r;
    }
    inv3_cv<float>(K, c->Ki);
    double Kd[9];
    for (int i = 0; i < 9; ++i) Kd[i] = (double)K[i];
    inv3_cv<double>(Kd, c->Kid);
}

void make_undistort_args(const float K[9], const float* dist, int ndist, UndistortArgs* a)
{
    a->fx = (double)K[0];
    a->fy = (double)K[4];
    a->cx = (double)K[2];
    a->cy = (double)K[5];
    for (int i = 0; i < 5; ++i) a->k[i] = (dist && i < ndist) ? (double)dist[i] : 0.0;
}

KeyFormat make_key_format(size_t n_px)
{
    int idx_bits = 1;
    while (((size_t)1 << idx_bits) <= n_px + 1) ++idx_bits;  // source index + 1 fits
    KeyFormat kf;
    kf.shift = 32 + idx_bits;
    const int epoch_bits = 64 - kf.shift > 20 ? 20 : 64 - kf.shift;
    kf.epoch_max = (1 << epoch_bits) - 1;
    return kf;
}

void make_pose(const float K[9], const float R[9], const float T[3], int valid, int epoch, PoseDev* p)
{
    float Ki[9];
    inv3_cv<float>(K, Ki);
    for (int i = 0; i < 9; ++i) p->R[i] = R[i];
    for (int i = 0; i < 3; ++i) p->T[i] = T[i];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            p->RK[3 * i + j] = h_dot3(R[3 * i], Ki[j], R[3 * i + 1], Ki[3 + j], R[3 * i + 2], Ki[6 + j]);
    p->valid = valid;
    p->epoch = epoch;
    p->pad = 0;
}

// ------------------------------------------------------------------------------------------------
// K0 gray: g = (c0*k0 + c1*19235 + c2*k2 + 16384) >> 15
// one thread = 4 pixels = 12 input bytes (3 x u32 when aligned) -> one u32 store per output plane
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned gray_px(unsigned c0, unsigned c1, unsigned c2, unsigned k0, unsigned k2)
{
    return (c0 * k0 + c1 * 19235u + c2 * k2 + 16384u) >> 15;
}

__global__ void __launch_bounds__(256) k_gray(const uint8_t* __restrict__ bgr, size_t step, size_t stride_b, int w, int h,
                                              uint8_t* __restrict__ g_flow, size_t gstride_b, uint8_t* __restrict__ g_orb,
                                              int orb_order, size_t opitch, size_t ostride_b, int aligned)
{
    pdl_wait();
    const int b = blockIdx.z;
    const int groups = (w + 3) >> 2;
    const int gx = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y;
    if (gx >= groups) return;
    const uint8_t* row = bgr + (size_t)b * stride_b + (size_t)y * step;
    const int x0 = gx * 4;
    unsigned c[12];
    const int npx = min(4, w - x0);
    if (aligned && npx == 4) {
        const uint3 v = *reinterpret_cast<const uint3*>(row + (size_t)x0 * 3);
        const unsigned ww[3] = {v.x, v.y, v.z};
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = (ww[i >> 2] >> (8 * (i & 3))) & 0xffu;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) c[i] = (i < npx * 3) ? row[(size_t)x0 * 3 + i] : 0u;
    }
    unsigned of = 0, oo = 0;
    const unsigned ok0 = orb_order == 0 ? 3735u : 9798u, ok2 = orb_order == 0 ? 9798u : 3735u;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        of |= gray_px(c[3 * p], c[3 * p + 1], c[3 * p + 2], 3735u, 9798u) << (8 * p);
        oo |= gray_px(c[3 * p], c[3 * p + 1], c[3 * p + 2], ok0, ok2) << (8 * p);
    }
    const size_t o = (size_t)b * gstride_b + (size_t)y * w + x0;
    const size_t oo_off = (size_t)b * ostride_b + (size_t)y * opitch + x0;
    if (g_flow) {
        if (npx == 4 && (w & 3) == 0 && (gstride_b & 3) == 0)
            *reinterpret_cast<unsigned*>(g_flow + o) = of;
        else
            for (int p = 0; p < npx; ++p) g_flow[o + p] = (uint8_t)(of >> (8 * p));
    }
    if (g_orb) {
        if (npx == 4 && (opitch & 3) == 0 && (ostride_b & 3) == 0)
            *reinterpret_cast<unsigned*>(g_orb + oo_off) = oo;
        else
            for (int p = 0; p < npx; ++p) g_orb[oo_off + p] = (uint8_t)(oo >> (8 * p));
    }
}

int launch_gray(const uint8_t* bgr, size_t bgr_step, size_t bgr_stride_b, int w, int h, int batch, uint8_t* g_flow,
                size_t gray_stride_b, uint8_t* g_orb, int orb_order, size_t orb_pitch, size_t orb_stride_b, cudaStream_t s,
                LaunchStats* st)
{
    LaunchScope ls(st, s, "K0_gray", 1);
    const int aligned = ((reinterpret_cast<uintptr_t>(bgr) & 3) == 0 && (bgr_step & 3) == 0 && (bgr_stride_b & 3) == 0) ? 1 : 0;
    dim3 block(128), grid(cdiv((w + 3) / 4, 128), h, batch);
    GD_CUDA(launch_pdl(k_gray, grid, block, 0, s, bgr, bgr_step, bgr_stride_b, w, h, g_flow, gray_stride_b, g_orb, orb_order, orb_pitch, orb_stride_b, aligned));
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------
// K2a depth edge (GetEdge, GeoMaskMaker.cc:854-964).  The reference evaluates everything in FP64; its result per pixel is
// one bit (edge or not).  k_depth_edge evaluates the test in f32 INTERVAL form — every f32 quantity carries a rigorous bound
// on its distance from the reference's f64 value — and decides the pixel when the interval of thres_edge lies entirely on
// one side of 0.04.  The few undecided pixels (|thres - 0.04| within ~1e-4) are queued in shared memory and re-evaluated by
// edge_exact_f64(), the reference's f64 arithmetic operation by operation.  The output is bit-exact; the FP64 pipe is
// (nearly) idle.  k_depth_edge_f64 (the all-f64 kernel of round 1) stays as the path for cameras whose inv(K) has a
// non-trivial third row, and as the in-tree cross-check (GD_EDGE_F64=1).
//
// Notation of the error bounds stated at the kernel: u = 2^-24; S = max_i sum_j |Kinv_ij| * (W-1, H-1, 1) (>= 1); D = 3.5 m,
// the depth cut applied to interior pixels.
// ------------------------------------------------------------------------------------------------
constexpr int ET_W = 32, ET_H = 16;

// the reference's normal and vertex of tile pixel (ly, lx) [depth tile coordinates, halo 2] in f64, operation by operation
__device__ __forceinline__ void edge_nv_f64(const float (*sd)[ET_W + 4], int ly, int lx, int x, int y, int w, int h, const CamConst& cam,
                                            double n[3], double v[3])
{
    n[0] = n[1] = n[2] = 0.0;
    v[0] = v[1] = v[2] = 0.0;
    if (x >= 1 && y >= 1 && x < w - 1 && y < h - 1) {
        const double dc = (double)sd[ly][lx], dt = (double)sd[ly - 1][lx], dl = (double)sd[ly][lx - 1];
        if (dc != 0.0 && dt != 0.0 && dl != 0.0) {
            // cross((-1, 0, dl - dc), (0, -1, dt - dc)) = (dl - dc, dt - dc, 1): the products with the constant
            // components are exact and the differences are never -0, so the three components need no arithmetic
            const double c0 = dl - dc, c1 = dt - dc, c2 = 1.0;
            double ss = c0 * c0;  // cv::norm: 0 + c0^2 + c1^2 + c2^2 in this order (0 + x == x for x >= 0)
            ss += c1 * c1;
            ss += c2 * c2;
            const double nv = sqrt(ss);
            const double inv = nv != 0.0 ? 1.0 / nv : 0.0;
            n[0] = c0 * inv;
            n[1] = c1 * inv;
            n[2] = c2 * inv;
            const double px = (double)x, py = (double)y;
            const double h0 = cam.Kid[0] * px + cam.Kid[1] * py + cam.Kid[2] * 1.0;
            const double h1 = cam.Kid[3] * px + cam.Kid[4] * py + cam.Kid[5] * 1.0;
            const double h2 = cam.Kid[6] * px + cam.Kid[7] * py + cam.Kid[8] * 1.0;
            v[0] = h0 * dc;
            v[1] = h1 * dc;
            v[2] = h2 * dc;
        }
    }
}

// the reference's edge decision of an interior pixel with non-zero (clamped) depth, from normals / vertices given by nv(k, n, v)
// for k = 0..7 (neighbour k) and k = 8 (the pixel itself)
template <class NV>
__device__ __forceinline__ uint8_t edge_decide_f64(NV nv)
{
    double cn[3], cv[3];
    nv(8, cn, cv);
    bool zero_nb = false;
    double max_phi_d = -1.0, max_phi_c = -1.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        double jn[3], jv[3];
        nv(k, jn, jv);
        const double z = jv[2];
        if (z == 0.0) {
            zero_nb = true;
            continue;
        }
        // the leading "0 +" of the dot products is dropped: it can only turn a -0 sum into +0, which no later
        // comparison distinguishes
        double phi_d = (jv[0] - cv[0]) * cn[0];
        phi_d += (jv[1] - cv[1]) * cn[1];
        phi_d += (z - cv[2]) * cn[2];
        const double ad = fabs(phi_d);
        if (max_phi_d < ad) max_phi_d = ad;
        if (phi_d < 0.0) {
            if (max_phi_c < 0.0) max_phi_c = 0.0;
        } else {
            double dot = jn[0] * cn[0];
            dot += jn[1] * cn[1];
            dot += jn[2] * cn[2];
            const double phi_c = 1.0 - dot;
            if (phi_c > max_phi_c) max_phi_c = phi_c;
        }
    }
    if (zero_nb) return 255;
    if (!(max_phi_c == -1.0 || max_phi_d == -1.0)) {
        const double thres_edge = max_phi_d + 0.05 * max_phi_c;
        if (thres_edge > 0.04) return 255;
    }
    return 0;
}

// neighbour order of the reference (:905-906); compile-time functions so that the unrolled loops fold them into immediates
__host__ __device__ constexpr int edge_nx(int k) { return k == 0 || k == 1 || k == 7 ? -1 : (k == 2 || k == 6 ? 0 : (k < 8 ? 1 : 0)); }
__host__ __device__ constexpr int edge_ny(int k) { return k == 1 || k == 2 || k == 3 ? -1 : (k == 0 || k == 4 ? 0 : (k < 8 ? 1 : 0)); }

// depth tile with halo 2, interior pixels clamped like :870-874.  Thread (tx, ty) of the 32 x 16 block loads rows ty, ty + 16 and
// columns tx, tx + 32 (no index divisions).  Returns true when a loaded value is NaN or infinite (the f32 interval form is
// only used on finite depths).
__device__ __forceinline__ bool load_depth_tile(const float* __restrict__ dp, int w, int h, int x0, int y0, int tx, int ty,
                                                float (*sd)[ET_W + 4])
{
    bool bad = false;
#pragma unroll
    for (int ry = 0; ry < 2; ++ry) {
        const int ly = ty + ry * ET_H;
        if (ly >= ET_H + 4) break;
        const int y = y0 + ly - 2;
#pragma unroll
        for (int rx = 0; rx < 2; ++rx) {
            const int lx = tx + rx * ET_W;
            if (lx >= ET_W + 4) break;
            const int x = x0 + lx - 2;
            float v = 0.f;
            if (x >= 0 && y >= 0 && x < w && y < h) {
                v = __ldg(dp + (size_t)y * w + x);
                const bool interior = x >= 1 && y >= 1 && x < w - 1 && y < h - 1;
                if (interior && v > 3.5f) v = 0.f;  // :870-874 ((double)v > 3.5 <=> v > 3.5f: 3.5 is a float)
                bad = bad || !(fabsf(v) <= 3.0e38f);
            }
            sd[ly][lx] = v;
        }
    }
    return bad;
}

struct EdgeBounds {
    float Ki[6];  // first two rows of inv(K) rounded to f32 (the third row is (0, 0, 1) on this path)
    float e_d;    // bound on |phi_d_f32 - phi_d_f64|
    float e_c;    // bound on |phi_c_f32 - phi_c_f64|
};

// f32 interval form.  Shared memory holds ONE float4 per pixel, A = (n0, n1, n2, depth); the vertex is never stored:
// with h(x, y) = inv(K) (x, y, 1) and v = h * depth,
//     phi_d = (v_j - v_c) . n_c = d_j * G_j - d_c * g,     g = h_c . n_c,   G_j = g + dx * gx + dy * gy,
//     gx = Ki0 n0 + Ki3 n1,  gy = Ki1 n0 + Ki4 n1          (h_j - h_c = dx * column 0 + dy * column 1 of inv(K)),
// so a neighbour costs one vector load, one fused multiply-add for phi_d, three for phi_c and the interval bookkeeping.
// n2 = 1 / sqrt(ss) lies in [0.19, 1], so n2 == 0 marks "no normal / zero vertex" (border pixels, pixels whose own, top or
// left depth is missing).  Error bounds (u = 2^-24, S >= 1 and D = 3.5 as above; the reference's own f64 rounding is below
// 1e-15 and covered by the margins):
//   n_f: c = dl - dc rounds once (u), ss = fma(c0, c0, fma(c1, c1, 1)) is within 4 u, MUFU.RSQ within 2 ulp = 4 u + 2 u from
//        ss, the product one more: relative error <= 10 u                                             -> en = 12 u
//   h_f: |h_f - h| <= 6 u S;   g_f: 2 (6 + 12) u S + 12 u inputs + 3 roundings of <= 3 u S            -> 57 u S
//   gx_f, gy_f: <= 15 u (|Ki0| + |Ki3|) <= 30 u S each;  G_j: g, gx, gy and two roundings             -> 123 u S
//   phi_d_f = fma(d_j, G_j, -d_c g): D 123 u S + D 57 u S + 3 u S D + 7 u S D = 190 u S D             -> E_D = 256 u S D
//   phi_c_f = 1 - sum_i n_j,i n_c,i: 12 u sum_i (|n_c,i| + |n_j,i|) <= 42 u, + 4 u of roundings       -> E_C = 64 u
//
// Execution: the kernel is latency bound when every tile is its own CTA (ncu: barrier + long-scoreboard stalls, the depth
// images come from DRAM), so it is PERSISTENT — 2 CTAs per SM walk the tile list, and the depth values of tile i+1 are
// requested into registers right after tile i's values have been published to shared memory; they arrive while the normals
// and the neighbour loop of tile i run.  The depth tile is double buffered (the exact f64 pass of stragglers may still read
// the old one), the per-tile flags alternate between two slots, three barriers per tile.
struct EdgeTiles {
    int tiles_x, tiles_per_img, total;
    unsigned magic_x, magic_img;  // ceil(2^32 / d): __umulhi(t, magic) == t / d for t * d < 2^32
};

// raw depth (0 outside the image); the clamp of :870-874 is applied when the value is published to shared memory, so that
// nothing depends on the load while the previous tile is being evaluated
__device__ __forceinline__ float edge_load_raw(const float* __restrict__ dp, int x, int y, int w, int h)
{
    float v = 0.f;
    if (x >= 0 && y >= 0 && x < w && y < h) v = __ldg(dp + (size_t)y * w + x);
    return v;
}
// `bad`: the value leaves the range the f32 error bounds assume — interior pixels |d| <= D = 3.5 after the clamp (a NaN, an
// infinity or a large negative depth fails), border pixels (never clamped, they only enter the normals of their neighbours)
// |d| <= 1e4 so that no f32 intermediate overflows.  A tile with a bad value is evaluated by the f64 arithmetic throughout.
__device__ __forceinline__ float edge_clamp(float v, int x, int y, int w, int h, bool& bad)
{
    const bool interior = x >= 1 && y >= 1 && x < w - 1 && y < h - 1;
    v = interior && v > 3.5f ? 0.f : v;  // (double)v > 3.5 <=> v > 3.5f: 3.5 is a float
    bad = bad || !(fabsf(v) <= (interior ? 3.5f : 1.0e4f));
    return v;
}

__global__ void __launch_bounds__(ET_W* ET_H, 2) k_depth_edge(const float* __restrict__ depth, size_t dstride_b, int w, int h,
                                                               CamConst cam, EdgeBounds eb, EdgeTiles et, uint8_t* __restrict__ edge,
                                                               size_t estride_b)
{
    pdl_wait();
    __shared__ float sd2[2][ET_H + 4][ET_W + 4];   // clamped depth, halo 2 (double buffered)
    __shared__ float4 sA[ET_H + 2][ET_W + 2];      // (normal, clamped depth), halo 1
    __shared__ int s_todo[ET_W * ET_H];
    __shared__ int s_flags[2][2];                  // per tile parity: {number of undecided pixels, non-finite depth seen}
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int tid = ty * ET_W + tx;
    if (tid < 4) (&s_flags[0][0])[tid] = 0;
    // Besides its own element every thread of the first warps owns one element of the halo strips (no per-element divisions
    // inside the tile loop): depth tile = rows 16..19 (4 x 36) + columns 32..35 of rows 0..15 (16 x 4) = 208 elements,
    // normal tile = rows 16..17 (2 x 34) + columns 32..33 of rows 0..15 (16 x 2) = 100 elements.
    int dly = -1, dlx = 0, nly = -1, nlx = 0;
    if (tid < 4 * (ET_W + 4)) {
        dly = ET_H + tid / (ET_W + 4);
        dlx = tid % (ET_W + 4);
    } else if (tid < 4 * (ET_W + 4) + ET_H * 4) {
        const int k = tid - 4 * (ET_W + 4);
        dly = k >> 2;
        dlx = ET_W + (k & 3);
    }
    if (tid < 2 * (ET_W + 2)) {
        nly = ET_H + tid / (ET_W + 2);
        nlx = tid % (ET_W + 2);
    } else if (tid < 2 * (ET_W + 2) + ET_H * 2) {
        const int k = tid - 2 * (ET_W + 2);
        nly = k >> 1;
        nlx = ET_W + (k & 1);
    }
    int tile = blockIdx.x;
    int b = 0, x0 = 0, y0 = 0;
    auto decode = [&](int t, int& tb, int& tx0, int& ty0) {
        tb = (int)__umulhi((unsigned)t, et.magic_img);
        const int r = t - tb * et.tiles_per_img;
        const int tyi = (int)__umulhi((unsigned)r, et.magic_x);
        tx0 = (r - tyi * et.tiles_x) * ET_W;
        ty0 = tyi * ET_H;
    };
    float v0 = 0.f, v1 = 0.f;
    if (tile < et.total) {
        decode(tile, b, x0, y0);
        const float* dp = depth + (size_t)b * dstride_b;
        v0 = edge_load_raw(dp, x0 + tx - 2, y0 + ty - 2, w, h);
        if (dly >= 0) v1 = edge_load_raw(dp, x0 + dlx - 2, y0 + dly - 2, w, h);
    }
    __syncthreads();  // flags zeroed
    for (int it = 0; tile < et.total; ++it, tile += gridDim.x) {
        const int par = it & 1;
        float (*sd)[ET_W + 4] = sd2[par];
        bool bad = false;
        v0 = edge_clamp(v0, x0 + tx - 2, y0 + ty - 2, w, h, bad);
        sd[ty][tx] = v0;
        if (dly >= 0) {
            v1 = edge_clamp(v1, x0 + dlx - 2, y0 + dly - 2, w, h, bad);
            sd[dly][dlx] = v1;
        }
        if (bad) s_flags[par][1] = 1;
        __syncthreads();
        const int cb = b, cx0 = x0, cy0 = y0;  // the tile being evaluated
        {
            const int nt = tile + gridDim.x;
            if (tid < 2) s_flags[par ^ 1][tid] = 0;  // last read before the barrier above, next written after the third one
            if (nt < et.total) {
                decode(nt, b, x0, y0);
                const float* dp = depth + (size_t)b * dstride_b;
                v0 = edge_load_raw(dp, x0 + tx - 2, y0 + ty - 2, w, h);
                if (dly >= 0) v1 = edge_load_raw(dp, x0 + dlx - 2, y0 + dly - 2, w, h);
            }
        }
        auto normal_at = [&](int ly, int lx) {  // normal tile element (halo 1) from the depth tile (halo 2)
            const int x = cx0 + lx - 1, y = cy0 + ly - 1;
            float4 A = make_float4(0.f, 0.f, 0.f, 0.f);
            if (x >= 1 && y >= 1 && x < w - 1 && y < h - 1) {
                const float dc = sd[ly + 1][lx + 1], dt = sd[ly][lx + 1], dl = sd[ly + 1][lx];
                A.w = dc;
                if (dc != 0.f && dt != 0.f && dl != 0.f) {
                    const float c0 = dl - dc, c1 = dt - dc;
                    const float ss = fmaf(c0, c0, fmaf(c1, c1, 1.0f));
                    float inv;
                    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ss));
                    A.x = c0 * inv;
                    A.y = c1 * inv;
                    A.z = inv;
                }
            }
            sA[ly][lx] = A;
        };
        normal_at(ty, tx);
        if (nly >= 0) normal_at(nly, nlx);
        __syncthreads();
        const int x = cx0 + tx, y = cy0 + ty;
        const bool in_img = x < w && y < h;
        uint8_t e = 0;
        const int lx = tx + 1, ly = ty + 1;
        const float4 Ac = sA[ly][lx];
        if (in_img && x >= 1 && y >= 1 && x < w - 1 && y < h - 1 && Ac.w != 0.f) {
            if (s_flags[par][1]) {
                s_todo[atomicAdd(&s_flags[par][0], 1)] = (ly << 8) | lx;  // tile with a value outside the bounds' range: all exact
            } else if (Ac.z == 0.f) {
                // a pixel with depth but without a normal (its top or left neighbour has no depth): cn = cv = 0 -> a zero vertex
                // among the neighbours, or phi_d = 0 and phi_c = 1 for every (finite) neighbour -> 0.05 > 0.04: edge in both cases
                e = 255;
            } else {
                const float px = (float)x, py = (float)y;
                const float h0 = fmaf(eb.Ki[0], px, fmaf(eb.Ki[1], py, eb.Ki[2]));
                const float h1 = fmaf(eb.Ki[3], px, fmaf(eb.Ki[4], py, eb.Ki[5]));
                const float g = fmaf(h0, Ac.x, fmaf(h1, Ac.y, Ac.z));
                const float gx = fmaf(eb.Ki[0], Ac.x, eb.Ki[3] * Ac.y);
                const float gy = fmaf(eb.Ki[1], Ac.x, eb.Ki[4] * Ac.y);
                const float ncw = -(Ac.w * g);
                const float gxs[3] = {g - gx, g, g + gx};
                float minz = 1.f, maxd = 0.f, c_lo = 0.f, c_hi = 0.f;  // eight valid neighbours -> both maxima are >= 0 in the reference
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int dx = edge_nx(k), dy = edge_ny(k);
                    const float4 Aj = sA[ly + dy][lx + dx];
                    const float G = dy == 0 ? gxs[dx + 1] : (dy < 0 ? gxs[dx + 1] - gy : gxs[dx + 1] + gy);
                    minz = fminf(minz, Aj.z);
                    const float phi_d = fmaf(Aj.w, G, ncw);
                    const float phi_c = fmaf(-Aj.x, Ac.x, fmaf(-Aj.y, Ac.y, fmaf(-Aj.z, Ac.z, 1.0f)));
                    maxd = fmaxf(maxd, fabsf(phi_d));
                    // contribution of this neighbour to max_phi_c: phi_c when phi_d >= 0, 0 when phi_d < 0; undecided sign -> both
                    if (phi_d >= -eb.e_d) c_hi = fmaxf(c_hi, phi_c);
                    if (phi_d >= eb.e_d) c_lo = fmaxf(c_lo, phi_c);
                }
                if (minz == 0.f) {
                    e = 255;  // a zero vertex among the neighbours (:925-949)
                } else {
                    const float slack = eb.e_d * 1.125f + 1e-6f;  // + rounding of the few f32 operations below
                    const float t_lo = (maxd - slack) + 0.05f * (c_lo - eb.e_c);
                    const float t_hi = (maxd + slack) + 0.05f * (c_hi + eb.e_c);
                    if (t_lo > 0.04f)
                        e = 255;
                    else if (!(t_hi < 0.04f))
                        s_todo[atomicAdd(&s_flags[par][0], 1)] = (ly << 8) | lx;  // undecided: exact f64 evaluation below
                }
            }
        }
        uint8_t* ep = edge + (size_t)cb * estride_b;
        if (in_img) ep[(size_t)y * w + x] = e;
        __syncthreads();
        const int ntodo = s_flags[par][0];  // block-uniform
        for (int t = tid; t < ntodo; t += ET_W * ET_H) {
            const int ply = s_todo[t] >> 8, plx = s_todo[t] & 255;  // normal tile coordinates (halo 1)
            const int px = cx0 + plx - 1, py = cy0 + ply - 1;
            const uint8_t ex = edge_decide_f64([&](int k, double n[3], double v[3]) {
                const int dx = k < 8 ? edge_nx(k) : 0, dy = k < 8 ? edge_ny(k) : 0;
                edge_nv_f64(sd, ply + 1 + dy, plx + 1 + dx, px + dx, py + dy, w, h, cam, n, v);
            });
            ep[(size_t)py * w + px] = ex;
        }
    }
}

// the all-f64 form: one normal / vertex per pixel in shared memory (f64), then the reference's neighbour loop
__global__ void __launch_bounds__(ET_W* ET_H) k_depth_edge_f64(const float* __restrict__ depth, size_t dstride_b, int w, int h,
                                                                CamConst cam, uint8_t* __restrict__ edge, size_t estride_b)
{
    pdl_wait();
    __shared__ float sd[ET_H + 4][ET_W + 4];          // clamped depth, halo 2
    __shared__ double sn[ET_H + 2][ET_W + 2][3];      // normals, halo 1
    __shared__ double sv[ET_H + 2][ET_W + 2][3];      // vertices, halo 1
    const int b = blockIdx.z;
    const float* dp = depth + (size_t)b * dstride_b;
    const int x0 = blockIdx.x * ET_W, y0 = blockIdx.y * ET_H;
    const int tid = threadIdx.y * ET_W + threadIdx.x;
    load_depth_tile(dp, w, h, x0, y0, threadIdx.x, threadIdx.y, sd);
    __syncthreads();
    for (int i = tid; i < (ET_H + 2) * (ET_W + 2); i += ET_W * ET_H) {
        const int ly = i / (ET_W + 2), lx = i - ly * (ET_W + 2);
        double n[3], v[3];
        edge_nv_f64(sd, ly + 1, lx + 1, x0 + lx - 1, y0 + ly - 1, w, h, cam, n, v);
        sn[ly][lx][0] = n[0]; sn[ly][lx][1] = n[1]; sn[ly][lx][2] = n[2];
        sv[ly][lx][0] = v[0]; sv[ly][lx][1] = v[1]; sv[ly][lx][2] = v[2];
    }
    __syncthreads();
    const int x = x0 + threadIdx.x, y = y0 + threadIdx.y;
    if (x >= w || y >= h) return;
    uint8_t e = 0;
    const int lx = threadIdx.x + 1, ly = threadIdx.y + 1;
    if (x >= 1 && y >= 1 && x < w - 1 && y < h - 1 && sd[ly + 1][lx + 1] != 0.f)
        e = edge_decide_f64([&](int k, double n[3], double v[3]) {
            const int jx = lx + (k < 8 ? edge_nx(k) : 0), jy = ly + (k < 8 ? edge_ny(k) : 0);
            n[0] = sn[jy][jx][0]; n[1] = sn[jy][jx][1]; n[2] = sn[jy][jx][2];
            v[0] = sv[jy][jx][0]; v[1] = sv[jy][jx][1]; v[2] = sv[jy][jx][2];
        });
    edge[(size_t)b * estride_b + (size_t)y * w + x] = e;
}

int launch_depth_edge(const float* depth, size_t depth_stride_b, int w, int h, int batch, const CamConst& cam,
                      uint8_t* edge, size_t edge_stride_b, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "K2a_depth_edge", 1);
    dim3 block(ET_W, ET_H), grid(cdiv(w, ET_W), cdiv(h, ET_H), batch);
    const char* env = std::getenv("GD_EDGE_F64");  // read per launch: the parity test toggles it inside one process
    const bool force_f64 = env && std::atoi(env) != 0;
    // the f32 interval kernel assumes z = depth (third row of inv(K) = (0, 0, 1)) and finite bounds
    // (inv(K)[2][2] = fx fy * (1 / (fx fy)) may differ from 1 in the last bit: z = that * depth is then within 2^-52 of the
    // depth the kernel uses, far inside the bounds, and it is zero exactly when the depth is)
    const bool plain_k = cam.Kid[6] == 0.0 && cam.Kid[7] == 0.0 && std::fabs(cam.Kid[8] - 1.0) < 1e-12;
    double S = 0.0;
    for (int r = 0; r < 2; ++r)
        S = std::max(S, std::fabs(cam.Kid[3 * r]) * (w - 1) + std::fabs(cam.Kid[3 * r + 1]) * (h - 1) + std::fabs(cam.Kid[3 * r + 2]));
    S = std::max(S, 1.0);
    const double u = 5.9604644775390625e-08;  // 2^-24
    const double e_d = 256.0 * u * S * 3.5;
    if (force_f64 || !plain_k || !(e_d < 1e-2)) {
        GD_CUDA(launch_pdl(k_depth_edge_f64, grid, block, 0, s, depth, depth_stride_b, w, h, cam, edge, edge_stride_b));
    } else {
        EdgeBounds eb;
        for (int i = 0; i < 6; ++i) eb.Ki[i] = (float)cam.Kid[i];
        eb.e_d = (float)(e_d * 1.0000002);  // rounded up
        eb.e_c = (float)(64.0 * u);
        EdgeTiles et;
        et.tiles_x = (int)grid.x;
        et.tiles_per_img = (int)(grid.x * grid.y);
        et.total = et.tiles_per_img * batch;
        auto magic = [](unsigned d) { return d <= 1u ? 0xFFFFFFFFu : (unsigned)(((1ull << 32) + d - 1) / d); };
        et.magic_x = magic((unsigned)et.tiles_x);
        et.magic_img = magic((unsigned)et.tiles_per_img);
        const bool magic_ok = et.tiles_x > 1 && (unsigned long long)et.total * et.tiles_per_img < (1ull << 32);
        if (!magic_ok) {
            GD_CUDA(launch_pdl(k_depth_edge_f64, grid, block, 0, s, depth, depth_stride_b, w, h, cam, edge, edge_stride_b));
        } else {
            static int n_sm = 0;
            if (!n_sm) {
                int dev = 0;
                GD_CUDA(cudaGetDevice(&dev));
                GD_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
            }
            // persistent (2 CTAs per SM walking the tile list) only when every CTA gets at least 8 tiles; a small job runs
            // one tile per CTA, so that the kernels of other streams and handles keep finding free SM resources
            if (et.total >= 8 * 2 * n_sm) {
                GD_CUDA(launch_plain(k_depth_edge, dim3((unsigned)(2 * n_sm)), block, 0, s, depth, depth_stride_b, w, h, cam, eb, et, edge,
                                     edge_stride_b));
            } else {
                GD_CUDA(launch_pdl(k_depth_edge, dim3((unsigned)et.total), block, 0, s, depth, depth_stride_b, w, h, cam, eb, et, edge,
                                   edge_stride_b));
            }
        }
    }
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------
// K2b Mahalanobis + scatter.  One thread per source pixel.  Scatter = 64-bit atomicMax on
// ((src_index+1) << 32 | float_bits(value)): the high word orders writers by raster index, so the
// last raster-order writer wins exactly like the sequential loop (GeoMaskMaker.cc:269).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot3f(float a0, float b0, float a1, float b1, float a2, float b2)
{
    float t = a0 * b0;
    t = t + a1 * b1;
    t = t + a2 * b2;
    return t;
}

__device__ __forceinline__ float depth2std(float depth, float std_k)
{
    float r = std_k;  // ((1/fu)^2 * 0.5) * 0.5, rounded product by product on the host
    r = r * depth;
    r = r * depth;
    r = r * depth;
    r = r * depth;
    return r;
}

// Correctly rounded x / d for a divisor known in advance (Markstein: q = RN(x r), q' = RN(q + RN(x - q d) r) with
// r = RN(1/d) is RN(x/d) whenever the significand of d is not all ones and nothing over/underflows — focal lengths and
// metre-scale operands here).  Three FMA-class instructions instead of the IEEE division sequence; the zero it returns for
// x = -0 is +0, which only feeds products that are added to non-negative-zero sums.
__device__ __forceinline__ float div_by(float x, float d, float r)
{
    const float q = x * r;
    const float rem = fmaf(-q, d, x);
    return fmaf(rem, r, q);
}

// MB: resident CTAs per SM the register allocation aims at (6: 40 registers with ~140 bytes of spills, 4: 64 registers, none)
template <int MB>
__global__ void __launch_bounds__(256, MB) k_mahalanobis(const float2* __restrict__ flow, size_t fstride_b,
                                                     const float* __restrict__ depth_ref, const float* __restrict__ depth_cur,
                                                     size_t dstride_b, const uint8_t* __restrict__ edge_ref,
                                                     const uint8_t* __restrict__ edge_cur, size_t estride_b,
                                                     const float2* __restrict__ lut, int w, int h, CamConst cam,
                                                     const PoseDev* __restrict__ poses, int key_shift,
                                                     unsigned long long* __restrict__ keys, size_t kstride_b)
{
    pdl_wait();
    const int b = blockIdx.z;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    // the stream's pose goes through shared memory: 24 words read once per CTA instead of 24 dependent global loads per thread
    __shared__ PoseDev s_pose;
    {
        const int t = threadIdx.y * blockDim.x + threadIdx.x;
        if (t < (int)(sizeof(PoseDev) / 4)) reinterpret_cast<int*>(&s_pose)[t] = reinterpret_cast<const int*>(poses + b)[t];
    }
    const bool in_img = x < w && y < h;
    const size_t i = in_img ? (size_t)y * w + x : 0;
    // everything that depends only on the source pixel is requested before the flow arrives (one dependent round less)
    const float2 f = __ldg(flow + (size_t)b * fstride_b + i);
    __syncthreads();
    const PoseDev* __restrict__ P = &s_pose;
    if (!in_img || !P->valid) return;
    float rx = (float)x, ry = (float)y;
    if (lut) {
        const float2 a = __ldg(lut + i);
        rx = a.x; ry = a.y;
    }
    // without a LUT the undistorted positions are the integer pixel positions themselves: no float <-> int round trips
    // (they run on the conversion pipe, which bounds this kernel)
    const int rix = lut ? (int)rx : x, riy = lut ? (int)ry : y;
    const bool ref_in = rix >= 0 && riy >= 0 && rix < w && riy < h;
    const size_t ri = ref_in ? (size_t)riy * w + rix : 0;
    const float ref_depth = __ldg(depth_ref + (size_t)b * dstride_b + ri);
    const uint8_t ref_edge = __ldg(edge_ref + (size_t)b * estride_b + ri);
    const float cur_x = (float)x + f.x;
    const float cur_y = (float)y + f.y;
    if (!(cur_x == cur_x) || !(cur_y == cur_y)) return;
    if (cur_x < 0 || cur_y < 0 || cur_x > (float)(w - 1) || cur_y > (float)(h - 1)) return;
    const int icx = (int)cur_x, icy = (int)cur_y;
    float cx = (float)icx, cy = (float)icy;
    if (lut) {
        const float2 c = __ldg(lut + (size_t)icy * w + icx);
        cx = c.x; cy = c.y;
    }
    const int cix = lut ? (int)cx : icx, ciy = lut ? (int)cy : icy;
    if (!ref_in || cix < 0 || ciy < 0 || cix >= w || ciy >= h) return;
    const size_t ci = (size_t)ciy * w + cix;
    const float cur_depth = __ldg(depth_cur + (size_t)b * dstride_b + ci);
    const uint8_t cur_edge = __ldg(edge_cur + (size_t)b * estride_b + ci);
    if (ref_edge == 255 || cur_edge == 255) return;
    if (cur_depth == 0.f || cur_depth > 3.5f || ref_depth == 0.f || ref_depth > 3.5f) return;  // (double)d > 3.5 <=> d > 3.5f

    const float fu = cam.fu, fv = cam.fv, cu = cam.cu;
    const float U0 = dot3f(P->RK[0], rx, P->RK[1], ry, P->RK[2], 1.0f);
    const float U1 = dot3f(P->RK[3], rx, P->RK[4], ry, P->RK[5], 1.0f);
    const float U2 = dot3f(P->RK[6], rx, P->RK[7], ry, P->RK[8], 1.0f);
    const float C0 = dot3f(cam.Ki[0], cx, cam.Ki[1], cy, cam.Ki[2], 1.0f) * cur_depth;
    const float C1 = dot3f(cam.Ki[3], cx, cam.Ki[4], cy, cam.Ki[5], 1.0f) * cur_depth;
    const float C2 = dot3f(cam.Ki[6], cx, cam.Ki[7], cy, cam.Ki[8], 1.0f) * cur_depth;
    const float P0 = fmaf(U0, ref_depth, P->T[0]);  // cv::scaleAdd is fused in OpenCV 4.13
    const float P1 = fmaf(U1, ref_depth, P->T[1]);
    const float P2 = fmaf(U2, ref_depth, P->T[2]);
    const float e0 = C0 - P0, e1 = C1 - P1, e2 = C2 - P2;

    const float s2 = depth2std(ref_depth, cam.std_k);
    const float s5 = depth2std(cur_depth, cam.std_k);
    const float rfu = cam.rfu, rfv = cam.rfv;
    // J rows (index slips of the reference kept: J(1,1) uses ref_depth, J(1,2) uses x)
    float J[3][6];
    const float dxc = cx - cu;
    J[0][0] = div_by(cur_depth, fu, rfu);  J[0][1] = 0.f;  J[0][2] = div_by(dxc, fu, rfu);
    J[0][3] = div_by(-P->R[0] * ref_depth, fu, rfu); J[0][4] = div_by(-P->R[1] * ref_depth, fv, rfv); J[0][5] = -U0;
    J[1][0] = 0.f;  J[1][1] = div_by(ref_depth, fv, rfv);  J[1][2] = div_by(dxc, fv, rfv);
    J[1][3] = div_by(-P->R[3] * ref_depth, fu, rfu); J[1][4] = div_by(-P->R[4] * ref_depth, fv, rfv); J[1][5] = -U1;
    J[2][0] = 0.f;  J[2][1] = 0.f;  J[2][2] = 1.0f;
    J[2][3] = div_by(-P->R[6] * ref_depth, fu, rfu); J[2][4] = div_by(-P->R[7] * ref_depth, fv, rfv); J[2][5] = -U2;
    // (J S) J^T like cv::gemm: f32 operands, f64 accumulator, k ascending.  The product of two f32 is exact in f64, so
    // fma(a, b, s) rounds exactly like s + a * b; terms with a structural zero of J add +-0 to a sum that is never -0
    // and are skipped (J(0,1) = J(1,0) = J(2,0) = J(2,1) = 0).  Bit-identical to the dense loop, 38 DFMA instead of 108 ops.
    float Cm[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        double JS[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) JS[c] = (double)J[r][c];
        JS[2] = (double)(J[r][2] * s2);
        JS[5] = (double)(J[r][5] * s5);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                const bool zr = (k == 1 && r != 1) || (k == 0 && r != 0);
                const bool zc = (k == 1 && c != 1) || (k == 0 && c != 0);
                if (!zr && !zc) s = fma(JS[k], (double)J[c][k], s);
            }
            Cm[3 * r + c] = (float)s;
        }
    }
    float Ci[9];
    inv3_cv<float>(Cm, Ci);
    float q[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double s = fma((double)e0, (double)Ci[c], 0.0);  // exact products: fma == multiply then add
        s = fma((double)e1, (double)Ci[3 + c], s);
        s = fma((double)e2, (double)Ci[6 + c], s);
        q[c] = (float)s;
    }
    double l = fma((double)q[0], (double)e0, 0.0);
    l = fma((double)q[1], (double)e1, l);
    l = fma((double)q[2], (double)e2, l);
    float value = sqrtf((float)l);
    value = value + 0.0f;  // -0 -> +0 so that the bit pattern orders like the value
    const unsigned long long key = ((unsigned long long)(unsigned)P->epoch << key_shift) | ((unsigned long long)(unsigned)(i + 1) << 32) |
                                   (unsigned long long)__float_as_uint(value);
    atomicMax(keys + (size_t)b * kstride_b + (size_t)icy * w + icx, key);
}

int launch_mahalanobis(const float2* flow, size_t flow_stride_b, const float* depth_ref, const float* depth_cur,
                       size_t depth_stride_b, const uint8_t* edge_ref, const uint8_t* edge_cur, size_t edge_stride_b,
                       const float2* lut, int w, int h, int batch, const CamConst& cam, const PoseDev* poses, KeyFormat kf,
                       unsigned long long* keys, size_t keys_stride_b, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "K2b_mahalanobis", 1);
    dim3 block(32, 8), grid(cdiv(w, 32), cdiv(h, 8), batch);
    const char* env = std::getenv("GD_MAHA_MB");
    if (env && std::atoi(env) == 5)
        GD_CUDA(launch_pdl(k_mahalanobis<5>, grid, block, 0, s, flow, flow_stride_b, depth_ref, depth_cur, depth_stride_b, edge_ref, edge_cur,
                           edge_stride_b, lut, w, h, cam, poses, kf.shift, keys, keys_stride_b));
    else if (env && std::atoi(env) == 4)
        GD_CUDA(launch_pdl(k_mahalanobis<4>, grid, block, 0, s, flow, flow_stride_b, depth_ref, depth_cur, depth_stride_b, edge_ref, edge_cur,
                           edge_stride_b, lut, w, h, cam, poses, kf.shift, keys, keys_stride_b));
    else
        GD_CUDA(launch_pdl(k_mahalanobis<6>, grid, block, 0, s, flow, flow_stride_b, depth_ref, depth_cur, depth_stride_b, edge_ref, edge_cur,
                           edge_stride_b, lut, w, h, cam, poses, kf.shift, keys, keys_stride_b));
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------
// K3 pass 1: min / max over the resolved low words.  Unwritten pixels (stale epoch) contribute 0.0f; NaN values are
// skipped like cv::normalize's min/max scan skips them — except at pixel 0, where the scan starts (flag word [2]).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned key_value_bits(unsigned long long key, int key_shift, unsigned epoch)
{
    return (unsigned)(key >> key_shift) == epoch ? (unsigned)key : 0u;
}
__device__ __forceinline__ bool bits_is_nan(unsigned b) { return (b & 0x7FFFFFFFu) > 0x7F800000u; }

__global__ void __launch_bounds__(256) k_minmax(const unsigned long long* __restrict__ keys, size_t kstride_b, int n_px,
                                                const PoseDev* __restrict__ poses, int key_shift,
                                                unsigned int* __restrict__ minmax_bits)
{
    pdl_wait();
    const int b = blockIdx.y;
    const unsigned long long* kp = keys + (size_t)b * kstride_b;
    const unsigned epoch = (unsigned)poses[b].epoch;
    unsigned mn = 0xFFFFFFFFu, mx = 0u;
    const int n2 = n_px >> 1;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += gridDim.x * blockDim.x) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(kp) + i);
        const unsigned a = key_value_bits(v.x, key_shift, epoch), c = key_value_bits(v.y, key_shift, epoch);
        if (!bits_is_nan(a)) {
            mn = min(mn, a);
            mx = max(mx, a);
        } else if (i == 0) {
            minmax_bits[GD_MM_WORDS * b + 2] = 0u;
        }
        if (!bits_is_nan(c)) {
            mn = min(mn, c);
            mx = max(mx, c);
        }
    }
    if ((n_px & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const unsigned a = key_value_bits(kp[n_px - 1], key_shift, epoch);
        if (!bits_is_nan(a)) {
            mn = min(mn, a);
            mx = max(mx, a);
        } else if (n_px == 1) {
            minmax_bits[GD_MM_WORDS * b + 2] = 0u;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    __shared__ unsigned smn[8], smx[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
        smn[warp] = mn;
        smx[warp] = mx;
    }
    __syncthreads();
    if (warp == 0) {
        mn = lane < (blockDim.x >> 5) ? smn[lane] : 0xFFFFFFFFu;
        mx = lane < (blockDim.x >> 5) ? smx[lane] : 0u;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (lane == 0) {
            atomicMin(minmax_bits + GD_MM_WORDS * b, mn);
            atomicMin(minmax_bits + GD_MM_WORDS * b + 1, ~mx);
        }
    }
}

int launch_minmax_reset(unsigned int* minmax_bits, int batch, cudaStream_t s)
{
    GD_CUDA(cudaMemsetAsync(minmax_bits, 0xFF, sizeof(unsigned) * GD_MM_WORDS * batch, s));
    return GD_OK;
}

int launch_minmax(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses, KeyFormat kf,
                  unsigned int* minmax_bits, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "K3a_minmax", 1);
    int blocks = cdiv(n_px / 2, 256 * 4);
    if (blocks < 1) blocks = 1;
    dim3 grid(blocks, batch);
    GD_CUDA(launch_pdl(k_minmax, grid, dim3(256), 0, s, keys, keys_stride_b, n_px, poses, kf.shift, minmax_bits));
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------
// K3 pass 2: v*a+b (fused, like cv convertTo), round half even, saturate, (<20) -> {1,0}
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned mask_of(float d, float a, float bsh)
{
    const float v = fmaf(d, a, bsh);
    int r = (v == v) ? __float2int_rn(v) : 0;
    r = max(0, min(255, r));
    return r < 20 ? 1u : 0u;
}

__global__ void __launch_bounds__(256) k_normalize_mask(const unsigned long long* __restrict__ keys, size_t kstride_b, int n_px,
                                                        const unsigned int* __restrict__ minmax_bits,
                                                        const PoseDev* __restrict__ poses, int key_shift,
                                                        uint8_t* __restrict__ mask, size_t mstride_b)
{
    pdl_wait();
    const int b = blockIdx.y;
    const unsigned long long* kp = keys + (size_t)b * kstride_b;
    uint8_t* mp = mask + (size_t)b * mstride_b;
    const unsigned epoch = (unsigned)poses[b].epoch;
    // invalid pose -> all ones (:179-185); NaN at pixel 0 -> cv::normalize turns the whole image into NaN -> 8-bit 0 -> all ones
    const bool valid = poses[b].valid != 0 && minmax_bits[GD_MM_WORDS * b + 2] != 0u;
    const float smin_f = __uint_as_float(minmax_bits[GD_MM_WORDS * b]);
    const float smax_f = __uint_as_float(~minmax_bits[GD_MM_WORDS * b + 1]);
    const double smin = (double)smin_f, smax = (double)smax_f;
    const double scale = 255.0 * ((smax - smin) > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, bsh = (float)shift;
    const int g = blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
    const int i0 = g * 4;
    if (i0 >= n_px) return;
    if (i0 + 4 <= n_px && (mstride_b & 3) == 0 && (kstride_b & 1) == 0) {
        const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(kp + i0);
        const ulonglong2 v0 = __ldg(k2), v1 = __ldg(k2 + 1);
        const float d0 = __uint_as_float(key_value_bits(v0.x, key_shift, epoch)), d1 = __uint_as_float(key_value_bits(v0.y, key_shift, epoch));
        const float d2 = __uint_as_float(key_value_bits(v1.x, key_shift, epoch)), d3 = __uint_as_float(key_value_bits(v1.y, key_shift, epoch));
        unsigned m = 0x01010101u;
        if (valid) m = mask_of(d0, a, bsh) | (mask_of(d1, a, bsh) << 8) | (mask_of(d2, a, bsh) << 16) | (mask_of(d3, a, bsh) << 24);
        *reinterpret_cast<unsigned*>(mp + i0) = m;
    } else {
        for (int i = i0; i < min(i0 + 4, n_px); ++i) {
            const float d = __uint_as_float(key_value_bits(kp[i], key_shift, epoch));
            mp[i] = valid ? (uint8_t)mask_of(d, a, bsh) : (uint8_t)1;
        }
    }
}

int launch_normalize_mask(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch,
                          const unsigned int* minmax_bits, const PoseDev* poses, KeyFormat kf, uint8_t* mask, size_t mask_stride_b,
                          cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "K3b_normalize_mask", 1);
    dim3 grid(cdiv(cdiv(n_px, 4), 256), batch);
    GD_CUDA(launch_pdl(k_normalize_mask, grid, dim3(256), 0, s, keys, keys_stride_b, n_px, minmax_bits, poses, kf.shift, mask, mask_stride_b));
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------
// K3, single pass: a thread-block cluster of K3_CS CTAs per stream.  Every CTA resolves its slice of the key image into
// shared memory (f32, 4 bytes per pixel) while it reduces min / max; the partial results are exchanged through distributed
// shared memory, cluster.sync(), and each CTA normalises + thresholds its slice straight from shared memory.  The key image
// is read ONCE (8 bytes per pixel) and the mask written once: 9 bytes per pixel, the algorithmic minimum for this layout
// (the two-kernel form reads the keys twice: 17).  Used when a slice fits one CTA's shared memory (640 x 480: 150 KB).
// ------------------------------------------------------------------------------------------------
constexpr int K3_CS = 8;          // portable cluster size
constexpr int K3_THREADS = 1024;

__global__ void __cluster_dims__(K3_CS, 1, 1) __launch_bounds__(K3_THREADS, 1)
    k_minmax_mask_cluster(const unsigned long long* __restrict__ keys, size_t kstride_b, int n_px, int slice,
                          const PoseDev* __restrict__ poses, int key_shift, unsigned int* __restrict__ minmax_bits,
                          uint8_t* __restrict__ mask, size_t mstride_b)
{
    extern __shared__ __align__(16) unsigned char k3_sm[];
    unsigned* vals = reinterpret_cast<unsigned*>(k3_sm);  // [slice] resolved value bits of this CTA's pixels
    __shared__ unsigned s_part[4];                        // min bits, max bits, poison flag (pixel 0 is NaN), -
    __shared__ unsigned s_wmn[K3_THREADS / 32], s_wmx[K3_THREADS / 32];
    cg::cluster_group cluster = cg::this_cluster();
    const int b = blockIdx.y, rank = (int)cluster.block_rank(), tid = threadIdx.x;
    const unsigned long long* kp = keys + (size_t)b * kstride_b;
    const unsigned epoch = (unsigned)poses[b].epoch;
    const int i0 = rank * slice, i1 = max(i0, min(n_px, i0 + slice));  // slice is a multiple of 4; empty for tiny images
    unsigned mn = 0xFFFFFFFFu, mx = 0u, poison = 0u;
    if (tid == 0) s_part[2] = 0u;
    __syncthreads();
    // pass 1: two keys (16 bytes) per load, four loads in flight per thread
    const int npair = (i1 - i0) >> 1;
    const ulonglong2* k2 = reinterpret_cast<const ulonglong2*>(kp + i0);
    for (int p = tid; p < npair; p += K3_THREADS * 4) {
        ulonglong2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (p + u * K3_THREADS < npair) v[u] = __ldg(k2 + p + u * K3_THREADS);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int q = p + u * K3_THREADS;
            if (q >= npair) break;
            const unsigned a = key_value_bits(v[u].x, key_shift, epoch), c = key_value_bits(v[u].y, key_shift, epoch);
            reinterpret_cast<uint2*>(vals)[q] = make_uint2(a, c);
            if (!bits_is_nan(a)) {
                mn = min(mn, a);
                mx = max(mx, a);
            } else if (i0 + 2 * q == 0) {
                poison = 1u;
            }
            if (!bits_is_nan(c)) {
                mn = min(mn, c);
                mx = max(mx, c);
            }
        }
    }
    if (((i1 - i0) & 1) && tid == 0) {  // odd tail (only the last slice of an odd-sized image)
        const unsigned a = key_value_bits(kp[i1 - 1], key_shift, epoch);
        vals[i1 - 1 - i0] = a;
        if (!bits_is_nan(a)) {
            mn = min(mn, a);
            mx = max(mx, a);
        } else if (i1 - 1 == 0) {
            poison = 1u;
        }
    }
    if (poison) s_part[2] = 1u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((tid & 31) == 0) {
        s_wmn[tid >> 5] = mn;
        s_wmx[tid >> 5] = mx;
    }
    __syncthreads();
    if (tid < 32) {
        mn = s_wmn[tid];
        mx = s_wmx[tid];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        }
        if (tid == 0) {
            s_part[0] = mn;
            s_part[1] = mx;
        }
    }
    cluster.sync();  // every CTA's partial result is in its shared memory
    unsigned gmn = 0xFFFFFFFFu, gmx = 0u, gpoison = 0u;
#pragma unroll
    for (int r = 0; r < K3_CS; ++r) {
        const unsigned* rp = cluster.map_shared_rank(s_part, r);
        gmn = min(gmn, rp[0]);
        gmx = max(gmx, rp[1]);
        gpoison |= rp[2];
    }
    cluster.sync();  // nobody leaves (and frees its shared memory) while a peer may still read it
    if (rank == 0 && tid == 0) {  // same encoding as the two-kernel form (debug fetch)
        minmax_bits[GD_MM_WORDS * b] = gmn;
        minmax_bits[GD_MM_WORDS * b + 1] = ~gmx;
        minmax_bits[GD_MM_WORDS * b + 2] = gpoison ? 0u : 0xFFFFFFFFu;
    }
    const bool valid = poses[b].valid != 0 && gpoison == 0u;
    const double smin = (double)__uint_as_float(gmn), smax = (double)__uint_as_float(gmx);
    const double scale = 255.0 * ((smax - smin) > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, bsh = (float)shift;
    uint8_t* mp = mask + (size_t)b * mstride_b;
    const int nquad = (i1 - i0) >> 2;
    for (int q = tid; q < nquad; q += K3_THREADS) {
        const uint4 v = reinterpret_cast<const uint4*>(vals)[q];
        unsigned m = 0x01010101u;
        if (valid)
            m = mask_of(__uint_as_float(v.x), a, bsh) | (mask_of(__uint_as_float(v.y), a, bsh) << 8) |
                (mask_of(__uint_as_float(v.z), a, bsh) << 16) | (mask_of(__uint_as_float(v.w), a, bsh) << 24);
        *reinterpret_cast<unsigned*>(mp + i0 + 4 * q) = m;
    }
    for (int i = i0 + 4 * nquad + tid; i < i1; i += K3_THREADS)  // tail of an image whose size is not a multiple of 4
        mp[i] = valid ? (uint8_t)mask_of(__uint_as_float(vals[i - i0]), a, bsh) : (uint8_t)1;
}

// returns GD_OK and sets *used when the cluster form fits (slice in shared memory, aligned strides); otherwise *used = false
int launch_minmax_mask_cluster(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses,
                               KeyFormat kf, unsigned int* minmax_bits, uint8_t* mask, size_t mask_stride_b, cudaStream_t s,
                               LaunchStats* st, bool* used)
{
    *used = false;
    const char* env = std::getenv("GD_K3_CLUSTER");
    if (env && std::atoi(env) == 0) return GD_OK;
    const int slice = (int)align_up((size_t)cdiv(n_px, K3_CS), 4);
    const size_t smem = (size_t)slice * 4;
    if (smem > 200 * 1024 || (mask_stride_b & 3) != 0 || (keys_stride_b & 1) != 0 || (slice & 1) != 0) return GD_OK;
    static thread_local int attr_device = -1;
    int dev = 0;
    GD_CUDA(cudaGetDevice(&dev));
    if (attr_device != dev) {
        GD_CUDA(cudaFuncSetAttribute(k_minmax_mask_cluster, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_device = dev;
    }
    LaunchScope ls(st, s, "K3_minmax_mask", 1);
    k_minmax_mask_cluster<<<dim3(K3_CS, batch), K3_THREADS, smem, s>>>(keys, keys_stride_b, n_px, slice, poses, kf.shift, minmax_bits, mask,
                                                                      mask_stride_b);
    GD_CUDA(cudaGetLastError());
    *used = true;
    return GD_OK;
}

__global__ void __launch_bounds__(256) k_resolve_dist(const unsigned long long* __restrict__ keys, size_t kstride_b, int n_px,
                                                      const PoseDev* __restrict__ poses, int key_shift, float* __restrict__ dist,
                                                      size_t dstride_b)
{
    const int b = blockIdx.y;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_px) return;
    dist[(size_t)b * dstride_b + i] = __uint_as_float(key_value_bits(keys[(size_t)b * kstride_b + i], key_shift, (unsigned)poses[b].epoch));
}

int launch_resolve_dist(const unsigned long long* keys, size_t keys_stride_b, int n_px, int batch, const PoseDev* poses, KeyFormat kf,
                        float* dist_out, size_t dist_stride_b, cudaStream_t s)
{
    dim3 grid(cdiv(n_px, 256), batch);
    k_resolve_dist<<<grid, 256, 0, s>>>(keys, keys_stride_b, n_px, poses, kf.shift, dist_out, dist_stride_b);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

__global__ void k_fill_u8(uint8_t* dst, size_t n, uint8_t v)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = v;
}

int launch_fill_u8(uint8_t* dst, size_t n, uint8_t v, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "fill_u8", 1);
    k_fill_u8<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(dst, n, v);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd

// ------------------------------------------------------------------------------------------------
// "next" row (f)-2 of SURVEY section 8: the Frame constructor's mask erosion + keypoint filter
// (GD-SLAM src/Frame.cc:258-282): erode(mask, 31x31 MORPH_ELLIPSE) and keep keypoint i iff
// eroded((int)pt.y, (int)pt.x) == 1.  The eroded image is only ever sampled at the keypoints, so it is evaluated there:
// one warp per keypoint (lane = structuring-element row), then an order-preserving compaction per stream.
// ------------------------------------------------------------------------------------------------
namespace gd {

void make_ellipse31(EllipseSE* se)
{
    const int r = 15, c = 15;
    const double inv_r2 = 1.0 / ((double)r * r);
    for (int i = 0; i < 31; ++i) {  // cv::getStructuringElement(MORPH_ELLIPSE)
        const int dy = i - r;
        const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
        se->j1[i] = c - dx > 0 ? c - dx : 0;
        se->j2[i] = c + dx + 1 < 31 ? c + dx + 1 : 31;
    }
}

__global__ void __launch_bounds__(256) k_erode_filter(const uint8_t* __restrict__ mask, size_t mstride_b, int w, int h,
                                                      const gd_keypoint* __restrict__ kps, size_t kstride_b,
                                                      const int* __restrict__ n_kp, int n_fixed, EllipseSE se,
                                                      uint8_t* __restrict__ keep, size_t keep_stride_b)
{
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int n = n_kp ? n_kp[b] : n_fixed;
    if (g >= n) return;
    const gd_keypoint kp = kps[(size_t)b * kstride_b + g];
    const int x = (int)kp.x, y = (int)kp.y;  // Mask_dil.at<uchar>(pt.y, pt.x): float -> int truncation
    const uint8_t* m = mask + (size_t)b * mstride_b;
    int mn = 255;
    if (lane < 31) {
        const int yy = y + lane - 15;
        if (yy >= 0 && yy < h) {
            const int xa = max(x + se.j1[lane] - 15, 0), xb = min(x + se.j2[lane] - 15, w);  // border: outside pixels ignored
            for (int xx = xa; xx < xb; ++xx) mn = min(mn, (int)__ldg(m + (size_t)yy * w + xx));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(0xffffffffu, mn, o));
    if (lane == 0) keep[(size_t)b * keep_stride_b + g] = (uint8_t)(mn == 1);
}

__global__ void __launch_bounds__(256) k_compact_keypoints(const gd_keypoint* __restrict__ kps, const uint8_t* __restrict__ desc,
                                                           const uint8_t* __restrict__ keep, size_t cap, const int* __restrict__ n_kp,
                                                           gd_keypoint* __restrict__ out_kps, uint8_t* __restrict__ out_desc,
                                                           int* __restrict__ out_n)
{
    __shared__ int s_warp[8];
    __shared__ int s_base;
    const int b = blockIdx.x, n = n_kp[b];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_base = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 256) {
        const int i = base + threadIdx.x;
        const bool k = i < n && keep[b * cap + i];
        const unsigned bal = __ballot_sync(0xffffffffu, k);
        if (lane == 0) s_warp[warp] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int w2 = 0; w2 < 8; ++w2) {
            if (w2 < warp) woff += s_warp[w2];
            tot += s_warp[w2];
        }
        if (k) {
            const size_t o = b * cap + s_base + woff + __popc(bal & ((1u << lane) - 1));
            out_kps[o] = kps[b * cap + i];
            const uint4* s4 = reinterpret_cast<const uint4*>(desc + (b * cap + i) * 32);
            uint4* d4 = reinterpret_cast<uint4*>(out_desc + o * 32);
            d4[0] = s4[0];
            d4[1] = s4[1];
        }
        __syncthreads();
        if (threadIdx.x == 0) s_base += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) out_n[b] = s_base;
}

__global__ void __launch_bounds__(256) k_depth_u16_to_m(const uint16_t* __restrict__ raw, size_t rstride_b, float* __restrict__ depth,
                                                        size_t dstride_b, size_t n, float inv_factor)
{
    const size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const uint16_t* r = raw + (size_t)blockIdx.y * rstride_b;
    float* d = depth + (size_t)blockIdx.y * dstride_b;
    if (i + 4 <= n && (rstride_b & 3) == 0 && (dstride_b & 3) == 0) {
        const ushort4 v = *reinterpret_cast<const ushort4*>(r + i);
        *reinterpret_cast<float4*>(d + i) = make_float4((float)v.x * inv_factor, (float)v.y * inv_factor, (float)v.z * inv_factor,
                                                        (float)v.w * inv_factor);
    } else {
        for (size_t k = i; k < n && k < i + 4; ++k) d[k] = (float)r[k] * inv_factor;
    }
}

int launch_depth_u16_to_m(const uint16_t* raw, size_t raw_stride_b, float* depth, size_t depth_stride_b, size_t n, int batch,
                          float inv_factor, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "F4_depth_u16_to_m", 1);
    dim3 grid((unsigned)((n / 4 + 255) / 256 + 1), batch);
    k_depth_u16_to_m<<<grid, 256, 0, s>>>(raw, raw_stride_b, depth, depth_stride_b, n, inv_factor);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

constexpr int SG_MAX = 4096;     // keypoints per stream the grid kernel sorts in shared memory
constexpr int SG_CELLS = 64 * 48;

struct GridBounds {
    float minx, miny, inv_w, inv_h;  // mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv
    int distorted;
};

__global__ void __launch_bounds__(512) k_stereo_grid(const float* __restrict__ depth, size_t dstride_b, int w, int h,
                                                     const gd_keypoint* __restrict__ kps, size_t cap, const int* __restrict__ n_kp,
                                                     float bf, UndistortArgs und, GridBounds gb, float* __restrict__ depth_out,
                                                     float* __restrict__ uright, int* __restrict__ cell_start,
                                                     int* __restrict__ cell_items, float2* __restrict__ un_out)
{
    __shared__ unsigned keys[SG_MAX];  // (cell << 12 | index), 0xFFFFFFFF = not in the grid / padding
    const int b = blockIdx.x, n = min(n_kp[b], SG_MAX);
    const float* dp = depth + (size_t)b * dstride_b;
    int npad = 1;
    while (npad < n) npad <<= 1;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        unsigned key = 0xFFFFFFFFu;
        if (i < n) {
            const gd_keypoint kp = kps[(size_t)b * cap + i];
            float ux = kp.x, uy = kp.y;  // mvKeysUn (Frame.cc:576-606)
            if (gb.distorted) undistort_point_cv(und, kp.x, kp.y, &ux, &uy);
            if (un_out) un_out[(size_t)b * cap + i] = make_float2(ux, uy);
            const float d = dp[(size_t)(int)kp.y * w + (int)kp.x];  // depth at the DISTORTED keypoint (:826)
            float dv = -1.f, ur = -1.f;
            if (d > 0.f) {
                dv = d;
                ur = ux - bf / d;
            }
            depth_out[(size_t)b * cap + i] = dv;
            uright[(size_t)b * cap + i] = ur;
            const int px = (int)roundf((ux - gb.minx) * gb.inv_w), py = (int)roundf((uy - gb.miny) * gb.inv_h);
            if (px >= 0 && px < 64 && py >= 0 && py < 48) key = ((unsigned)(px * 48 + py) << 12) | (unsigned)i;
        }
        keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= npad; k <<= 1)  // bitonic sort: (cell, index) ascending = mGrid push_back order
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < npad; i += blockDim.x) {
                const int p = i ^ j;
                if (p > i) {
                    const unsigned a = keys[i], c = keys[p];
                    const bool up = (i & k) == 0;
                    if ((a > c) == up) {
                        keys[i] = c;
                        keys[p] = a;
                    }
                }
            }
            __syncthreads();
        }
    int* cs = cell_start + (size_t)b * (SG_CELLS + 1);
    int* ci = cell_items + (size_t)b * cap;
    for (int i = threadIdx.x; i < npad; i += blockDim.x) {
        const unsigned key = keys[i];
        if (key != 0xFFFFFFFFu) ci[i] = (int)(key & 0xFFFu);
        // cell_start[c] = first sorted position whose cell >= c
        const int cell = key == 0xFFFFFFFFu ? SG_CELLS : (int)(key >> 12);
        const int prev = i == 0 ? -1 : (keys[i - 1] == 0xFFFFFFFFu ? SG_CELLS : (int)(keys[i - 1] >> 12));
        for (int c = prev + 1; c <= cell; ++c) cs[c] = i;
    }
    if (threadIdx.x == 0) {
        const int last = npad == 0 ? -1 : (keys[npad - 1] == 0xFFFFFFFFu ? SG_CELLS : (int)(keys[npad - 1] >> 12));
        int total = 0;
        for (int i = 0; i < npad; ++i) total += keys[i] != 0xFFFFFFFFu;  // n <= 4096: trivial
        for (int c = last + 1; c <= SG_CELLS; ++c) cs[c] = total;
    }
}

int launch_stereo_grid(const float* depth, size_t depth_stride_b, int w, int h, int batch, const gd_keypoint* kps, size_t cap,
                       const int* n_kp, float bf, const UndistortArgs& und, float* depth_out, float* uright, int* cell_start,
                       int* cell_items, float2* un_out, cudaStream_t s, LaunchStats* st)
{
    GD_REQUIRE(cap <= (size_t)SG_MAX, "keypoint capacity above 4096 is not supported by the grid kernel");
    LaunchScope ls(st, s, "F3_stereo_grid", 1);
    // Frame::ComputeImageBounds (Frame.cc:608-636): undistorted image corners when mDistCoef(0) != 0, else the image itself
    GridBounds gb;
    float minx = 0.f, maxx = (float)w, miny = 0.f, maxy = (float)h;
    gb.distorted = und.k[0] != 0.0 ? 1 : 0;
    if (gb.distorted) {
        float c[4][2];
        undistort_point_cv(und, 0.f, 0.f, &c[0][0], &c[0][1]);
        undistort_point_cv(und, (float)w, 0.f, &c[1][0], &c[1][1]);
        undistort_point_cv(und, 0.f, (float)h, &c[2][0], &c[2][1]);
        undistort_point_cv(und, (float)w, (float)h, &c[3][0], &c[3][1]);
        minx = std::min(c[0][0], c[2][0]);
        maxx = std::max(c[1][0], c[3][0]);
        miny = std::min(c[0][1], c[1][1]);
        maxy = std::max(c[2][1], c[3][1]);
    }
    gb.minx = minx;
    gb.miny = miny;
    gb.inv_w = 64.f / (maxx - minx);
    gb.inv_h = 48.f / (maxy - miny);
    k_stereo_grid<<<batch, 512, 0, s>>>(depth, depth_stride_b, w, h, kps, cap, n_kp, bf, und, gb, depth_out, uright, cell_start, cell_items,
                                        un_out);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

int launch_erode_filter(const uint8_t* mask, size_t mask_stride_b, int w, int h, int batch, const gd_keypoint* kps, size_t cap,
                        const int* n_kp, int n_fixed, uint8_t* keep, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "F2_erode_filter", 1);
    EllipseSE se;
    make_ellipse31(&se);
    dim3 grid(cdiv((int)cap, 8), batch);
    k_erode_filter<<<grid, 256, 0, s>>>(mask, mask_stride_b, w, h, kps, cap, n_kp, n_fixed, se, keep, cap);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

int launch_compact_keypoints(const gd_keypoint* kps, const uint8_t* desc, const uint8_t* keep, size_t cap, int batch,
                             const int* n_kp, gd_keypoint* out_kps, uint8_t* out_desc, int* out_n, cudaStream_t s, LaunchStats* st)
{
    LaunchScope ls(st, s, "F2_compact_keypoints", 1);
    k_compact_keypoints<<<batch, 256, 0, s>>>(kps, desc, keep, cap, n_kp, out_kps, out_desc, out_n);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

}  // namespace gd
