// orb.cu — ORBextractor for sm_100a (built with --fmad=false: integer stages and individually rounded f32).
//
// Replaces ORB_SLAM2::ORBextractor::operator() (GD-SLAM src/ORBextractor.cc:1043-1105), bit-exact:
//   K4a k_orb_resize     ComputePyramid               :1107-1132  (cv::resize 8U INTER_LINEAR, 11-bit fixed point)
//   K4b k_orb_fast       cell loop + cv::FAST          :789-829    (FAST-9/16 score, cell-local NMS, threshold fallback,
//                                                                   raster-ordered compaction per cell)
//   K4c k_orb_quadtree   DistributeOctTree/DivideNode  :539-763, :481-537  (one CTA per level and stream; the std::list
//                                                                   algorithm restated with level-synchronous passes,
//                                                                   prefix sums and ranking — see the kernel comment)
//   K4e k_orb_blur       GaussianBlur 7x7 sigma 2      :1085-1086  (integer taps, single rounding)
//   K4d/e k_orb_describe IC_Angle + computeOrbDescriptor :77-147, fix-up :837-847, scaling :1095-1101
#include "orb.cuh"

#include <climits>
#include <cmath>

namespace gd {

// global (not __constant__): every lane reads a different 32-byte slice, which the constant cache would serialise
__device__ __align__(16) signed char d_pattern[1024] = {
#include "orb_pattern.inc"
};

// ------------------------------------------------------------------------------------------------ plan (host)
static int cv_round_f(float v) { return (int)lrintf(v); }
static int cv_round_dd(double v) { return (int)lrint(v); }
static int cv_floor_d(double v)
{
    int i = (int)v;
    return i - (i > v);
}
static int cv_ceil_d(double v)
{
    int i = (int)v;
    return i + (i < v);
}

int orb_make_plan(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int w, int h, OrbPlan* p)
{
    GD_REQUIRE(nlevels >= 1 && nlevels <= ORB_MAX_LEVELS, "nlevels out of range");
    GD_REQUIRE(nfeatures >= 1 && scale_factor > 1.0f, "bad nfeatures / scale factor");
    *p = OrbPlan();
    p->nlevels = nlevels;
    p->nfeatures = nfeatures;
    p->iniTh = ini_th;
    p->minTh = min_th;
    p->w = w;
    p->h = h;
    // ORBextractor::ORBextractor, :410-446 (scaleFactor is a double member initialised from the float argument)
    const double scaleFactor = (double)scale_factor;
    float sc[ORB_MAX_LEVELS], inv[ORB_MAX_LEVELS];
    sc[0] = 1.0f;
    for (int i = 1; i < nlevels; i++) sc[i] = (float)(sc[i - 1] * scaleFactor);
    for (int i = 0; i < nlevels; i++) inv[i] = 1.0f / sc[i];
    int nper[ORB_MAX_LEVELS];
    const float factor = (float)(1.0f / scaleFactor);
    float nDesired = (float)(nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels)));
    int sum = 0;
    for (int l = 0; l < nlevels - 1; l++) {
        nper[l] = cv_round_f(nDesired);
        sum += nper[l];
        nDesired *= factor;
    }
    nper[nlevels - 1] = std::max(nfeatures - sum, 0);
    // umax, :452-469
    {
        int v, v0, vmax = cv_floor_d(ORB_HALF * std::sqrt(2.f) / 2 + 1);
        int vmin = cv_ceil_d(ORB_HALF * std::sqrt(2.f) / 2);
        const double hp2 = ORB_HALF * ORB_HALF;
        for (v = 0; v <= vmax; ++v) p->umax[v] = cv_round_dd(std::sqrt(hp2 - v * v));
        for (v = ORB_HALF, v0 = 0; v >= vmin; --v) {
            while (p->umax[v0] == p->umax[v0 + 1]) ++v0;
            p->umax[v] = v0;
            ++v0;
        }
    }
    size_t off = 0;
    int cells = 0, cand = 0, kept = 0;
    for (int l = 0; l < nlevels; ++l) {
        OrbLevel& L = p->lv[l];
        L.w = cv_round_f((float)w * inv[l]);
        L.h = cv_round_f((float)h * inv[l]);
        GD_REQUIRE(L.w >= 2 * ORB_BORDER + 30 && L.h >= 2 * ORB_BORDER + 30, "image too small for this pyramid depth");
        L.pitch = (int)align_up((size_t)L.w, 64);
        L.off = off;
        off += align_up((size_t)L.pitch * L.h, 256);
        L.scale = sc[l];
        L.kp_size = (float)(int)(31 * sc[l]);
        L.N = nper[l];
        if (l > 0) {
            L.scale_x = (double)p->lv[l - 1].w / L.w;
            L.scale_y = (double)p->lv[l - 1].h / L.h;
        } else {
            L.scale_x = L.scale_y = 1.0;
        }
        // cell grid, :773-787
        const int minB = ORB_BORDER, maxBX = L.w - ORB_EDGE + 3, maxBY = L.h - ORB_EDGE + 3;
        const float width = (float)(maxBX - minB), height = (float)(maxBY - minB);
        L.nCols = (int)(width / 30.f);
        L.nRows = (int)(height / 30.f);
        GD_REQUIRE(L.nCols >= 1 && L.nRows >= 1, "level too small for one FAST cell");
        L.wCell = (int)std::ceil(width / L.nCols);
        L.hCell = (int)std::ceil(height / L.nRows);
        GD_REQUIRE((long long)L.nCols * L.nRows * L.nCols < (1ll << 20), "too many FAST cells per level for the index split");
        L.cell_start = cells;
        cells += L.nCols * L.nRows;
        p->max_cells_level = std::max(p->max_cells_level, L.nCols * L.nRows);
        p->tile_w = std::max(p->tile_w, (int)align_up((size_t)L.wCell + 6 + 3, 4));  // + alignment columns, pitch % 4 == 0
        p->tile_h = std::max(p->tile_h, L.hCell + 6);
        // DistributeOctTree initial nodes, :543-545
        L.nIni = (int)std::round((float)(maxBX - minB) / (maxBY - minB));
        GD_REQUIRE(L.nIni >= 1, "aspect ratio not supported by DistributeOctTree (nIni = 0)");
        L.hX = (float)(maxBX - minB) / L.nIni;
        L.cand_off = cand;
        // non-maximum suppression runs per cell: no two 8-adjacent survivors INSIDE a cell, but survivors of neighbouring
        // cells can touch across the cell border -> the bound is the sum of the per-cell bounds
        L.cand_cap = L.nCols * L.nRows * ((L.wCell + 1) / 2) * ((L.hCell + 1) / 2);
        cand += (int)align_up((size_t)L.cand_cap, 64);
        L.kept_off = kept;
        kept += L.N + 8;
        p->max_N = std::max(p->max_N, L.N);
        p->kp_capacity += L.N + 3;
    }
    p->pyr_bytes = off;
    p->total_cells = cells;
    p->cell_cap = ((p->tile_w - 6 + 1) / 2) * ((p->tile_h - 6 + 1) / 2);
    GD_REQUIRE(p->tile_w - 9 <= 64 && p->tile_h - 6 <= 64, "FAST cell larger than the keep bitmap (64 x 64)");
    p->cand_total = cand;
    p->kept_total = kept;
    GD_REQUIRE(p->max_N + 8 < 30000 && p->tile_w * p->tile_h * 2 < 200 * 1024, "plan exceeds kernel limits");
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------ K4a resize
// The source index and the two 11-bit weights of cv::resize depend on one coordinate only: they are tabulated per level
// on the host (ushort4 = {i0, i1, w0, w1}; same f64 -> f32 -> round-half-even arithmetic as the per-pixel form) so the
// kernel is four byte loads and the fixed-point blend per pixel.
constexpr int RS_ROWS = 4;
__global__ void __launch_bounds__(256) k_orb_resize(const uint8_t* __restrict__ src, int spitch, uint8_t* __restrict__ dst, int dw,
                                                    int dh, int dpitch, size_t stride_b, const ushort4* __restrict__ xt,
                                                    const ushort4* __restrict__ yt)
{
    pdl_wait();
    const int dx = blockIdx.x * 32 + threadIdx.x;
    if (dx >= dw) return;
    const uint8_t* s = src + (size_t)blockIdx.z * stride_b;
    uint8_t* d = dst + (size_t)blockIdx.z * stride_b;
    const ushort4 X = __ldg(xt + dx);
    const int a0 = X.z, a1 = X.w;
#pragma unroll
    for (int j = 0; j < RS_ROWS; ++j) {
        const int dy = (blockIdx.y * RS_ROWS + j) * 8 + threadIdx.y;
        if (dy >= dh) break;
        const ushort4 Y = __ldg(yt + dy);
        const uint8_t* r0 = s + (size_t)Y.x * spitch;
        const uint8_t* r1 = s + (size_t)Y.y * spitch;
        const int h0 = __ldg(r0 + X.x) * a0 + __ldg(r0 + X.y) * a1;
        const int h1 = __ldg(r1 + X.x) * a0 + __ldg(r1 + X.y) * a1;
        const int v = ((((int)Y.z * (h0 >> 4)) >> 16) + (((int)Y.w * (h1 >> 4)) >> 16) + 2) >> 2;
        d[(size_t)dy * dpitch + dx] = (uint8_t)max(0, min(255, v));
    }
}

// host side of the tables (cv::resize INTER_LINEAR 8U: coordinate in f64, cast to f32, weights rounded to 11 bits)
static void orb_resize_axis_table(int dn, int sn, double scale, bool clamp_like_x, ushort4* t)
{
    for (int dpos = 0; dpos < dn; ++dpos) {
        float f = (float)((dpos + 0.5) * scale - 0.5);
        int si = (int)std::floor(f);
        f -= si;
        int i0, i1;
        if (clamp_like_x) {
            if (si < 0) { f = 0; si = 0; }
            if (si >= sn - 1) { f = 0; si = sn - 1; }
            i0 = si;
            i1 = si < sn - 1 ? si + 1 : si;
        } else {
            i0 = std::max(0, std::min(sn - 1, si));
            i1 = std::max(0, std::min(sn - 1, si + 1));
        }
        const int w0 = (int)std::nearbyintf((1.f - f) * 2048), w1 = (int)std::nearbyintf(f * 2048);
        t[dpos] = make_ushort4((unsigned short)i0, (unsigned short)i1, (unsigned short)w0, (unsigned short)w1);
    }
}

// ------------------------------------------------------------------------------------------------ K4b FAST
struct FastLevelDev {
    int w, h, pitch;
    unsigned long long off;
    int nCols, nRows, wCell, hCell, cell_start;
    unsigned inv_nCols;  // (1 << 20) / nCols + 1: (ci * inv) >> 20 == ci / nCols for ci * nCols < 2^20
};
struct FastArgs {
    int nlevels, total_cells, cell_cap, tile_w, tile_h, iniTh, minTh;
    FastLevelDev lv[ORB_MAX_LEVELS];
};

// NOTE (toolchain hazard, CUDA 12.9 / sm_100a): `max(best, max(mn, -mx))` is folded by ptxas into VIMNMX3 with the
// negation DROPPED (verified on B200 with a 20-line repro: device returns max|d| where the host returns the FAST score).
// The dark-arc score is therefore computed as a min over the negated differences (ring - v) — no negated min/max operand.
// necessary condition for S' > th: a 9-arc contains at least two of the four compass points
// Packed form (see fast_full for the offset packing): half = d + 256 + (0x2000 - (th + 257)) has bit 13 set exactly when
// d > th (halves stay inside 0..0x3fff for th <= 255), the four masked flags add up without carries, and "at least two" is
// bit 14 or 15 of a half.
__device__ __forceinline__ bool fast_quick(const uint8_t* p, int tp, int th)
{
    th = max(-256, min(th, 255));  // outside this range the outcome no longer depends on th
    const int v = p[0];
    const unsigned r0 = p[3 * tp], r4 = p[3], r8 = p[-3 * tp], r12 = p[-3];
    unsigned C = (unsigned)(v + 0x2000 - 1 - th) + ((unsigned)(0x2000 - 1 - th - v) << 16);
    asm("" : "+r"(C));
    const unsigned M = 0x20002000u;
    const unsigned s = ((r0 * 65535u + C) & M) + ((r4 * 65535u + C) & M) + ((r8 * 65535u + C) & M) + ((r12 * 65535u + C) & M);
    return (s & 0xC000C000u) != 0u;
}

__device__ __forceinline__ int fast_full(const uint8_t* p, int tp)
{
    const int v = p[0];
    int r[16];  // ring pixels, clockwise from (0,+3)
    r[0] = p[3 * tp];
    r[4] = p[3];
    r[8] = p[-3 * tp];
    r[12] = p[-3];
    r[1] = p[3 * tp + 1];
    r[2] = p[2 * tp + 2];
    r[3] = p[tp + 3];
    r[5] = p[-tp + 3];
    r[6] = p[-2 * tp + 2];
    r[7] = p[-3 * tp + 1];
    r[9] = p[-3 * tp - 1];
    r[10] = p[-2 * tp - 2];
    r[11] = p[-tp - 3];
    r[13] = p[tp - 3];
    r[14] = p[2 * tp - 2];
    r[15] = p[3 * tp - 1];
    // Both polarities at once with packed 16-bit SIMD min/max (VIMNMX3.S16x2), on OFFSET differences so that one multiply-add
    // packs a ring pixel: low half = (v - r) + 256 (centre brighter), high half = (r - v) + 256 (centre darker), both in
    // 1..511;  r * 65535 + C = (r + 256 - v) * 65536 + (v + 256 - r)  with  C = (v + 256) + (256 - v) * 65536  (the low half
    // is positive, so nothing borrows from the high half).  min / max commute with the offset.
    unsigned C = (unsigned)(v + 256) + ((unsigned)(256 - v) << 16);
    asm("" : "+r"(C));  // keep C as one register: the compiler otherwise rewrites r * 65535 + C as (r - v) * 65535 + const
    unsigned q[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) q[k] = (unsigned)r[k] * 65535u + C;
    // min over every 9-arc: q3[k] = min(q[k..k+2]), arc(k) = min(q3[k], q3[k+3], q3[k+6]); S' = max over the 16 arcs
    unsigned q3[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) q3[k] = __vminu2(__vminu2(q[k], q[(k + 1) & 15]), q[(k + 2) & 15]);
    unsigned a9[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a9[k] = __vminu2(__vminu2(q3[k], q3[(k + 3) & 15]), q3[(k + 6) & 15]);
    unsigned m5[5];
#pragma unroll
    for (int k = 0; k < 5; ++k) m5[k] = __vmaxu2(__vmaxu2(a9[3 * k], a9[3 * k + 1]), a9[3 * k + 2]);
    const unsigned bestp = __vmaxu2(__vmaxu2(__vmaxu2(m5[0], m5[1]), m5[2]), __vmaxu2(__vmaxu2(m5[3], m5[4]), a9[15]));
    const int blo = (int)(bestp & 0xffffu), bhi = (int)(bestp >> 16);
    return max(blo, bhi) - 256;  // S'
}

constexpr int FAST_THREADS = 256;
constexpr int FAST_KEEP_WORDS = 128;  // keep bitmap: 64 bits per interior row of a cell (cells are < 60 x 60)
constexpr int FAST_KW = FAST_KEEP_WORDS / 32;

// All loops run over (row, column) with a warp per row, list entries are (y << 6 | x) and the keep bitmap is row aligned
// (bit y * 64 + x): no index divisions anywhere.  ncu of the first form of this kernel: 23 % of the instructions were the
// raster-ordered output replicated in all eight warps, 25 % the 4-point pass, 16 % the tile load, 8 % the prologue.
__global__ void __launch_bounds__(FAST_THREADS) k_orb_fast(const uint8_t* __restrict__ pyr, size_t pyr_stride_b, FastArgs a,
                                                          int* __restrict__ cell_cnt, ushort4* __restrict__ slabs)
{
    pdl_wait();
    extern __shared__ __align__(16) unsigned char sm[];
    const int tp = a.tile_w;                       // tile pitch (multiple of 4, >= widest cell + 3 alignment columns)
    uint8_t* tile_base = sm;                       // [tile_h][tp], rows start at the 4-byte aligned column below x0
    uint8_t* sc_base = sm + tp * a.tile_h;         // S' of the pixels that pass the 4-point test and the threshold, else 0
    __shared__ unsigned s_keep[FAST_KEEP_WORDS];
    const int b = blockIdx.y;
    const int cell = blockIdx.x;
    // level of the cell: binary search over the ascending cell_start table (unused levels hold INT_MAX)
    static_assert(ORB_MAX_LEVELS == 16, "four search steps");
    int l = cell >= a.lv[8].cell_start ? 8 : 0;
    l += cell >= a.lv[l + 4].cell_start ? 4 : 0;
    l += cell >= a.lv[l + 2].cell_start ? 2 : 0;
    l += cell >= a.lv[l + 1].cell_start ? 1 : 0;
    const FastLevelDev L = a.lv[l];  // by value: the level is a run-time index, every later use would be an indexed constant read
    const int ci = cell - L.cell_start;
    const int i = (int)(((unsigned)ci * L.inv_nCols) >> 20), j = ci - i * L.nCols;
    const int maxBX = L.w - ORB_EDGE + 3, maxBY = L.h - ORB_EDGE + 3;
    const int x0 = ORB_BORDER + j * L.wCell, y0 = ORB_BORDER + i * L.hCell;
    int x1 = x0 + L.wCell + 6, y1 = y0 + L.hCell + 6;
    int* cnt_out = cell_cnt + (size_t)b * a.total_cells + cell;
    if (y0 >= maxBY - 3 || x0 >= maxBX - 6) {  // :794, :803
        if (threadIdx.x == 0) *cnt_out = 0;
        return;
    }
    x1 = min(x1, maxBX);
    y1 = min(y1, maxBY);
    const int cw = x1 - x0, ch = y1 - y0;
    const int iw = cw - 6, ih = ch - 6;
    if (iw <= 0 || ih <= 0) {
        if (threadIdx.x == 0) *cnt_out = 0;
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int NW = FAST_THREADS / 32;
    const int xo = x0 & 3;                         // the tile holds image columns [x0 - xo, x1)
    const int nw = (cw + xo + 3) >> 2;             // 32-bit words per tile row (<= 17)
    const uint8_t* img = pyr + (size_t)b * pyr_stride_b + L.off + (size_t)y0 * L.pitch + (x0 - xo);
    // tile rows: half a warp per row, one word per lane (+ word 16 for the widest cells); the loads of three rounds of rows
    // are issued before the first store; score rows cleared alongside
    {
        const int k = lane & 15, yb = 2 * warp + (lane >> 4);
        const bool k1 = k + 16 < nw;  // nw <= 17
        auto ld = [&](int y, int kk) { return __ldg(reinterpret_cast<const unsigned*>(img + (size_t)y * L.pitch) + kk); };
        auto st = [&](int y, int kk, unsigned v) {
            reinterpret_cast<unsigned*>(tile_base + y * tp)[kk] = v;
            reinterpret_cast<unsigned*>(sc_base + y * tp)[kk] = 0u;
        };
        unsigned v0[3], v1[3];
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int y = yb + 2 * NW * it;
            v0[it] = (y < ch && k < nw) ? ld(y, k) : 0u;
            v1[it] = (y < ch && k1) ? ld(y, k + 16) : 0u;
        }
#pragma unroll
        for (int it = 0; it < 3; ++it) {
            const int y = yb + 2 * NW * it;
            if (y < ch && k < nw) st(y, k, v0[it]);
            if (y < ch && k1) st(y, k + 16, v1[it]);
        }
        for (int y = yb + 6 * NW; y < ch; y += 2 * NW) {
            if (k < nw) st(y, k, ld(y, k));
            if (k1) st(y, k + 16, ld(y, k + 16));
        }
    }
    if (threadIdx.x < FAST_KEEP_WORDS) s_keep[threadIdx.x] = 0u;
    static_assert(FAST_KEEP_WORDS <= FAST_THREADS, "bitmap is cleared by one pass of the block");
    const uint8_t* tile = tile_base + xo;
    uint8_t* sc = sc_base + xo;
    // survivors of the 4-point test as (y << 6 | x), one private list segment per warp (a warp owns the rows y = warp mod 8:
    // no atomics, and the 16-point pass of a warp only needs its own list).  A segment holds ceil(ih / 8) * iw entries at most,
    // which is below tp * tile_h / 8 for every cell.
    unsigned short* plist = reinterpret_cast<unsigned short*>(sm + 2 * tp * a.tile_h) + warp * ((tp * a.tile_h) / NW);
    int tot = 0;
    unsigned kw[FAST_KW];  // keep words lane + 32 r (every warp holds the whole bitmap)
    // first pass at iniThFAST only, like the reference's first cv::FAST call; if no corner survives the NMS the cell is
    // redone at minThFAST (:812-816).  Per pass:
    //   1. 4-point necessary test for every pixel of the warp's rows, survivors appended to the warp's list
    //   2. the 16-point network over the list (densely packed lanes whatever the pass rate of the 4-point test)
    //   3. cell-local strict NMS over the listed corners only -> one keep bit per interior pixel
    //   4. raster-ordered compaction straight from the bitmap (one warp)
    for (int pass = 0; pass < 2 && tot == 0; ++pass) {
        const int th = pass == 0 ? a.iniTh : a.minTh;
        __syncthreads();  // tile / cleared scores visible (first pass); the previous pass is done with the bitmap
        int nl = 0;       // warp-uniform
        for (int y = warp; y < ih; y += NW) {
            const uint8_t* row = tile + (y + 3) * tp + 3;
            for (int xb = 0; xb < iw; xb += 32) {
                const int x = xb + lane;
                const bool qk = x < iw && fast_quick(row + x, tp, th);
                const unsigned bal = __ballot_sync(0xffffffffu, qk);
                if (qk) plist[nl + __popc(bal & ((1u << lane) - 1))] = (unsigned short)((y << 6) | x);
                nl += __popc(bal);
            }
        }
        __syncwarp();
        for (int q = lane; q < nl; q += 32) {
            const int t = plist[q];
            const int pp = ((t >> 6) + 3) * tp + (t & 63) + 3;
            const int sv = fast_full(tile + pp, tp);
            sc[pp] = (uint8_t)(sv > th ? sv : 0);
        }
        __syncthreads();
        // NMS: keep iff S' > th and S' > S'_nb for every neighbour that is itself a corner at th (sc is 0 otherwise)
        for (int q = lane; q < nl; q += 32) {
            const int t = plist[q];
            const uint8_t* c = sc + ((t >> 6) + 3) * tp + (t & 63) + 3;
            const int sv = c[0];
            if (sv == 0) continue;
            const int m0 = max(max(c[-tp - 1], c[-tp]), max(c[-tp + 1], c[-1]));
            const int m1 = max(max(c[tp - 1], c[tp]), max(c[tp + 1], c[1]));
            if (max(m0, m1) < sv) atomicOr(&s_keep[t >> 5], 1u << (t & 31));
        }
        __syncthreads();
        int cnt = 0;
#pragma unroll
        for (int r = 0; r < FAST_KW; ++r) {
            kw[r] = s_keep[lane + 32 * r];
            cnt += __popc(kw[r]);
        }
        tot = __reduce_add_sync(0xffffffffu, cnt);
    }
    if (warp != 0) return;
    // raster order = ascending bit index.  Exclusive prefix of the per-word counts (word k lives in lane k & 31, register
    // k >> 5), then every lane writes the corners of its own words.
    ushort4* slab = slabs + ((size_t)b * a.total_cells + cell) * a.cell_cap;
    int run = 0;
#pragma unroll
    for (int r = 0; r < FAST_KW; ++r) {
        if (16 * r >= ih) break;  // words 32 r .. 32 r + 31 hold rows 16 r .. 16 r + 15
        unsigned word = kw[r];
        const int c = __popc(word);
        int inc = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, inc, d);
            if (lane >= d) inc += v;
        }
        int my = run + inc - c;
        run += __shfl_sync(0xffffffffu, inc, 31);
        const int tbase = (lane + 32 * r) * 32;
        while (word) {
            const int t = tbase + __ffs(word) - 1;
            word &= word - 1;
            if (my < a.cell_cap) {
                const int y = t >> 6, x = t & 63;
                slab[my] = make_ushort4((unsigned short)(x + 3 + j * L.wCell), (unsigned short)(y + 3 + i * L.hCell),
                                        (unsigned short)(sc[(y + 3) * tp + x + 3] - 1), 0);
            }
            ++my;
        }
    }
    if (lane == 0) *cnt_out = min(tot, a.cell_cap);
}

// ------------------------------------------------------------------------------------------------ whole-level FAST (GetRt)
// cv::FAST(threshold, nonmaxSuppression = true) on a whole image, as cv::ORB runs it per pyramid level (GetRt):
// (1) score map S' (0 unless S' > th) on regular 64 x 16 tiles, 4-point test -> compacted list -> 16-point network;
// (2) strict 8-neighbour NMS over the map (pixels closer than 3 to the border have no score).  The second kernel writes a
// map that holds S' at the surviving corners and 0 elsewhere; raster order of its non-zeros = cv::FAST's output order.
constexpr int FW_W = 64, FW_H = 16, FW_P = 72;

__global__ void __launch_bounds__(256) k_fast_whole_score(const uint8_t* __restrict__ img, int w, int h, int pitch, int th,
                                                          uint8_t* __restrict__ score)
{
    __shared__ __align__(16) uint8_t tile[(FW_H + 6) * FW_P];
    __shared__ __align__(16) uint8_t sc[FW_H * FW_W];
    __shared__ unsigned short plist[FW_H * FW_W];
    __shared__ int s_nlist;
    const int X0 = FW_W * blockIdx.x, Y0 = FW_H * blockIdx.y;  // first scored pixel of the tile
    const int tid = threadIdx.x;
    for (int q = tid; q < (FW_H + 6) * FW_P; q += 256) {  // rows Y0 - 3 .. Y0 + 18, columns X0 - 4 .. X0 + 67
        const int row = q / FW_P, col = q - row * FW_P;
        const int gy = Y0 - 3 + row, gx = X0 - 4 + col;
        tile[q] = (gy >= 0 && gy < h && gx >= 0 && gx < w) ? img[(size_t)gy * pitch + gx] : (uint8_t)0;
    }
    reinterpret_cast<unsigned*>(sc)[tid] = 0u;
    if (tid == 0) s_nlist = 0;
    __syncthreads();
    const int lane = tid & 31;
#pragma unroll
    for (int base = 0; base < FW_H * FW_W; base += 256) {
        const int q = base + tid, ly = q >> 6, lx = q & 63;
        const int gx = X0 + lx, gy = Y0 + ly;
        bool qk = gx >= 3 && gx < w - 3 && gy >= 3 && gy < h - 3;
        if (qk) qk = fast_quick(tile + (ly + 3) * FW_P + lx + 4, FW_P, th);
        const unsigned bal = __ballot_sync(0xffffffffu, qk);
        int wbase = 0;
        if (lane == 0 && bal) wbase = atomicAdd(&s_nlist, __popc(bal));
        wbase = __shfl_sync(0xffffffffu, wbase, 0);
        if (qk) plist[wbase + __popc(bal & ((1u << lane) - 1))] = (unsigned short)q;
    }
    __syncthreads();
    const int nl = s_nlist;
    for (int e = tid; e < nl; e += 256) {
        const int q = plist[e], ly = q >> 6, lx = q & 63;
        const int sv = fast_full(tile + (ly + 3) * FW_P + lx + 4, FW_P);
        sc[q] = (uint8_t)(sv > th ? sv : 0);
    }
    __syncthreads();
    for (int q = tid; q < FW_H * FW_W; q += 256) {
        const int ly = q >> 6, lx = q & 63, gx = X0 + lx, gy = Y0 + ly;
        if (gx < w && gy < h) score[(size_t)gy * pitch + gx] = sc[q];
    }
}

__global__ void __launch_bounds__(256) k_fast_whole_nms(const uint8_t* __restrict__ score, int w, int h, int pitch,
                                                        uint8_t* __restrict__ kept)
{
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= w || y >= h) return;
    const uint8_t* c = score + (size_t)y * pitch + x;
    const int s = c[0];
    int out = 0;
    if (s > 0) {  // scored pixels are at least 3 from the border: the 8 neighbours exist
        const int m0 = max(max(c[-pitch - 1], c[-pitch]), max(c[-pitch + 1], c[-1]));
        const int m1 = max(max(c[pitch - 1], c[pitch]), max(c[pitch + 1], c[1]));
        if (max(m0, m1) < s) out = s;
    }
    kept[(size_t)y * pitch + x] = (uint8_t)out;
}

int orb_fast_whole(const uint8_t* d_img, int w, int h, int pitch, int th, uint8_t* d_score, uint8_t* d_kept, cudaStream_t s)
{
    k_fast_whole_score<<<dim3(cdiv(w, FW_W), cdiv(h, FW_H)), 256, 0, s>>>(d_img, w, h, pitch, th, d_score);
    GD_CUDA(cudaGetLastError());
    k_fast_whole_nms<<<dim3(cdiv(w, 32), cdiv(h, 8)), dim3(32, 8), 0, s>>>(d_score, w, h, pitch, d_kept);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ---- the same over all pyramid levels of `batch` streams in ONE launch (resident GetRt stage).
// Levels are dense (pitch = width) at byte offset lv[l].off inside a stream's pyramid; blockIdx = (tile, level, stream).
// The NMS kernel also counts, per level row, the surviving corners inside the cv::ORB border (edge <= x < w - edge, same
// for y: KeyPointsFilter::runByImageBorder): the selection kernel turns the counts into raster-ordered list positions.
// One kernel: a CTA scores the 66 x 18 pixels around its 64 x 16 output tile (the 1-pixel ring of scores the 8-neighbour
// suppression needs is recomputed instead of being exchanged through a score map in HBM), suppresses and writes the map of
// kept corners.  The image tile is assembled from ALIGNED 32-bit words (dense levels: rows start at any byte) with a funnel
// shift; words that would lie outside the level are clamped to its first / last word — they only hold pixels outside the
// image, which no scored pixel (3 <= x < w - 3, 3 <= y < h - 3) ever reads.
constexpr int FK_TP = 80, FK_TR = FW_H + 8;   // tile: 24 rows x 80 bytes (19 words loaded), image (Y0 - 4 + r, X0 - 4 + c)
constexpr int FK_SW = FW_W + 2, FK_SH = FW_H + 2, FK_SP = 68;  // scores: 18 rows x 66 (pitch 68), image (Y0 - 1 + sy, X0 - 1 + sx)

__global__ void __launch_bounds__(256) k_cv_fast_kept_levels(const uint8_t* __restrict__ pyr, size_t stride_b, CvPyrArgs a, int th,
                                                             int edge, uint8_t* __restrict__ kept, int* __restrict__ rowcnt,
                                                             size_t rowcnt_stride)
{
    const CvLevelDev L = a.lv[blockIdx.y];
    const int tiles_x = (L.w + FW_W - 1) / FW_W, tiles_y = (L.h + FW_H - 1) / FW_H;
    if ((int)blockIdx.x >= tiles_x * tiles_y) return;
    const int w = L.w, h = L.h, pitch = L.w;
    const uint8_t* img = pyr + (size_t)blockIdx.z * stride_b + L.off;  // 256-byte aligned
    __shared__ __align__(16) uint8_t tile[FK_TR * FK_TP];
    __shared__ __align__(16) uint8_t sc[FK_SH * FK_SP];
    constexpr int FK_ITERS = (FK_SH * FK_SW + 255) / 256;        // 5 rounds of 256 positions
    __shared__ unsigned short plist[8][FK_ITERS * 32];           // one private survivor list per warp: no atomics
    const int ty = (int)blockIdx.x / tiles_x;
    const int X0 = FW_W * ((int)blockIdx.x - ty * tiles_x), Y0 = FW_H * ty;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        const int last = (w * h - 1) & ~3;
        const unsigned* iw = reinterpret_cast<const unsigned*>(img);
        for (int q = tid; q < FK_TR * 19; q += 256) {
            const int r = (q * 3450) >> 16, k = q - 19 * r;  // q / 19 for q < 4681
            const int gy = min(max(Y0 - 4 + r, 0), h - 1);
            const int o = gy * pitch + X0 - 4 + 4 * k, m = o & 3, oa = o - m;
            const unsigned w0 = __ldg(iw + (min(max(oa, 0), last) >> 2)), w1 = __ldg(iw + (min(max(oa + 4, 0), last) >> 2));
            reinterpret_cast<unsigned*>(tile + r * FK_TP)[k] = __funnelshift_r(w0, w1, 8 * m);
        }
        for (int q = tid; q < FK_SH * FK_SP / 4; q += 256) reinterpret_cast<unsigned*>(sc)[q] = 0u;
    }
    __syncthreads();
    int nl = 0;  // warp-uniform
    unsigned short* wl = plist[warp];
#pragma unroll
    for (int base = 0; base < FK_ITERS * 256; base += 256) {
        const int q = base + tid;
        const int sy = (q * 993) >> 16, sx = q - FK_SW * sy;  // q / 66 for q < 32768
        const int gx = X0 - 1 + sx, gy = Y0 - 1 + sy;
        bool qk = q < FK_SH * FK_SW && gx >= 3 && gx < w - 3 && gy >= 3 && gy < h - 3;
        if (qk) qk = fast_quick(tile + (sy + 3) * FK_TP + sx + 3, FK_TP, th);
        const unsigned bal = __ballot_sync(0xffffffffu, qk);
        if (qk) wl[nl + __popc(bal & ((1u << lane) - 1))] = (unsigned short)((sy << 7) | sx);
        nl += __popc(bal);
    }
    __syncwarp();
    for (int e = lane; e < nl; e += 32) {
        const int q = wl[e], sy = q >> 7, sx = q & 127;
        const int sv = fast_full(tile + (sy + 3) * FK_TP + sx + 3, FK_TP);
        sc[sy * FK_SP + sx] = (uint8_t)(sv > th ? sv : 0);
    }
    __syncthreads();
    uint8_t* kp = kept + (size_t)blockIdx.z * stride_b + L.off;
#pragma unroll
    for (int base = 0; base < FW_H * FW_W; base += 256) {
        const int q = base + tid, ly = q >> 6, lx = q & 63;
        const int x = X0 + lx, y = Y0 + ly;
        int out = 0;
        if (x < w && y < h) {
            const uint8_t* c = sc + (ly + 1) * FK_SP + lx + 1;
            const int sv = c[0];
            if (sv > 0) {
                const int m0 = max(max(c[-FK_SP - 1], c[-FK_SP]), max(c[-FK_SP + 1], c[-1]));
                const int m1 = max(max(c[FK_SP - 1], c[FK_SP]), max(c[FK_SP + 1], c[1]));
                if (max(m0, m1) < sv) out = sv;
            }
            kp[(size_t)y * pitch + x] = (uint8_t)out;
        }
        const bool inside = out > 0 && x >= edge && x < w - edge && y >= edge && y < h - edge;
        const unsigned bal = __ballot_sync(0xffffffffu, inside);  // one warp = 32 consecutive pixels of one row
        if (lane == 0 && bal) atomicAdd(rowcnt + (size_t)blockIdx.z * rowcnt_stride + L.row_off + y, __popc(bal));
    }
}

int orb_cv_fast_levels(const uint8_t* pyr, size_t stride_b, const CvPyrArgs& a, int batch, int th, int edge, uint8_t* score,
                       uint8_t* kept, int* rowcnt, size_t rowcnt_stride, cudaStream_t s)
{
    (void)score;  // the score map stays on chip
    int t_score = 0;
    for (int l = 0; l < a.nlevels; ++l) {
        t_score = std::max(t_score, cdiv(a.lv[l].w, FW_W) * cdiv(a.lv[l].h, FW_H));
        GD_REQUIRE(((size_t)a.lv[l].off & 3) == 0 && a.lv[l].w >= 8 && a.lv[l].h >= 8, "cv::ORB level layout");
    }
    GD_REQUIRE((stride_b & 3) == 0 && ((uintptr_t)pyr & 3) == 0, "cv::ORB pyramid alignment");
    k_cv_fast_kept_levels<<<dim3(t_score, a.nlevels, batch), 256, 0, s>>>(pyr, stride_b, a, th, edge, kept, rowcnt, rowcnt_stride);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// ------------------------------------------------------------------------------------------------ K4c quadtree
// DistributeOctTree restated for one CTA.  The reference keeps a std::list of nodes: every pass of its first loop
// splits ALL nodes holding more than one key (children are push_front'ed, the parent erased); once another full pass
// would overshoot N it switches to splitting one node at a time in order of (key count desc, newest first) until the
// list holds >= N nodes.  Observations that make it data parallel:
//   * list order after a pass = children of the last processed parent first (n4,n3,n2,n1), ..., then the untouched
//     single-key nodes in their old order  -> positions by prefix sums over the old list;
//   * "newest first" among nodes created in the same pass = smaller list position (SURVEY B-3 canonical rule);
//   * which child a key lands in depends only on the parent's bounds -> keys carry a node index, no key lists;
//   * the best key of a node = max response, ties to the earliest candidate (stable partitions keep candidate order).
struct QtLevelDev {
    int nCols, nRows, cell_start, N, nIni, cand_off, cand_cap, kept_off, maxX, maxY;
    float hX;
};
struct QtArgs {
    int nlevels, total_cells, cell_cap, cand_total, kept_total, LN, max_cells;
    QtLevelDev lv[ORB_MAX_LEVELS];
};

constexpr int QT_THREADS = 512;

// exclusive scan of data[0..n) by warp 0; all threads must call; result in place, returns the total
__device__ int block_excl_scan(int* data, int n, int* s_total)
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

__global__ void __launch_bounds__(QT_THREADS) k_orb_quadtree(QtArgs a, const int* __restrict__ cell_cnt,
                                                             const ushort4* __restrict__ slabs, ushort4* __restrict__ cand,
                                                             uint8_t* __restrict__ cand_q, int* __restrict__ kept,
                                                             int* __restrict__ kept_cnt, int* __restrict__ cand_cnt,
                                                             int* __restrict__ err)
{
    pdl_wait();
    extern __shared__ __align__(16) unsigned char qsm[];
    const int LN = a.LN;
    // shared arrays (ints unless noted)
    short4* ndA = reinterpret_cast<short4*>(qsm);             // node bounds (ulx, uly, brx, bry), ping
    short4* ndB = ndA + LN;                                     // pong
    int* szA = reinterpret_cast<int*>(ndB + LN);
    int* szB = szA + LN;
    int* cc = szB + LN;          // child key counts [LN*4]
    int* cpos = cc + 4 * LN;     // child new positions [LN*4]
    int* kpos = cpos + 4 * LN;   // new position of a kept node [LN]
    int* sA = kpos + LN;         // scan scratch A [max(LN, max_cells)]
    int* sB = sA + max(LN, a.max_cells);  // scan scratch B [LN]
    int* rk = sB + LN;           // rank of a candidate [LN]
    int* byrank = rk + LN;       // node index by rank [LN]
    unsigned char* candA = reinterpret_cast<unsigned char*>(byrank + LN);  // candidate flag ping [LN]
    unsigned char* candB = candA + LN;
    unsigned char* dv = candB + LN;  // divide flag [LN]
    unsigned long long* best = reinterpret_cast<unsigned long long*>(
        (reinterpret_cast<uintptr_t>(dv + LN) + 7) & ~(uintptr_t)7);  // [LN]
    __shared__ int s_total, s_flag, s_rstar;

    const int l = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    const QtLevelDev& L = a.lv[l];
    ushort4* cd = cand + (size_t)b * a.cand_total + L.cand_off;
    uint8_t* cq = cand_q + (size_t)b * a.cand_total + L.cand_off;
    int* kp = kept + (size_t)b * a.kept_total + L.kept_off;
    const int N = L.N;

    // ---- compact the per-cell lists (cells row-major, raster inside a cell = vToDistributeKeys order)
    const int ncell = L.nCols * L.nRows;
    const int* cc_in = cell_cnt + (size_t)b * a.total_cells + L.cell_start;
    for (int c = tid; c < ncell; c += QT_THREADS) sA[c] = cc_in[c];
    int n = block_excl_scan(sA, ncell, &s_total);
    if (n > L.cand_cap) {  // per-cell NMS density bound; flag instead of corrupting memory
        if (tid == 0) atomicOr(err, 1);
        n = L.cand_cap;
    }
    {
        const int warp = tid >> 5, lane = tid & 31;
        for (int c = warp; c < ncell; c += QT_THREADS / 32) {
            const int cnt = cc_in[c], o = sA[c];
            const ushort4* src = slabs + ((size_t)b * a.total_cells + L.cell_start + c) * a.cell_cap;
            for (int e = lane; e < cnt; e += 32)
                if (o + e < n) cd[o + e] = src[e];
        }
    }
    if (tid == 0) cand_cnt[b * a.nlevels + l] = n;
    __syncthreads();
    if (n == 0) {
        if (tid == 0) kept_cnt[b * a.nlevels + l] = 0;
        return;
    }
    // ---- initial nodes (:543-585)
    int cnt = L.nIni;
    for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) {
        ndA[i2] = make_short4((short)(int)(L.hX * (float)i2), 0, (short)(int)(L.hX * (float)(i2 + 1)), (short)L.maxY);
        szA[i2] = 0;
        candA[i2] = 0;
    }
    __syncthreads();
    for (int k = tid; k < n; k += QT_THREADS) {
        ushort4 c = cd[k];
        const int ni = (int)((float)c.x / L.hX);
        c.w = (unsigned short)ni;
        cd[k] = c;
        atomicAdd(&szA[ni], 1);
    }
    __syncthreads();
    {  // erase empty initial nodes, keep order
        for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) sB[i2] = szA[i2] > 0 ? 1 : 0;
        const int newcnt = block_excl_scan(sB, cnt, &s_total);
        if (newcnt != cnt) {
            for (int i2 = tid; i2 < cnt; i2 += QT_THREADS)
                if (szA[i2] > 0) {
                    ndB[sB[i2]] = ndA[i2];
                    szB[sB[i2]] = szA[i2];
                    candB[sB[i2]] = 0;
                }
            __syncthreads();
            for (int k = tid; k < n; k += QT_THREADS) {
                ushort4 c = cd[k];
                c.w = (unsigned short)sB[c.w];
                cd[k] = c;
            }
            __syncthreads();
            for (int i2 = tid; i2 < newcnt; i2 += QT_THREADS) {
                ndA[i2] = ndB[i2];
                szA[i2] = szB[i2];
                candA[i2] = 0;
            }
            cnt = newcnt;
            __syncthreads();
        }
    }
    short4* nd = ndA;
    short4* nd2 = ndB;
    int* sz = szA;
    int* sz2 = szB;
    unsigned char* cf = candA;
    unsigned char* cf2 = candB;

    // one splitting step: nodes with dv[i] set are divided; `order` gives the processing order of divided nodes:
    //   order == nullptr : list order (phase 1);  otherwise rank order (phase 2), byrank[r] = node, nproc = #processed
    auto split_step = [&](bool by_rank, int nproc) -> int {
        for (int i2 = tid; i2 < 4 * cnt; i2 += QT_THREADS) cc[i2] = 0;
        __syncthreads();
        for (int k = tid; k < n; k += QT_THREADS) {
            const ushort4 c = cd[k];
            const int ni = c.w;
            if (by_rank ? cf[ni] : dv[ni]) {  // phase 2 counts children of every candidate (needed to find the stop)
                const short4 bnd = nd[ni];
                const int mx = bnd.x + ((bnd.z - bnd.x + 1) >> 1), my = bnd.y + ((bnd.w - bnd.y + 1) >> 1);
                const int q = ((int)c.x < mx) ? (((int)c.y < my) ? 0 : 2) : (((int)c.y < my) ? 1 : 3);
                cq[k] = (uint8_t)q;
                atomicAdd(&cc[4 * ni + q], 1);
            }
        }
        __syncthreads();
        return 0;
    };
    auto build_lists = [&](bool by_rank, int nproc) -> int {
        // sA: per processing slot number of non-empty children (exclusive scanned); sB: keep flags (scanned)
        const int nslots = by_rank ? nproc : cnt;
        for (int s2 = tid; s2 < nslots; s2 += QT_THREADS) {
            const int ni = by_rank ? byrank[s2] : s2;
            int c = 0;
            if (dv[ni]) c = (cc[4 * ni] > 0) + (cc[4 * ni + 1] > 0) + (cc[4 * ni + 2] > 0) + (cc[4 * ni + 3] > 0);
            sA[s2] = c;
        }
        const int totalC = block_excl_scan(sA, nslots, &s_total);
        for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) sB[i2] = dv[i2] ? 0 : 1;
        const int totalK = block_excl_scan(sB, cnt, &s_total);
        const int newcnt = totalC + totalK;
        if (newcnt > LN) {  // cannot happen: the list never exceeds N + 3
            if (tid == 0) atomicOr(err, 2);
            return -1;
        }
        for (int s2 = tid; s2 < nslots; s2 += QT_THREADS) {
            const int ni = by_rank ? byrank[s2] : s2;
            if (!dv[ni]) continue;
            const short4 bnd = nd[ni];
            const int mx = bnd.x + ((bnd.z - bnd.x + 1) >> 1), my = bnd.y + ((bnd.w - bnd.y + 1) >> 1);
            int c = 0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = cc[4 * ni + q];
                if (k > 0) {
                    const int pos = totalC - 1 - (sA[s2] + c);
                    short4 cb;
                    cb.x = (q & 1) ? (short)mx : bnd.x;
                    cb.z = (q & 1) ? bnd.z : (short)mx;
                    cb.y = (q & 2) ? (short)my : bnd.y;
                    cb.w = (q & 2) ? bnd.w : (short)my;
                    nd2[pos] = cb;
                    sz2[pos] = k;
                    cf2[pos] = k > 1 ? 1 : 0;
                    cpos[4 * ni + q] = pos;
                    ++c;
                }
            }
        }
        for (int i2 = tid; i2 < cnt; i2 += QT_THREADS)
            if (!dv[i2]) {
                const int pos = totalC + sB[i2];
                kpos[i2] = pos;
                nd2[pos] = nd[i2];
                sz2[pos] = sz[i2];
                cf2[pos] = 0;
            }
        __syncthreads();
        for (int k = tid; k < n; k += QT_THREADS) {
            ushort4 c = cd[k];
            const int ni = c.w;
            c.w = (unsigned short)(dv[ni] ? cpos[4 * ni + cq[k]] : kpos[ni]);
            cd[k] = c;
        }
        __syncthreads();
        short4* t1 = nd; nd = nd2; nd2 = t1;
        int* t2 = sz; sz = sz2; sz2 = t2;
        unsigned char* t3 = cf; cf = cf2; cf2 = t3;
        return newcnt;
    };
    auto count_flag = [&](const unsigned char* f, int m) -> int {
        if (tid == 0) s_flag = 0;
        __syncthreads();
        int c = 0;
        for (int i2 = tid; i2 < m; i2 += QT_THREADS) c += f[i2] ? 1 : 0;
        if (c) atomicAdd(&s_flag, c);
        __syncthreads();
        const int r = s_flag;
        __syncthreads();
        return r;
    };

    bool finish = false;
    while (!finish) {
        const int prev = cnt;
        for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) dv[i2] = sz[i2] > 1 ? 1 : 0;
        __syncthreads();
        split_step(false, 0);
        const int newcnt = build_lists(false, 0);
        if (newcnt < 0) break;
        cnt = newcnt;
        const int nToExpand = count_flag(cf, cnt);
        if (cnt >= N || cnt == prev)
            finish = true;
        else if (cnt + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = cnt;
                // rank the candidates: key count descending, then list position ascending (newest first)
                for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) {
                    int r = -1;
                    if (cf[i2]) {
                        r = 0;
                        const int si = sz[i2];
                        for (int j2 = 0; j2 < cnt; ++j2)
                            if (cf[j2] && (sz[j2] > si || (sz[j2] == si && j2 < i2))) ++r;
                    }
                    rk[i2] = r;
                }
                __syncthreads();
                const int C = count_flag(cf, cnt);
                for (int i2 = tid; i2 < cnt; i2 += QT_THREADS)
                    if (rk[i2] >= 0) byrank[rk[i2]] = i2;
                __syncthreads();
                split_step(true, C);  // child counts of every candidate
                // gain in rank order -> first rank where the list reaches N
                for (int r = tid; r < C; r += QT_THREADS) {
                    const int ni = byrank[r];
                    sA[r] = (cc[4 * ni] > 0) + (cc[4 * ni + 1] > 0) + (cc[4 * ni + 2] > 0) + (cc[4 * ni + 3] > 0) - 1;
                }
                block_excl_scan(sA, C, &s_total);
                if (tid == 0) s_rstar = C - 1;
                __syncthreads();
                for (int r = tid; r < C; r += QT_THREADS) {
                    const int ni = byrank[r];
                    const int gain = (cc[4 * ni] > 0) + (cc[4 * ni + 1] > 0) + (cc[4 * ni + 2] > 0) + (cc[4 * ni + 3] > 0) - 1;
                    if (cnt + sA[r] + gain >= N) atomicMin(&s_rstar, r);
                }
                __syncthreads();
                const int nproc = s_rstar + 1;
                for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) dv[i2] = (rk[i2] >= 0 && rk[i2] < nproc) ? 1 : 0;
                __syncthreads();
                const int nc2 = build_lists(true, nproc);
                if (nc2 < 0) {
                    finish = true;
                    break;
                }
                cnt = nc2;
                if (cnt >= N || cnt == prev2) finish = true;
            }
        }
    }
    // ---- best key per node (:741-760): max response, first in candidate order
    for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) best[i2] = 0ull;
    __syncthreads();
    for (int k = tid; k < n; k += QT_THREADS) {
        const ushort4 c = cd[k];
        const unsigned long long key = ((unsigned long long)c.z << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)k);
        atomicMax(&best[c.w], key);
    }
    __syncthreads();
    for (int i2 = tid; i2 < cnt; i2 += QT_THREADS) kp[i2] = (int)(0xFFFFFFFFu - (unsigned)(best[i2] & 0xFFFFFFFFull));
    if (tid == 0) kept_cnt[b * a.nlevels + l] = cnt;
}

static size_t qt_smem_bytes(int LN, int max_cells)
{
    size_t s = 0;
    s += sizeof(short4) * 2 * LN;
    s += sizeof(int) * (2 * LN + 4 * LN + 4 * LN + LN + std::max(LN, max_cells) + LN + LN + LN);
    s += 3 * (size_t)LN + 8;
    s += sizeof(unsigned long long) * LN;
    return align_up(s, 16);
}

// ------------------------------------------------------------------------------------------------ K4e blur
struct BlurLevelDev {
    int w, h, pitch, tiles_x, tile_start;
    unsigned long long off;
};
struct BlurArgs {
    int nlevels, total_tiles;
    BlurLevelDev lv[ORB_MAX_LEVELS];
};
// One warp blurs a 120 x 36 block: lane i owns the 4-pixel group at x0 - 4 + 4 i (lanes 0 and 31 only supply halo words),
// marches down 36 + 6 rows with the 7-row window of horizontal sums in registers.  Horizontal pass on packed 16-bit
// pairs (sums <= 255 * 256 fit a lane), vertical pass in 32 bit, one rounding at the end: the integer arithmetic is
// order independent, so the result equals the row-then-column cv::GaussianBlur fixed-point path bit for bit.
constexpr int BT_W = 120, BT_H = 36;  // BT_H + 6 = 42 rows = six passes of the seven-row body

__device__ __forceinline__ int reflect101i(int p, int len)
{
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * len - 2 - p;
    return p;
}

__global__ void __launch_bounds__(256, 4) k_orb_blur(const uint8_t* __restrict__ pyr, uint8_t* __restrict__ out, size_t stride_b,
                                                  BlurArgs a)
{
    pdl_wait();
    const int lane = threadIdx.x & 31;
    int t = blockIdx.x * 8 + (threadIdx.x >> 5), l = 0;
    if (t >= a.total_tiles) return;  // warp-uniform
    while (l + 1 < a.nlevels && t >= a.lv[l + 1].tile_start) ++l;
    const BlurLevelDev& L = a.lv[l];
    t -= L.tile_start;
    const int ty = t / L.tiles_x, tx = t - ty * L.tiles_x;
    const int w = L.w, h = L.h, pitch = L.pitch;
    const int x = tx * BT_W - 4 + 4 * lane, y0 = ty * BT_H;
    const uint8_t* img = pyr + (size_t)blockIdx.y * stride_b + L.off;
    uint8_t* dst = out + (size_t)blockIdx.y * stride_b + L.off;
    const bool fast = x >= 0 && x + 3 < w;   // aligned word fully inside the row
    const bool need = x < w + 3;             // groups further right are never read
    int xr[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) xr[k] = need ? reflect101i(x + k, w) : 0;
    const bool owner = lane >= 1 && lane <= 30 && x < w;
    unsigned hw[7][4];
    // the rows are requested BLUR_PF iterations ahead (the kernel is latency bound: one dependent load per row otherwise)
    auto load_row = [&](int r) -> unsigned {
        const int yy = reflect101i(y0 + r - 3, h);
        const uint8_t* row = img + (size_t)yy * pitch;
        if (fast) return __ldg(reinterpret_cast<const unsigned*>(row + x));
        return (unsigned)__ldg(row + xr[0]) | ((unsigned)__ldg(row + xr[1]) << 8) | ((unsigned)__ldg(row + xr[2]) << 16) |
               ((unsigned)__ldg(row + xr[3]) << 24);
    };
    // Seven rows (the period of the window ring and of the prefetch ring) are unrolled, the six passes over them are a real
    // loop: the fully unrolled 38-row form spent more stall cycles on instruction fetch than on memory (ncu: no_instruction).
    constexpr int BLUR_PF = 7;
    static_assert((BT_H + 6) % 7 == 0, "whole passes of seven rows");
    unsigned Bq[BLUR_PF];
#pragma unroll
    for (int r = 0; r < BLUR_PF; ++r) Bq[r] = load_row(r);
#pragma unroll 1
    for (int rb = 0; rb < BT_H + 6; rb += 7) {
#pragma unroll
        for (int rr = 0; rr < 7; ++rr) {
            const int r = rb + rr;  // r % 7 == rr
            const unsigned B = Bq[rr];
            if (r + BLUR_PF < BT_H + 6) Bq[rr] = load_row(r + BLUR_PF);
            const unsigned A = __shfl_up_sync(0xffffffffu, B, 1), C = __shfl_down_sync(0xffffffffu, B, 1);
            // byte windows starting k pixels from the group start, split into 16-bit pairs P_k = (b[k], b[k+1])
            const unsigned wm3 = __funnelshift_r(A, B, 8), wm2 = __funnelshift_r(A, B, 16);
            const unsigned wp1 = __funnelshift_r(B, C, 8), wp2 = __funnelshift_r(B, C, 16), wp3 = __funnelshift_r(B, C, 24);
            const unsigned Pm3 = __byte_perm(wm3, 0u, 0x4140), Pm1 = __byte_perm(wm3, 0u, 0x4342);
            const unsigned Pm2 = __byte_perm(wm2, 0u, 0x4140), P0 = __byte_perm(wm2, 0u, 0x4342);
            const unsigned P1 = __byte_perm(wp1, 0u, 0x4140), P3 = __byte_perm(wp1, 0u, 0x4342);
            const unsigned P2 = __byte_perm(wp2, 0u, 0x4140), P4 = __byte_perm(wp2, 0u, 0x4342);
            const unsigned P5 = __byte_perm(wp3, 0u, 0x4342);
            const unsigned h01 = 18u * (Pm3 + P3) + 34u * (Pm2 + P2) + 48u * (Pm1 + P1) + 56u * P0;
            const unsigned h23 = 18u * (Pm1 + P5) + 34u * (P0 + P4) + 48u * (P1 + P3) + 56u * P2;
            unsigned* hn = hw[rr];
            hn[0] = h01 & 0xffffu; hn[1] = h01 >> 16; hn[2] = h23 & 0xffffu; hn[3] = h23 >> 16;
            if (r >= 6) {
                const int y = y0 + r - 6;
                unsigned o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const unsigned acc = 18u * (hw[(rr + 1) % 7][j] + hw[rr][j]) + 34u * (hw[(rr + 2) % 7][j] + hw[(rr + 6) % 7][j]) +
                                         48u * (hw[(rr + 3) % 7][j] + hw[(rr + 5) % 7][j]) + 56u * hw[(rr + 4) % 7][j];
                    o[j] = (acc + 32768u) >> 16;
                }
                if (owner && y < h) {
                    uint8_t* q = dst + (size_t)y * pitch + x;
                    if (x + 3 < w) {
                        *reinterpret_cast<unsigned*>(q) = o[0] | (o[1] << 8) | (o[2] << 16) | (o[3] << 24);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (x + j < w) q[j] = (uint8_t)o[j];
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ K4d/e describe
struct DescLevelDev {
    int pitch, cand_off, kept_off;
    unsigned long long off;
    float scale, kp_size;
};
struct DescArgs {
    int nlevels, cand_total, kept_total, kp_capacity;
    int umax[ORB_HALF + 1];
    DescLevelDev lv[ORB_MAX_LEVELS];
};

__device__ __forceinline__ float fast_atan2_dev(float y, float x)
{
    const float scale = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * scale, p3 = -0.3258083974640975f * scale, p5 = 0.1555786518463281f * scale,
                p7 = -0.04432655554792128f * scale;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + 2.2204460492503131e-16f);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + 2.2204460492503131e-16f);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

// one warp per output keypoint
__global__ void __launch_bounds__(256) k_orb_describe(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur,
                                                      size_t stride_b, DescArgs a, const ushort4* __restrict__ cand,
                                                      const int* __restrict__ kept, const int* __restrict__ kept_cnt,
                                                      gd_keypoint* __restrict__ out_kp, uint8_t* __restrict__ out_desc,
                                                      int* __restrict__ out_n)
{
    pdl_wait();
    const int b = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int l = -1, within = 0, total = 0;
    for (int q = 0; q < a.nlevels; ++q) {
        const int c = kept_cnt[b * a.nlevels + q];
        if (l < 0 && g < total + c) {
            l = q;
            within = g - total;
        }
        total += c;
    }
    if (g == 0 && lane == 0) out_n[b] = total;
    if (l < 0 || g >= a.kp_capacity) return;
    const DescLevelDev& L = a.lv[l];
    const int ci = kept[(size_t)b * a.kept_total + L.kept_off + within];
    const ushort4 c = cand[(size_t)b * a.cand_total + L.cand_off + ci];
    const int x = c.x + ORB_BORDER, y = c.y + ORB_BORDER;  // :843-844
    const uint8_t* center = pyr + (size_t)b * stride_b + L.off + (size_t)y * L.pitch + x;
    // IC_Angle: lane <-> u = lane - 15
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF;
    if (lane < 2 * ORB_HALF + 1) {
        m10 = u * center[u];
        const int au = abs(u);
#pragma unroll
        for (int v = 1; v <= ORB_HALF; ++v) {
            if (au <= a.umax[v]) {
                const int vp = center[u + v * L.pitch], vm = center[u - v * L.pitch];
                m01 += v * (vp - vm);
                m10 += u * (vp + vm);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    const float angle = fast_atan2_dev((float)m01, (float)m10);
    // steered BRIEF on the blurred level: lane -> descriptor byte
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ar = angle * factorPI;
    const float ca = (float)cos((double)ar), sa = (float)sin((double)ar);
    const uint8_t* bc = blur + (size_t)b * stride_b + L.off + (size_t)y * L.pitch + x;
    int val = 0;
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2);
    const int4 q1 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2 + 1);
    const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};  // one test = 4 signed bytes (x0,y0,x1,y1)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int wd = words[k];
        const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff),
                    x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
        const int t0 = bc[__float2int_rn(x0 * sa + y0 * ca) * L.pitch + __float2int_rn(x0 * ca - y0 * sa)];
        const int t1 = bc[__float2int_rn(x1 * sa + y1 * ca) * L.pitch + __float2int_rn(x1 * ca - y1 * sa)];
        val |= (t0 < t1) << k;
    }
    out_desc[((size_t)b * a.kp_capacity + g) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        gd_keypoint k;
        k.x = (float)x;
        k.y = (float)y;
        if (l != 0) {  // :1095-1101
            k.x = k.x * L.scale;
            k.y = k.y * L.scale;
        }
        k.size = L.kp_size;
        k.angle = angle;
        k.response = (float)c.z;
        k.octave = l;
        k.class_id = -1;
        out_kp[(size_t)b * a.kp_capacity + g] = k;
    }
}

// ------------------------------------------------------------------------------------------------ core
// IC_Angle + steered BRIEF of cv::ORB for a list of keypoints at integer level coordinates (one warp per keypoint; the same
// arithmetic as k_orb_describe, cv::ORB only differs in the image the tests read: its float-blurred level)
struct CvDescArgs {
    int umax[ORB_HALF + 1];
};
__global__ void __launch_bounds__(256) k_cvorb_describe(const uint8_t* __restrict__ img, const uint8_t* __restrict__ blur, int pitch,
                                                        const int* __restrict__ xs, const int* __restrict__ ys, int n, CvDescArgs a,
                                                        float* __restrict__ angle_out, uint8_t* __restrict__ desc_out)
{
    const int lane = threadIdx.x & 31;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (g >= n) return;
    const int x = xs[g], y = ys[g];
    const uint8_t* center = img + (size_t)y * pitch + x;
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF;
    if (lane < 2 * ORB_HALF + 1) {
        m10 = u * center[u];
        const int au = abs(u);
#pragma unroll
        for (int v = 1; v <= ORB_HALF; ++v) {
            if (au <= a.umax[v]) {
                const int vp = center[u + v * pitch], vm = center[u - v * pitch];
                m01 += v * (vp - vm);
                m10 += u * (vp + vm);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    const float angle = fast_atan2_dev((float)m01, (float)m10);
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ar = angle * factorPI;
    const float ca = (float)cos((double)ar), sa = (float)sin((double)ar);
    const uint8_t* bc = blur + (size_t)y * pitch + x;
    int val = 0;
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2);
    const int4 q1 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2 + 1);
    const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int wd = words[k];
        const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff),
                    x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
        const int t0 = bc[__float2int_rn(x0 * sa + y0 * ca) * pitch + __float2int_rn(x0 * ca - y0 * sa)];
        const int t1 = bc[__float2int_rn(x1 * sa + y1 * ca) * pitch + __float2int_rn(x1 * ca - y1 * sa)];
        val |= (t0 < t1) << k;
    }
    desc_out[(size_t)g * 32 + lane] = (uint8_t)val;
    if (lane == 0) angle_out[g] = angle;
}

int orb_cv_describe(const uint8_t* d_img, const uint8_t* d_blur, int pitch, const int* d_xs, const int* d_ys, int n, float* d_angle,
                    uint8_t* d_desc, cudaStream_t s)
{
    if (n <= 0) return GD_OK;
    static const int umax[ORB_HALF + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};  // ORBextractor.cc:456-469
    CvDescArgs a;
    for (int i = 0; i <= ORB_HALF; ++i) a.umax[i] = umax[i];
    k_cvorb_describe<<<cdiv(n, 8), 256, 0, s>>>(d_img, d_blur, pitch, d_xs, d_ys, n, a, d_angle, d_desc);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

// IC_Angle + steered BRIEF for the keypoints the GetRt selection kernel kept (all levels, all streams, one launch): record g
// of a stream = the g-th entry of the level-ordered concatenation of the per-level lists sel[level][0 .. sel_n[level]).
// sel entry: .x = y * w + x at the level, .y = bits of the Harris response.  Output = cv::KeyPoint records in cv::ORB's
// order (pt scaled to level 0, size 31 * scale, octave) and 32-byte descriptors.
__global__ void __launch_bounds__(256) k_cv_describe_sel(const uint8_t* __restrict__ pyr, const uint8_t* __restrict__ blur, size_t stride_b,
                                                         CvPyrArgs a, CvDescArgs da, const uint2* __restrict__ sel, int sel_cap,
                                                         const int* __restrict__ sel_n, gd_keypoint* __restrict__ out_kp,
                                                         uint8_t* __restrict__ out_desc, int* __restrict__ out_n, int feat_cap)
{
    const int lane = threadIdx.x & 31, b = blockIdx.y;
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    int l = -1, within = 0, total = 0;
    for (int q = 0; q < a.nlevels; ++q) {
        const int c = min(sel_n[b * a.nlevels + q], sel_cap);
        if (l < 0 && g < total + c) {
            l = q;
            within = g - total;
        }
        total += c;
    }
    if (g == 0 && lane == 0) out_n[b] = min(total, feat_cap);
    if (l < 0 || g >= feat_cap) return;
    const CvLevelDev L = a.lv[l];
    const uint2 e = sel[((size_t)b * a.nlevels + l) * sel_cap + within];
    const int y = (int)(e.x / (unsigned)L.w), x = (int)(e.x - (unsigned)y * L.w), pitch = L.w;
    const uint8_t* center = pyr + (size_t)b * stride_b + L.off + (size_t)y * pitch + x;
    int m01 = 0, m10 = 0;
    const int u = lane - ORB_HALF;
    if (lane < 2 * ORB_HALF + 1) {
        m10 = u * center[u];
        const int au = abs(u);
#pragma unroll
        for (int v = 1; v <= ORB_HALF; ++v) {
            if (au <= da.umax[v]) {
                const int vp = center[u + v * pitch], vm = center[u - v * pitch];
                m01 += v * (vp - vm);
                m10 += u * (vp + vm);
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        m10 += __shfl_xor_sync(0xffffffffu, m10, o);
    }
    const float angle = fast_atan2_dev((float)m01, (float)m10);
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float ar = angle * factorPI;
    const float ca = (float)cos((double)ar), sa = (float)sin((double)ar);
    const uint8_t* bc = blur + (size_t)b * stride_b + L.off + (size_t)y * pitch + x;
    int val = 0;
    const int4 q0 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2);
    const int4 q1 = __ldg(reinterpret_cast<const int4*>(d_pattern) + lane * 2 + 1);
    const int words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const int wd = words[k];
        const float x0 = (float)(signed char)(wd & 0xff), y0 = (float)(signed char)((wd >> 8) & 0xff),
                    x1 = (float)(signed char)((wd >> 16) & 0xff), y1 = (float)(signed char)((wd >> 24) & 0xff);
        const int t0 = bc[__float2int_rn(x0 * sa + y0 * ca) * pitch + __float2int_rn(x0 * ca - y0 * sa)];
        const int t1 = bc[__float2int_rn(x1 * sa + y1 * ca) * pitch + __float2int_rn(x1 * ca - y1 * sa)];
        val |= (t0 < t1) << k;
    }
    out_desc[((size_t)b * feat_cap + g) * 32 + lane] = (uint8_t)val;
    if (lane == 0) {
        gd_keypoint k;
        k.x = (float)x * L.scale;  // cv::ORB: pt *= scale for every level (scale of level 0 is 1)
        k.y = (float)y * L.scale;
        k.size = 31.0f * L.scale;
        k.angle = angle;
        k.response = __uint_as_float(e.y);
        k.octave = l;
        k.class_id = -1;
        out_kp[(size_t)b * feat_cap + g] = k;
    }
}

int orb_cv_describe_sel(const uint8_t* pyr, const uint8_t* blur, size_t stride_b, const CvPyrArgs& a, int batch, const uint2* sel,
                        int sel_cap, const int* sel_n, gd_keypoint* out_kp, uint8_t* out_desc, int* out_n, int feat_cap, cudaStream_t s)
{
    static const int umax[ORB_HALF + 1] = {15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3};  // ORBextractor.cc:456-469
    CvDescArgs da;
    for (int i = 0; i <= ORB_HALF; ++i) da.umax[i] = umax[i];
    k_cv_describe_sel<<<dim3(cdiv(feat_cap, 8), batch), 256, 0, s>>>(pyr, blur, stride_b, a, da, sel, sel_cap, sel_n, out_kp, out_desc, out_n,
                                                                    feat_cap);
    GD_CUDA(cudaGetLastError());
    return GD_OK;
}

int OrbCore::init(int nfeatures_, float scale_factor_, int nlevels_, int ini_th, int min_th, int max_width, int max_height,
                  int device_, int batch_, cudaStream_t s, LaunchStats* st)
{
    GD_REQUIRE(batch_ >= 1, "bad batch");
    GD_TRY(select_device(device_));
    device = device_;
    batch = batch_;
    stats = st;
    nfeatures = nfeatures_;
    scale_factor = scale_factor_;
    nlevels = nlevels_;
    iniTh = ini_th;
    minTh = min_th;
    max_w = max_width;
    max_h = max_height;
    GD_TRY(orb_make_plan(nfeatures, scale_factor, nlevels, iniTh, minTh, max_w, max_h, &plan));
    if (s) {
        stream = s;
    } else {
        GD_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        own_stream = true;
    }
    const size_t B = (size_t)batch;
    GD_TRY(gray_in.alloc(B * (size_t)max_w * max_h));
    GD_TRY(pyr.alloc(B * plan.pyr_bytes));
    GD_TRY(blur.alloc(B * plan.pyr_bytes));
    GD_TRY(cell_cnt.alloc(B * plan.total_cells * sizeof(int)));
    GD_TRY(slabs.alloc(B * (size_t)plan.total_cells * plan.cell_cap * sizeof(ushort4)));
    GD_TRY(cand.alloc(B * (size_t)plan.cand_total * sizeof(ushort4)));
    GD_TRY(cand_q.alloc(B * (size_t)plan.cand_total));
    GD_TRY(kept.alloc(B * (size_t)plan.kept_total * sizeof(int)));
    GD_TRY(kept_cnt.alloc(B * nlevels * sizeof(int)));
    GD_TRY(cand_cnt.alloc(B * nlevels * sizeof(int)));
    GD_TRY(out_kp.alloc(B * (size_t)plan.kp_capacity * sizeof(gd_keypoint)));
    GD_TRY(out_desc.alloc(B * (size_t)plan.kp_capacity * 32));
    GD_TRY(out_n.alloc(B * sizeof(int)));
    GD_TRY(err.alloc(sizeof(int)));
    GD_TRY(h_n.alloc((B + 1) * sizeof(int)));
    GD_TRY(rs_tab.alloc((size_t)(max_w + max_h) * nlevels * sizeof(ushort4)));
    GD_TRY(upload_resize_tables());
    GD_CUDA(cudaMemsetAsync(err.p, 0, sizeof(int), stream));
    GD_CUDA(cudaMemsetAsync(pyr.p, 0, pyr.bytes, stream));
    const size_t fast_smem = (size_t)plan.tile_w * plan.tile_h * 4;
    if (fast_smem > 48 * 1024) GD_CUDA(cudaFuncSetAttribute(k_orb_fast, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fast_smem));
    const size_t qs = qt_smem_bytes(plan.max_N + 8, plan.max_cells_level);
    GD_REQUIRE(qs <= 220 * 1024, "nfeatures too large for the quadtree kernel's shared memory");
    if (qs > 48 * 1024) GD_CUDA(cudaFuncSetAttribute(k_orb_quadtree, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)qs));
    GD_CUDA(cudaStreamSynchronize(stream));
    return GD_OK;
}

int OrbCore::set_size(int w, int h)
{
    if (w == plan.w && h == plan.h) return GD_OK;
    GD_REQUIRE(w <= max_w && h <= max_h, "image larger than the size given at creation");
    OrbPlan p;
    GD_TRY(orb_make_plan(nfeatures, scale_factor, nlevels, iniTh, minTh, w, h, &p));
    // a smaller image needs no more memory than the plan the buffers were sized for
    OrbPlan big;
    GD_TRY(orb_make_plan(nfeatures, scale_factor, nlevels, iniTh, minTh, max_w, max_h, &big));
    GD_REQUIRE(p.pyr_bytes <= big.pyr_bytes && (size_t)p.total_cells * p.cell_cap <= (size_t)big.total_cells * big.cell_cap &&
                   p.cand_total <= big.cand_total && p.kept_total <= big.kept_total && p.kp_capacity <= big.kp_capacity &&
                   p.total_cells <= big.total_cells,
               "image size needs more memory than the maximum size given at creation");
    plan = p;
    plain_runs = 0;
    return upload_resize_tables();
}

int OrbCore::upload_resize_tables()
{
    std::vector<ushort4> tab;
    for (int l = 1; l < plan.nlevels; ++l) {
        const OrbLevel &S = plan.lv[l - 1], &D = plan.lv[l];
        rs_x_off[l] = (int)tab.size();
        tab.resize(tab.size() + D.w);
        orb_resize_axis_table(D.w, S.w, D.scale_x, true, tab.data() + rs_x_off[l]);
        rs_y_off[l] = (int)tab.size();
        tab.resize(tab.size() + D.h);
        orb_resize_axis_table(D.h, S.h, D.scale_y, false, tab.data() + rs_y_off[l]);
    }
    GD_REQUIRE(tab.size() * sizeof(ushort4) <= rs_tab.bytes, "resize tables exceed their buffer");
    if (!tab.empty()) {
        GD_CUDA(cudaStreamSynchronize(stream));  // a previous plan's tables may still be in use
        GD_CUDA(cudaMemcpyAsync(rs_tab.p, tab.data(), tab.size() * sizeof(ushort4), cudaMemcpyHostToDevice, stream));
        GD_CUDA(cudaStreamSynchronize(stream));
    }
    return GD_OK;
}

OrbCore::~OrbCore()
{
    if (own_stream && stream) cudaStreamDestroy(stream);
}

int OrbCore::extract_resident()
{
    // stand-alone handles replay the launch sequence as a graph once it has run twice un-captured at this image size
    if (plain_runs < 2) {
        ++plain_runs;
        return enqueue_extract();
    }
    return graphs.run(((unsigned long long)plan.w << 20) | (unsigned long long)plan.h, stream, stats, [&] { return enqueue_extract(); });
}

int OrbCore::enqueue_extract()
{
    PdlScope pdl_scope(batch);
    const OrbPlan& P = plan;
    uint8_t* py = pyr.as<uint8_t>();
    // K4a: pyramid chain
    for (int l = 1; l < P.nlevels; ++l) {
        LaunchScope ls(stats, stream, "K4a_pyramid_resize", 1);
        const OrbLevel &S = P.lv[l - 1], &D = P.lv[l];
        dim3 block(32, 8), grid(cdiv(D.w, 32), cdiv(D.h, 8 * RS_ROWS), batch);
        GD_CUDA(launch_pdl(k_orb_resize, grid, block, 0, stream, py + S.off, S.pitch, py + D.off, D.w, D.h, D.pitch, P.pyr_bytes,
                                                 rs_tab.as<ushort4>() + rs_x_off[l], rs_tab.as<ushort4>() + rs_y_off[l]));
        GD_CUDA(cudaGetLastError());
    }
    {  // K4b: FAST over every cell of every level
        LaunchScope ls(stats, stream, "K4b_fast_cells", 1);
        FastArgs fa;
        fa.nlevels = P.nlevels; fa.total_cells = P.total_cells; fa.cell_cap = P.cell_cap;
        fa.tile_w = P.tile_w; fa.tile_h = P.tile_h; fa.iniTh = P.iniTh; fa.minTh = P.minTh;
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            fa.lv[l] = {L.w, L.h, L.pitch, (unsigned long long)L.off, L.nCols, L.nRows, L.wCell, L.hCell, L.cell_start,
                        (1u << 20) / (unsigned)L.nCols + 1u};
        }
        for (int l = P.nlevels; l < ORB_MAX_LEVELS; ++l) {
            fa.lv[l] = fa.lv[0];
            fa.lv[l].cell_start = INT_MAX;
        }
        dim3 grid(P.total_cells, batch);
        GD_CUDA(launch_pdl(k_orb_fast, grid, dim3(FAST_THREADS), (size_t)P.tile_w * P.tile_h * 4, stream, py, P.pyr_bytes, fa, cell_cnt.as<int>(),
                                                                                      slabs.as<ushort4>()));
        GD_CUDA(cudaGetLastError());
    }
    {  // K4c: quadtree, one CTA per (level, stream)
        LaunchScope ls(stats, stream, "K4c_quadtree", 1);
        QtArgs qa;
        qa.nlevels = P.nlevels; qa.total_cells = P.total_cells; qa.cell_cap = P.cell_cap; qa.cand_total = P.cand_total;
        qa.kept_total = P.kept_total; qa.LN = P.max_N + 8; qa.max_cells = P.max_cells_level;
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            qa.lv[l] = {L.nCols, L.nRows, L.cell_start, L.N, L.nIni, L.cand_off, L.cand_cap, L.kept_off,
                        L.w - ORB_EDGE + 3 - ORB_BORDER, L.h - ORB_EDGE + 3 - ORB_BORDER, L.hX};
        }
        dim3 grid(P.nlevels, batch);
        GD_CUDA(launch_pdl(k_orb_quadtree, grid, dim3(QT_THREADS), qt_smem_bytes(qa.LN, qa.max_cells), stream, 
            qa, cell_cnt.as<int>(), slabs.as<ushort4>(), cand.as<ushort4>(), cand_q.as<uint8_t>(), kept.as<int>(),
            kept_cnt.as<int>(), cand_cnt.as<int>(), err.as<int>()));
        GD_CUDA(cudaGetLastError());
    }
    {  // K4e: 7x7 Gaussian of every level
        LaunchScope ls(stats, stream, "K4e_blur7", 1);
        BlurArgs ba;
        ba.nlevels = P.nlevels;
        int tiles = 0;
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            ba.lv[l] = {L.w, L.h, L.pitch, cdiv(L.w, BT_W), tiles, (unsigned long long)L.off};
            tiles += cdiv(L.w, BT_W) * cdiv(L.h, BT_H);
        }
        ba.total_tiles = tiles;
        dim3 grid(cdiv(tiles, 8), batch);  // one warp per 120 x 32 block
        GD_CUDA(launch_pdl(k_orb_blur, grid, dim3(256), 0, stream, py, blur.as<uint8_t>(), P.pyr_bytes, ba));
        GD_CUDA(cudaGetLastError());
    }
    {  // K4d/e: orientation + descriptors + output records
        LaunchScope ls(stats, stream, "K4de_orient_describe", 1);
        DescArgs da;
        da.nlevels = P.nlevels; da.cand_total = P.cand_total; da.kept_total = P.kept_total; da.kp_capacity = P.kp_capacity;
        for (int i = 0; i <= ORB_HALF; ++i) da.umax[i] = P.umax[i];
        for (int l = 0; l < P.nlevels; ++l) {
            const OrbLevel& L = P.lv[l];
            da.lv[l] = {L.pitch, L.cand_off, L.kept_off, (unsigned long long)L.off, L.scale, L.kp_size};
        }
        dim3 grid(cdiv(P.kp_capacity, 8), batch);
        GD_CUDA(launch_pdl(k_orb_describe, grid, dim3(256), 0, stream, py, blur.as<uint8_t>(), P.pyr_bytes, da, cand.as<ushort4>(), kept.as<int>(),
                                                 kept_cnt.as<int>(), out_kp.as<gd_keypoint>(), out_desc.as<uint8_t>(),
                                                 out_n.as<int>()));
        GD_CUDA(cudaGetLastError());
    }
    return GD_OK;
}

}  // namespace gd
