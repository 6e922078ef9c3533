// orb_oracle.cpp — CPU restatement of ORB_SLAM2::ORBextractor as used by GD-SLAM (TEST INFRASTRUCTURE, see
// gd_oracle.h).  Follows /root/reference/src/ORBextractor.cc:
//   ctor tables                :410-470   -> OrbCfg::OrbCfg
//   ComputePyramid             :1107-1132 -> pyramid()
//   ComputeKeyPointsOctTree    :765-853   -> candidates() (cell loop :789-829), distribute() (:539-763, :481-537),
//                                            fix-up :837-847, IC_Angle :77-104
//   operator()                 :1043-1105 -> gdo_orb_extract (blur :1085-1086, computeOrbDescriptor :108-147, scaling :1095-1101)
// The quadtree is restated with an index-linked list over a node pool instead of std::list: list position semantics
// (push_front / erase / iteration order) are kept; the reference's tie-break by list-node ADDRESS
// (sort of pair<int,ExtractorNode*>, :684) becomes the canonical "later-created node = higher address" (SURVEY B-3),
// i.e. what the verbatim reference code does under a monotonic allocator (oracle/_ref is built that way).
// cos/sin of the keypoint angle: correctly rounded f32, (float)cos((double)a) (SURVEY B-5).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "orb_prims.hpp"

namespace {

const int PATCH_SIZE = 31, HALF_PATCH_SIZE = 15, EDGE_THRESHOLD = 19;
const int kPattern[1024] = {
#include "orb_pattern.inc"
};

struct OrbCfg {
    int nfeatures, nlevels, iniTh, minTh;
    double scaleFactor;
    std::vector<float> scale, invScale;
    std::vector<int> nPerLevel, umax;
    OrbCfg(int nf, float sf, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(sf)
    {
        scale.resize(nl);
        invScale.resize(nl);
        scale[0] = 1.0f;
        for (int i = 1; i < nl; i++) scale[i] = (float)(scale[i - 1] * scaleFactor);
        for (int i = 0; i < nl; i++) invScale[i] = 1.0f / scale[i];
        nPerLevel.resize(nl);
        float factor = (float)(1.0f / scaleFactor);
        float nDesired = (float)(nf * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nl)));
        int sum = 0;
        for (int l = 0; l < nl - 1; l++) {
            nPerLevel[l] = gdo::cv_round(nDesired);
            sum += nPerLevel[l];
            nDesired *= factor;
        }
        nPerLevel[nl - 1] = std::max(nf - sum, 0);
        umax.resize(HALF_PATCH_SIZE + 1);
        int v, v0, vmax = gdo::cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
        int vmin = gdo::cv_ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
        const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
        for (v = 0; v <= vmax; ++v) umax[v] = gdo::cv_round(std::sqrt(hp2 - v * v));
        for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
            while (umax[v0] == umax[v0 + 1]) ++v0;
            umax[v] = v0;
            ++v0;
        }
    }
    void level_size(int w, int h, int l, int* lw, int* lh) const
    {
        *lw = gdo::cv_round((float)w * invScale[l]);
        *lh = gdo::cv_round((float)h * invScale[l]);
    }
};

struct Cand {
    float x, y;  // relative to (minBorderX, minBorderY)
    float response;
};

// cell loop of ComputeKeyPointsOctTree (:771-829)
void candidates(const uint8_t* img, int cols, int rows, int iniTh, int minTh, std::vector<Cand>& out)
{
    out.clear();
    const float W = 30;
    const int minBorderX = EDGE_THRESHOLD - 3, minBorderY = minBorderX;
    const int maxBorderX = cols - EDGE_THRESHOLD + 3, maxBorderY = rows - EDGE_THRESHOLD + 3;
    const float width = (float)(maxBorderX - minBorderX), height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W), nRows = (int)(height / W);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<gdo::FastKp> cell;
    for (int i = 0; i < nRows; i++) {
        const float iniY = (float)(minBorderY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBorderY - 3) continue;
        if (maxY > maxBorderY) maxY = (float)maxBorderY;
        for (int j = 0; j < nCols; j++) {
            const float iniX = (float)(minBorderX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBorderX - 6) continue;
            if (maxX > maxBorderX) maxX = (float)maxBorderX;
            const int x0 = (int)iniX, x1 = (int)maxX, y0 = (int)iniY, y1 = (int)maxY;
            gdo::fast_detect(img + (size_t)y0 * cols + x0, x1 - x0, y1 - y0, (size_t)cols, iniTh, cell);
            if (cell.empty()) gdo::fast_detect(img + (size_t)y0 * cols + x0, x1 - x0, y1 - y0, (size_t)cols, minTh, cell);
            for (const auto& k : cell) out.push_back({(float)k.x + j * wCell, (float)k.y + i * hCell, (float)k.response});
        }
    }
}

// ---- DistributeOctTree (:539-763) with ExtractorNode::DivideNode (:481-537) ------------------------------------
struct QNode {
    int ulx, uly, brx, bry;  // UL = (ulx,uly), UR = (brx,uly), BL = (ulx,bry), BR = (brx,bry)
    std::vector<int> keys;
    bool noMore = false;
    int prev = -1, next = -1;
    bool alive = false;
};

struct QList {
    std::vector<QNode> pool;  // index = creation sequence number ("address")
    int head = -1, tail = -1, count = 0;
    int push_back(QNode&& n)
    {
        pool.push_back(std::move(n));
        const int i = (int)pool.size() - 1;
        pool[i].alive = true;
        pool[i].prev = tail;
        pool[i].next = -1;
        if (tail >= 0) pool[tail].next = i; else head = i;
        tail = i;
        ++count;
        return i;
    }
    int push_front(QNode&& n)
    {
        pool.push_back(std::move(n));
        const int i = (int)pool.size() - 1;
        pool[i].alive = true;
        pool[i].next = head;
        pool[i].prev = -1;
        if (head >= 0) pool[head].prev = i; else tail = i;
        head = i;
        ++count;
        return i;
    }
    int erase(int i)  // returns next
    {
        const int p = pool[i].prev, n = pool[i].next;
        if (p >= 0) pool[p].next = n; else head = n;
        if (n >= 0) pool[n].prev = p; else tail = p;
        pool[i].alive = false;
        --count;
        return n;
    }
};

void divide(const QNode& nd, const std::vector<Cand>& c, QNode ch[4])
{
    const int halfX = (int)std::ceil((float)(nd.brx - nd.ulx) / 2);
    const int halfY = (int)std::ceil((float)(nd.bry - nd.uly) / 2);
    const int mx = nd.ulx + halfX, my = nd.uly + halfY;
    ch[0].ulx = nd.ulx; ch[0].uly = nd.uly; ch[0].brx = mx;     ch[0].bry = my;
    ch[1].ulx = mx;     ch[1].uly = nd.uly; ch[1].brx = nd.brx; ch[1].bry = my;
    ch[2].ulx = nd.ulx; ch[2].uly = my;     ch[2].brx = mx;     ch[2].bry = nd.bry;
    ch[3].ulx = mx;     ch[3].uly = my;     ch[3].brx = nd.brx; ch[3].bry = nd.bry;
    for (int k : nd.keys) {
        const Cand& kp = c[k];
        if (kp.x < (float)mx) {
            if (kp.y < (float)my) ch[0].keys.push_back(k); else ch[2].keys.push_back(k);
        } else if (kp.y < (float)my)
            ch[1].keys.push_back(k);
        else
            ch[3].keys.push_back(k);
    }
    for (int i = 0; i < 4; ++i) ch[i].noMore = ch[i].keys.size() == 1;
}

void distribute(const std::vector<Cand>& c, int minX, int maxX, int minY, int maxY, int N, std::vector<int>& result)
{
    result.clear();
    if (c.empty()) return;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    const float hX = (float)(maxX - minX) / nIni;
    QList L;
    L.pool.reserve(4 * c.size() + 16);
    std::vector<int> ini(nIni);
    for (int i = 0; i < nIni; i++) {
        QNode n;
        n.ulx = (int)(hX * (float)i);
        n.brx = (int)(hX * (float)(i + 1));
        n.uly = 0;
        n.bry = maxY - minY;
        ini[i] = L.push_back(std::move(n));
    }
    for (int k = 0; k < (int)c.size(); ++k) L.pool[ini[(int)(c[k].x / hX)]].keys.push_back(k);
    for (int it = L.head; it >= 0;) {
        if (L.pool[it].keys.size() == 1) {
            L.pool[it].noMore = true;
            it = L.pool[it].next;
        } else if (L.pool[it].keys.empty())
            it = L.erase(it);
        else
            it = L.pool[it].next;
    }
    bool finish = false;
    std::vector<std::pair<int, int>> sizeAndNode;  // (key count, pool index)
    auto add_children = [&](QNode ch[4], int* nToExpand) {
        for (int q = 0; q < 4; ++q)
            if (!ch[q].keys.empty()) {
                const int sz = (int)ch[q].keys.size();
                const int idx = L.push_front(std::move(ch[q]));
                if (sz > 1) {
                    if (nToExpand) ++*nToExpand;
                    sizeAndNode.push_back({sz, idx});
                }
            }
    };
    while (!finish) {
        const int prevSize = L.count;
        int nToExpand = 0;
        sizeAndNode.clear();
        for (int it = L.head; it >= 0;) {
            if (L.pool[it].noMore) {
                it = L.pool[it].next;
                continue;
            }
            QNode ch[4];
            divide(L.pool[it], c, ch);
            add_children(ch, &nToExpand);
            it = L.erase(it);
        }
        if (L.count >= N || L.count == prevSize)
            finish = true;
        else if (L.count + nToExpand * 3 > N) {
            while (!finish) {
                const int prevSize2 = L.count;
                std::vector<std::pair<int, int>> prev = sizeAndNode;
                sizeAndNode.clear();
                std::sort(prev.begin(), prev.end());
                for (int j = (int)prev.size() - 1; j >= 0; j--) {
                    QNode ch[4];
                    divide(L.pool[prev[j].second], c, ch);
                    add_children(ch, nullptr);
                    L.erase(prev[j].second);
                    if (L.count >= N) break;
                }
                if (L.count >= N || L.count == prevSize2) finish = true;
            }
        }
    }
    for (int it = L.head; it >= 0; it = L.pool[it].next) {
        const std::vector<int>& ks = L.pool[it].keys;
        int best = ks[0];
        float maxResponse = c[best].response;
        for (size_t k = 1; k < ks.size(); k++)
            if (c[ks[k]].response > maxResponse) {
                best = ks[k];
                maxResponse = c[ks[k]].response;
            }
        result.push_back(best);
    }
}

// IC_Angle (:77-104)
float ic_angle(const uint8_t* img, int cols, int x, int y, const std::vector<int>& umax)
{
    int m_01 = 0, m_10 = 0;
    const uint8_t* center = img + (size_t)y * cols + x;
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * center[u];
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
        int v_sum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int val_plus = center[u + v * cols], val_minus = center[u - v * cols];
            v_sum += (val_plus - val_minus);
            m_10 += u * (val_plus + val_minus);
        }
        m_01 += v * v_sum;
    }
    return gdo::fast_atan2((float)m_01, (float)m_10);
}

// computeOrbDescriptor (:108-147)
void orb_descriptor(const uint8_t* img, int cols, int x, int y, float angle_deg, uint8_t* desc)
{
    const float factorPI = (float)(3.1415926535897932384626433832795 / 180.f);
    const float angle = angle_deg * factorPI;
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const uint8_t* center = img + (size_t)y * cols + x;
    const int* p = kPattern;
    auto val = [&](int idx) -> int {
        const int px = p[2 * idx], py = p[2 * idx + 1];
        const float fy = (float)px * b + (float)py * a;
        const float fx = (float)px * a - (float)py * b;
        return center[gdo::cv_round(fy) * cols + gdo::cv_round(fx)];
    };
    for (int i = 0; i < 32; ++i, p += 32) {
        int v = 0;
        for (int k = 0; k < 8; ++k) v |= (val(2 * k) < val(2 * k + 1)) << k;
        desc[i] = (uint8_t)v;
    }
}

struct Kp {
    float x, y, size, angle, response;
    int32_t octave, class_id;
};

}  // namespace

extern "C" {

void gdo_orb_config(int nfeatures, float scale, int nlevels, int w, int h, int* n_per_level, int* level_sizes, float* scales,
                    int* umax)
{
    OrbCfg cfg(nfeatures, scale, nlevels, 20, 7);
    for (int l = 0; l < nlevels; ++l) {
        if (n_per_level) n_per_level[l] = cfg.nPerLevel[l];
        if (level_sizes) cfg.level_size(w, h, l, &level_sizes[2 * l], &level_sizes[2 * l + 1]);
        if (scales) scales[l] = cfg.scale[l];
    }
    if (umax)
        for (int i = 0; i <= HALF_PATCH_SIZE; ++i) umax[i] = cfg.umax[i];
}

float gdo_fast_atan2(float y, float x) { return gdo::fast_atan2(y, x); }

void gdo_resize_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh)
{
    gdo::resize_linear_u8(src, sw, sh, (size_t)sw, dst, dw, dh, (size_t)dw);
}

void gdo_gaussian7_u8(const uint8_t* src, int w, int h, uint8_t* dst) { gdo::gaussian7_u8(src, w, h, (size_t)w, dst, (size_t)w); }

// cv::FAST on a whole (sub-)image; out: (x, y, response) int triples, returns count (<= capacity written)
int gdo_fast_detect(const uint8_t* img, int cols, int rows, int threshold, int* out, int capacity)
{
    std::vector<gdo::FastKp> k;
    gdo::fast_detect(img, cols, rows, (size_t)cols, threshold, k);
    for (int i = 0; i < (int)k.size() && i < capacity; ++i) {
        out[3 * i] = k[i].x;
        out[3 * i + 1] = k[i].y;
        out[3 * i + 2] = k[i].response;
    }
    return (int)k.size();
}

// per-level candidate list of ComputeKeyPointsOctTree's cell loop; out: (x, y, response) float triples (border-relative)
int gdo_orb_candidates(const uint8_t* img, int cols, int rows, int ini_th, int min_th, float* out, int capacity)
{
    std::vector<Cand> c;
    candidates(img, cols, rows, ini_th, min_th, c);
    for (int i = 0; i < (int)c.size() && i < capacity; ++i) {
        out[3 * i] = c[i].x;
        out[3 * i + 1] = c[i].y;
        out[3 * i + 2] = c[i].response;
    }
    return (int)c.size();
}

// DistributeOctTree on a candidate list; writes the kept candidate indices in result order
int gdo_orb_distribute(const float* cand, int n, int minX, int maxX, int minY, int maxY, int N, int* kept, int capacity)
{
    std::vector<Cand> c(n);
    for (int i = 0; i < n; ++i) c[i] = {cand[3 * i], cand[3 * i + 1], cand[3 * i + 2]};
    std::vector<int> r;
    distribute(c, minX, maxX, minY, maxY, N, r);
    for (int i = 0; i < (int)r.size() && i < capacity; ++i) kept[i] = r[i];
    return (int)r.size();
}

// cv::KeyPointsFilter::retainBest as OpenCV writes it (std::nth_element + std::partition on the responses): returns the
// number kept and the permutation the routine leaves behind (perm[i] = original index of the i-th kept element).  The
// order is whatever libstdc++'s introselect produces; it is what cv::ORB's output order consists of.
int gdo_retain_best_order(const float* resp, int n, int n_points, int* perm)
{
    struct E {
        float r;
        int i;
    };
    std::vector<E> v((size_t)n);
    for (int i = 0; i < n; ++i) v[i] = {resp[i], i};
    if (n_points >= 0 && n > n_points) {
        if (n_points == 0) return 0;
        std::nth_element(v.begin(), v.begin() + n_points - 1, v.end(), [](const E& a, const E& b) { return a.r > b.r; });
        const float amb = v[n_points - 1].r;
        auto e = std::partition(v.begin() + n_points, v.end(), [amb](const E& a) { return a.r >= amb; });
        v.resize((size_t)(e - v.begin()));
    }
    for (size_t i = 0; i < v.size(); ++i) perm[i] = v[i].i;
    return (int)v.size();
}

// std::sort(matches.begin(), matches.end()) of GeoMaskMaker.cc:95 (cv::DMatch::operator< compares the distance only): the
// order among equal distances is whatever libstdc++'s introsort produces; perm[i] = original index of the i-th match.
void gdo_sort_matches_order(const float* dist, int n, int* perm)
{
    struct E {
        float d;
        int i;
        bool operator<(const E& o) const { return d < o.d; }
    };
    std::vector<E> v((size_t)n);
    for (int i = 0; i < n; ++i) v[i] = {dist[i], i};
    std::sort(v.begin(), v.end());
    for (int i = 0; i < n; ++i) perm[i] = v[i].i;
}

// rBRIEF descriptor of one keypoint on an already blurred image (computeOrbDescriptor, ORBextractor.cc:108-147)
void gdo_orb_descriptor(const uint8_t* blurred, int cols, int x, int y, float angle_deg, uint8_t* desc32)
{
    orb_descriptor(blurred, cols, x, y, angle_deg, desc32);
}

float gdo_ic_angle(const uint8_t* img, int cols, int x, int y)
{
    OrbCfg cfg(1500, 1.2f, 8, 20, 7);
    return ic_angle(img, cols, x, y, cfg.umax);
}

// ORBextractor::operator().  kps: 7 x 4-byte records (cv::KeyPoint layout); desc: n x 32.
// pyramid_out (optional): levels concatenated; n_level_out (optional): keypoints per level.
int gdo_orb_extract(const uint8_t* gray, int w, int h, size_t step, int nfeatures, float scale, int nlevels, int ini_th,
                    int min_th, void* kps_out, uint8_t* desc_out, int capacity, uint8_t* pyramid_out, int* n_level_out)
{
    OrbCfg cfg(nfeatures, scale, nlevels, ini_th, min_th);
    std::vector<std::vector<uint8_t>> pyr(nlevels);
    std::vector<int> lw(nlevels), lh(nlevels);
    for (int l = 0; l < nlevels; ++l) {
        cfg.level_size(w, h, l, &lw[l], &lh[l]);
        pyr[l].resize((size_t)lw[l] * lh[l]);
        if (l == 0)
            for (int y = 0; y < h; ++y) std::memcpy(&pyr[0][(size_t)y * w], gray + (size_t)y * step, (size_t)w);
        else
            gdo::resize_linear_u8(pyr[l - 1].data(), lw[l - 1], lh[l - 1], (size_t)lw[l - 1], pyr[l].data(), lw[l], lh[l], (size_t)lw[l]);
    }
    if (pyramid_out) {
        size_t off = 0;
        for (int l = 0; l < nlevels; ++l) {
            std::memcpy(pyramid_out + off, pyr[l].data(), pyr[l].size());
            off += pyr[l].size();
        }
    }
    Kp* out = (Kp*)kps_out;
    int total = 0;
    std::vector<Cand> cand;
    std::vector<int> kept;
    std::vector<uint8_t> blurred;
    for (int l = 0; l < nlevels; ++l) {
        const int minB = EDGE_THRESHOLD - 3;
        const int maxBX = lw[l] - EDGE_THRESHOLD + 3, maxBY = lh[l] - EDGE_THRESHOLD + 3;
        candidates(pyr[l].data(), lw[l], lh[l], ini_th, min_th, cand);
        distribute(cand, minB, maxBX, minB, maxBY, cfg.nPerLevel[l], kept);
        if (n_level_out) n_level_out[l] = (int)kept.size();
        if (kept.empty()) continue;
        const int scaledPatchSize = (int)(PATCH_SIZE * cfg.scale[l]);
        blurred.resize(pyr[l].size());
        gdo::gaussian7_u8(pyr[l].data(), lw[l], lh[l], (size_t)lw[l], blurred.data(), (size_t)lw[l]);
        for (int idx : kept) {
            Kp k;
            k.x = cand[idx].x + minB;
            k.y = cand[idx].y + minB;
            k.response = cand[idx].response;
            k.octave = l;
            k.class_id = -1;
            k.size = (float)scaledPatchSize;
            const int ix = gdo::cv_round(k.x), iy = gdo::cv_round(k.y);
            k.angle = ic_angle(pyr[l].data(), lw[l], ix, iy, cfg.umax);
            if (total < capacity) {
                if (desc_out) orb_descriptor(blurred.data(), lw[l], ix, iy, k.angle, desc_out + (size_t)total * 32);
                if (l != 0) {
                    k.x *= cfg.scale[l];
                    k.y *= cfg.scale[l];
                }
                if (out) out[total] = k;
            }
            ++total;
        }
    }
    return total;
}

}  // extern "C"
