// orb.cuh — ORBextractor (ORB-SLAM2 / GD-SLAM src/ORBextractor.cc) as a batched CUDA pipeline.
#pragma once
#include "gd_internal.h"

namespace gd {

constexpr int ORB_MAX_LEVELS = 16;
constexpr int ORB_EDGE = 19;   // EDGE_THRESHOLD
constexpr int ORB_HALF = 15;   // HALF_PATCH_SIZE
constexpr int ORB_BORDER = 16; // minBorder = EDGE_THRESHOLD - 3

struct OrbLevel {
    int w, h, pitch;      // level size, row pitch in bytes
    size_t off;           // byte offset of the level inside one stream's pyramid buffer
    int nCols, nRows, wCell, hCell;  // FAST cell grid (ComputeKeyPointsOctTree, :781-787)
    int cell_start;       // index of the level's first cell in the per-stream cell arrays
    int N;                // mnFeaturesPerLevel
    int nIni;             // initial quadtree nodes
    float hX;
    int cand_off, cand_cap;  // slice of the compact candidate arrays
    int kept_off;            // slice of the kept-index arrays (capacity N + 4)
    float scale;             // mvScaleFactor[level]
    float kp_size;           // (float)(int)(31 * scale)
    double scale_x, scale_y; // resize ratios from the previous level
};

struct OrbPlan {
    int nlevels = 0, nfeatures = 0, iniTh = 20, minTh = 7;
    int w = 0, h = 0;
    OrbLevel lv[ORB_MAX_LEVELS];
    size_t pyr_bytes = 0;   // one stream's pyramid
    int total_cells = 0;
    int max_cells_level = 0;
    int cell_cap = 0;       // slab capacity per cell (NMS density bound)
    int tile_w = 0, tile_h = 0;  // largest cell (incl. the 6-px overlap)
    int cand_total = 0;     // compact candidate capacity per stream
    int kept_total = 0;
    int max_N = 0;
    int kp_capacity = 0;    // sum(N + 3)
    int umax[ORB_HALF + 1];
};

int orb_make_plan(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int w, int h, OrbPlan* plan);

struct OrbCore {
    int device = 0, batch = 0;
    OrbPlan plan;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    LaunchStats* stats = nullptr;
    int max_w = 0, max_h = 0;
    int nfeatures = 0, nlevels = 0, iniTh = 0, minTh = 0;
    float scale_factor = 0;
    size_t alloc_px = 0;

    DevBuf gray_in;   // [B][max_h][max_w] staging when the caller passes host gray
    DevBuf pyr;       // [B][pyr_bytes]   level 0 is filled by the caller (copy or K0)
    DevBuf blur;      // [B][pyr_bytes]
    DevBuf cell_cnt;  // [B][total_cells] int
    DevBuf slabs;     // [B][total_cells][cell_cap] ushort4 (x, y, response, -)
    DevBuf cand;      // [B][cand_total] ushort4 (x, y, response, node)
    DevBuf cand_q;    // [B][cand_total] u8
    DevBuf kept;      // [B][kept_total] int (candidate index)
    DevBuf kept_cnt;  // [B][nlevels] int
    DevBuf cand_cnt;  // [B][nlevels] int
    DevBuf out_kp;    // [B][kp_capacity] gd_keypoint
    DevBuf out_desc;  // [B][kp_capacity][32]
    DevBuf out_n;     // [B] int
    DevBuf err;       // [1] int  (capacity overflow flags)
    DevBuf rs_tab;    // cv::resize index/weight tables per level: x table then y table (ushort4)
    int rs_x_off[ORB_MAX_LEVELS] = {0}, rs_y_off[ORB_MAX_LEVELS] = {0};
    PinnedBuf h_n;

    int init(int nfeatures, float scale_factor, int nlevels, int ini_th, int min_th, int max_width, int max_height,
             int device_, int batch_, cudaStream_t s, LaunchStats* st);
    int set_size(int w, int h);  // (re)plans for an image size <= max
    int upload_resize_tables();
    uint8_t* level0(int b = 0) { return pyr.as<uint8_t>() + (size_t)b * plan.pyr_bytes; }
    // level 0 of every stream already resident in `pyr` (pitch = plan.lv[0].pitch): run the whole extraction
    int extract_resident();
    int enqueue_extract();  // the launches of extract_resident() (graph capturable, no host state)
    GraphCache graphs;      // enabled for stand-alone handles only (the batched front-end captures its whole step)
    int plain_runs = 0;
    ~OrbCore();
};

// IC_Angle + rBRIEF for keypoints at integer coordinates of one level (cv::ORB: d_blur = the float-blurred level)
int orb_cv_describe(const uint8_t* d_img, const uint8_t* d_blur, int pitch, const int* d_xs, const int* d_ys, int n, float* d_angle,
                    uint8_t* d_desc, cudaStream_t s);
// cv::FAST(th, nonmax) on a whole image (cv::ORB per level, GetRt): kept[y][x] = S' at surviving corners, else 0
int orb_fast_whole(const uint8_t* d_img, int w, int h, int pitch, int th, uint8_t* d_score, uint8_t* d_kept, cudaStream_t s);

// ---- batched cv::ORB blocks of the resident GetRt stage (getrt.cu): dense pyramid levels (pitch = width)
struct CvLevelDev {
    int w, h;
    unsigned long long off;  // byte offset of the level inside one stream's pyramid
    int row_off;             // offset of the level's rows inside the per-stream row-count array
    float scale;             // (float)pow((double)1.2f, level)
};
struct CvPyrArgs {
    int nlevels;
    CvLevelDev lv[8];
};
// cv::FAST(th, nonmax) on every level of every stream: score map, then the NMS map (S' at surviving corners) and, per level
// row, the number of survivors inside the `edge` border (rowcnt must be zero on entry)
int orb_cv_fast_levels(const uint8_t* pyr, size_t stride_b, const CvPyrArgs& a, int batch, int th, int edge, uint8_t* score,
                       uint8_t* kept, int* rowcnt, size_t rowcnt_stride, cudaStream_t s);
// orientation + descriptors + cv::KeyPoint records for the selected keypoints (sel[b][level][sel_cap]: .x = y * w + x, .y = bits
// of the Harris response), concatenated level by level
int orb_cv_describe_sel(const uint8_t* pyr, const uint8_t* blur, size_t stride_b, const CvPyrArgs& a, int batch, const uint2* sel,
                        int sel_cap, const int* sel_n, gd_keypoint* out_kp, uint8_t* out_desc, int* out_n, int feat_cap, cudaStream_t s);

}  // namespace gd
