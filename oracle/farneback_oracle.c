/*
 * farneback_oracle.c — CPU restatement of cv::calcOpticalFlowFarneback (flags = 0, box window) as
 * invoked by GeoMaskMaker::GetFlow (/root/reference/src/GeoMaskMaker.cc:158-166) with
 * (pyr_scale 0.5, levels 3, winsize 15, iterations 3, poly_n 5, poly_sigma 1.2, flags 0).
 * TEST INFRASTRUCTURE (see gd_oracle.h).
 *
 * The algorithm lives in OpenCV (video/optflowgf.cpp), which the reference does not vendor and
 * does not pin (CMakeLists.txt:48-54; the committed binary linked 3.4).  The contract is OpenCV
 * 4.13.0 semantics (the cv2 wheel of this image).  This file restates the published algorithm
 * (SURVEY.md A4); it is pinned against cv2.calcOpticalFlowFarneback outputs committed under
 * tests/golden/ (tolerance |d| <= 1e-4*max(1,|ref|), see tests/test_oracle_farneback.py).
 *
 * Stage map (names follow OpenCV's):
 *   GaussianBlur(f32, ksize s, sigma) on the FULL-RESOLUTION image, REFLECT_101
 *   resize(INTER_LINEAR) to the level size
 *   FarnebackPolyExp(n=5, sigma=1.2)            -> R (5 x f32 / px)
 *   FarnebackUpdateMatrices(R0,R1,flow)         -> M (5 x f32 / px)
 *   FarnebackUpdateFlow_Blur(block 15)          -> flow; 3 iterations per level
 */
#include "gd_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

static int cv_round(double v) { return (int)lrint(v); }
static int reflect101(int p, int len)
{
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * len - 2 - p;
    }
    return p;
}

/* cv::getGaussianKernel(n, sigma, CV_32F) */
static void gaussian_kernel(int n, double sigma, float* k)
{
    if (sigma <= 0 && n == 3) {
        k[0] = 0.25f; k[1] = 0.5f; k[2] = 0.25f;
        return;
    }
    const double sx = sigma > 0 ? sigma : ((n - 1) * 0.5 - 1) * 0.3 + 0.8;
    const double scale2x = -0.5 / (sx * sx);
    double sum = 0, t[64];
    for (int i = 0; i < n; ++i) {
        const double x = i - (n - 1) * 0.5;
        t[i] = exp(scale2x * x * x);
        sum += t[i];
    }
    sum = 1.0 / sum;
    for (int i = 0; i < n; ++i) k[i] = (float)(t[i] * sum);
}

/* separable f32 Gaussian, row pass then column pass, REFLECT_101 */
static void gaussian_blur_f32(const float* src, int w, int h, int ksize, double sigma, float* dst)
{
    float k[64];
    gaussian_kernel(ksize, sigma, k);
    const int r = ksize / 2;
    float* tmp = (float*)malloc((size_t)w * h * sizeof(float));
    for (int y = 0; y < h; ++y) {
        const float* s = src + (size_t)y * w;
        for (int x = 0; x < w; ++x) {
            float acc = k[0] * s[reflect101(x - r, w)];
            for (int i = 1; i < ksize; ++i) acc += k[i] * s[reflect101(x - r + i, w)];
            tmp[(size_t)y * w + x] = acc;
        }
    }
    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            float acc = k[r] * tmp[(size_t)y * w + x];
            for (int i = 1; i <= r; ++i)
                acc += k[r + i] * (tmp[(size_t)reflect101(y + i, h) * w + x] + tmp[(size_t)reflect101(y - i, h) * w + x]);
            dst[(size_t)y * w + x] = acc;
        }
    }
    free(tmp);
}

/* cv::resize INTER_LINEAR for f32, cn interleaved channels */
static void resize_linear_f32(const float* src, int sw, int sh, float* dst, int dw, int dh, int cn)
{
    if (sw == dw && sh == dh) {
        memcpy(dst, src, (size_t)sw * sh * cn * sizeof(float));
        return;
    }
    const double scale_x = (double)sw / dw, scale_y = (double)sh / dh;
    int* xofs = (int*)malloc(dw * sizeof(int));
    float* xa = (float*)malloc(dw * sizeof(float));
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)floorf(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= sw - 1) { fx = 0; sx = sw - 1; }
        xofs[dx] = sx;
        xa[dx] = fx;
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)floorf(fy);
        fy -= sy;
        /* OpenCV clips the two source rows individually and keeps the unclamped weights */
        int sy1 = sy + 1;
        sy = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy);
        sy1 = sy1 < 0 ? 0 : (sy1 > sh - 1 ? sh - 1 : sy1);
        const float b0 = 1.f - fy, b1 = fy;
        for (int dx = 0; dx < dw; ++dx) {
            const int sx = xofs[dx];
            const int sx1 = sx < sw - 1 ? sx + 1 : sx;
            const float a0 = 1.f - xa[dx], a1 = xa[dx];
            for (int c = 0; c < cn; ++c) {
                const float h0 = src[((size_t)sy * sw + sx) * cn + c] * a0 + src[((size_t)sy * sw + sx1) * cn + c] * a1;
                const float h1 = src[((size_t)sy1 * sw + sx) * cn + c] * a0 + src[((size_t)sy1 * sw + sx1) * cn + c] * a1;
                dst[((size_t)dy * dw + dx) * cn + c] = h0 * b0 + h1 * b1;
            }
        }
    }
    free(xofs);
    free(xa);
}

/* FarnebackPrepareGaussian */
static void prepare_gaussian(int n, double sigma, float* g, float* xg, float* xxg, double* ig11, double* ig03,
                             double* ig33, double* ig55)
{
    if (sigma < FLT_EPSILON) sigma = n * 0.3;
    double s = 0.;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)exp(-x * x / (2 * sigma * sigma));
        s += g[x];
    }
    s = 1. / s;
    for (int x = -n; x <= n; x++) {
        g[x] = (float)(g[x] * s);
        xg[x] = (float)(x * g[x]);
        xxg[x] = (float)(x * x * g[x]);
    }
    double G[6][6];
    memset(G, 0, sizeof(G));
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
    /* invert (Gauss-Jordan, f64; OpenCV uses Cholesky — same to ~1e-16 relative) */
    double A[6][12];
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) {
            A[i][j] = G[i][j];
            A[i][6 + j] = i == j ? 1.0 : 0.0;
        }
    for (int c = 0; c < 6; ++c) {
        int p = c;
        for (int r = c + 1; r < 6; ++r)
            if (fabs(A[r][c]) > fabs(A[p][c])) p = r;
        if (p != c)
            for (int j = 0; j < 12; ++j) { double t = A[c][j]; A[c][j] = A[p][j]; A[p][j] = t; }
        const double d = 1.0 / A[c][c];
        for (int j = 0; j < 12; ++j) A[c][j] *= d;
        for (int r = 0; r < 6; ++r)
            if (r != c) {
                const double f = A[r][c];
                if (f != 0.0)
                    for (int j = 0; j < 12; ++j) A[r][j] -= f * A[c][j];
            }
    }
    *ig11 = A[1][7];
    *ig03 = A[0][9];
    *ig33 = A[3][9];
    *ig55 = A[5][11];
}

/* FarnebackPolyExp: src f32 w*h -> dst 5 x f32 / px */
static void poly_exp(const float* src, int width, int height, int n, double sigma, float* dst)
{
    float* kbuf = (float*)malloc((n * 6 + 3) * sizeof(float));
    float* _row = (float*)malloc((size_t)(width + n * 2) * 3 * sizeof(float));
    float* g = kbuf + n;
    float* xg = g + n * 2 + 1;
    float* xxg = xg + n * 2 + 1;
    float* row = _row + n * 3;
    double ig11, ig03, ig33, ig55;
    prepare_gaussian(n, sigma, g, xg, xxg, &ig11, &ig03, &ig33, &ig55);

    for (int y = 0; y < height; y++) {
        float g0 = g[0], g1, g2;
        const float* srow0 = src + (size_t)y * width;
        const float* srow1 = 0;
        float* drow = dst + (size_t)y * width * 5;
        for (int x = 0; x < width; x++) {
            row[x * 3] = srow0[x] * g0;
            row[x * 3 + 1] = row[x * 3 + 2] = 0.f;
        }
        for (int k = 1; k <= n; k++) {
            g0 = g[k]; g1 = xg[k]; g2 = xxg[k];
            srow0 = src + (size_t)(y - k > 0 ? y - k : 0) * width;
            srow1 = src + (size_t)(y + k < height - 1 ? y + k : height - 1) * width;
            for (int x = 0; x < width; x++) {
                float p = srow0[x] + srow1[x];
                float t0 = row[x * 3] + g0 * p;
                float t1 = row[x * 3 + 1] + g1 * (srow1[x] - srow0[x]);
                float t2 = row[x * 3 + 2] + g2 * p;
                row[x * 3] = t0;
                row[x * 3 + 1] = t1;
                row[x * 3 + 2] = t2;
            }
        }
        for (int x = 0; x < n * 3; x++) {
            row[-1 - x] = row[2 - x];
            row[width * 3 + x] = row[width * 3 + x - 3];
        }
        for (int x = 0; x < width; x++) {
            g0 = g[0];
            double b1 = row[x * 3] * g0, b2 = 0, b3 = row[x * 3 + 1] * g0, b4 = 0, b5 = row[x * 3 + 2] * g0, b6 = 0;
            for (int k = 1; k <= n; k++) {
                double tg = row[(x + k) * 3] + row[(x - k) * 3];
                g0 = g[k];
                b1 += tg * g0;
                b4 += tg * xxg[k];
                b2 += (row[(x + k) * 3] - row[(x - k) * 3]) * xg[k];
                b3 += (row[(x + k) * 3 + 1] + row[(x - k) * 3 + 1]) * g0;
                b6 += (row[(x + k) * 3 + 1] - row[(x - k) * 3 + 1]) * xg[k];
                b5 += (row[(x + k) * 3 + 2] + row[(x - k) * 3 + 2]) * g0;
            }
            drow[x * 5 + 1] = (float)(b2 * ig11);
            drow[x * 5] = (float)(b3 * ig11);
            drow[x * 5 + 3] = (float)(b1 * ig03 + b4 * ig33);
            drow[x * 5 + 2] = (float)(b1 * ig03 + b5 * ig33);
            drow[x * 5 + 4] = (float)(b6 * ig55);
        }
    }
    free(kbuf);
    free(_row);
}

/* FarnebackUpdateMatrices over all rows */
static void update_matrices(const float* R0a, const float* R1, const float* flowa, float* Ma, int width, int height)
{
    enum { BORDER = 5 };
    static const float border[BORDER] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
    const size_t step1 = (size_t)width * 5;
    for (int y = 0; y < height; y++) {
        const float* flow = flowa + (size_t)y * width * 2;
        const float* R0 = R0a + (size_t)y * step1;
        float* M = Ma + (size_t)y * step1;
        for (int x = 0; x < width; x++) {
            float dx = flow[x * 2], dy = flow[x * 2 + 1];
            float fx = x + dx, fy = y + dy;
            int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
            float r2, r3, r4, r5, r6;
            fx -= x1;
            fy -= y1;
            if ((unsigned)x1 < (unsigned)(width - 1) && (unsigned)y1 < (unsigned)(height - 1)) {
                const float* ptr = R1 + (size_t)y1 * step1 + (size_t)x1 * 5;
                float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
                r2 = a00 * ptr[0] + a01 * ptr[5] + a10 * ptr[step1] + a11 * ptr[step1 + 5];
                r3 = a00 * ptr[1] + a01 * ptr[6] + a10 * ptr[step1 + 1] + a11 * ptr[step1 + 6];
                r4 = a00 * ptr[2] + a01 * ptr[7] + a10 * ptr[step1 + 2] + a11 * ptr[step1 + 7];
                r5 = a00 * ptr[3] + a01 * ptr[8] + a10 * ptr[step1 + 3] + a11 * ptr[step1 + 8];
                r6 = a00 * ptr[4] + a01 * ptr[9] + a10 * ptr[step1 + 4] + a11 * ptr[step1 + 9];
                r4 = (R0[x * 5 + 2] + r4) * 0.5f;
                r5 = (R0[x * 5 + 3] + r5) * 0.5f;
                r6 = (R0[x * 5 + 4] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f;
                r4 = R0[x * 5 + 2];
                r5 = R0[x * 5 + 3];
                r6 = R0[x * 5 + 4] * 0.5f;
            }
            r2 = (R0[x * 5] - r2) * 0.5f;
            r3 = (R0[x * 5 + 1] - r3) * 0.5f;
            r2 += r4 * dy + r6 * dx;
            r3 += r6 * dy + r5 * dx;
            if ((unsigned)(x - BORDER) >= (unsigned)(width - BORDER * 2) ||
                (unsigned)(y - BORDER) >= (unsigned)(height - BORDER * 2)) {
                float scale = (x < BORDER ? border[x] : 1.f) * (x >= width - BORDER ? border[width - x - 1] : 1.f) *
                              (y < BORDER ? border[y] : 1.f) * (y >= height - BORDER ? border[height - y - 1] : 1.f);
                r2 *= scale; r3 *= scale; r4 *= scale; r5 *= scale; r6 *= scale;
            }
            M[x * 5] = r4 * r4 + r6 * r6;
            M[x * 5 + 1] = (r4 + r5) * r6;
            M[x * 5 + 2] = r5 * r5 + r6 * r6;
            M[x * 5 + 3] = r4 * r2 + r6 * r3;
            M[x * 5 + 4] = r6 * r2 + r5 * r3;
        }
    }
}

/* FarnebackUpdateFlow_Blur, written as the clean two-phase form (box sums in f64, replicate border,
 * then the 2x2 solve); OpenCV's striped in-loop UpdateMatrices is equivalent. */
static void update_flow_blur(const float* M, float* flow, int width, int height, int block_size)
{
    const int m = block_size / 2;
    const double scale = 1. / (block_size * block_size);
    double* vs = (double*)malloc((size_t)width * 5 * sizeof(double));
    for (int y = 0; y < height; y++) {
        for (int x = 0; x < width * 5; x++) vs[x] = 0.0;
        for (int dy = -m; dy <= m; dy++) {
            int yy = y + dy;
            yy = yy < 0 ? 0 : (yy > height - 1 ? height - 1 : yy);
            const float* s = M + (size_t)yy * width * 5;
            for (int x = 0; x < width * 5; x++) vs[x] += s[x];
        }
        for (int x = 0; x < width; x++) {
            double g11 = 0, g12 = 0, g22 = 0, h1 = 0, h2 = 0;
            for (int dx = -m; dx <= m; dx++) {
                int xx = x + dx;
                xx = xx < 0 ? 0 : (xx > width - 1 ? width - 1 : xx);
                g11 += vs[xx * 5];
                g12 += vs[xx * 5 + 1];
                g22 += vs[xx * 5 + 2];
                h1 += vs[xx * 5 + 3];
                h2 += vs[xx * 5 + 4];
            }
            double g11_ = g11 * scale, g12_ = g12 * scale, g22_ = g22 * scale, h1_ = h1 * scale, h2_ = h2 * scale;
            double idet = 1. / (g11_ * g22_ - g12_ * g12_ + 1e-3);
            flow[((size_t)y * width + x) * 2] = (float)((g11_ * h2_ - g12_ * h1_) * idet);
            flow[((size_t)y * width + x) * 2 + 1] = (float)((g22_ * h1_ - g12_ * h2_) * idet);
        }
    }
    free(vs);
}

static void level_geometry(int w, int h, double pyr_scale, int k, double* scale, double* sigma, int* ksize, int* lw,
                           int* lh)
{
    double s = 1;
    for (int i = 0; i < k; i++) s *= pyr_scale;
    *scale = s;
    *sigma = (1. / s - 1) * 0.5;
    int sz = cv_round(*sigma * 5) | 1;
    *ksize = sz > 3 ? sz : 3;
    *lw = cv_round(w * s);
    *lh = cv_round(h * s);
}

void gdo_farneback_polyexp_level(const uint8_t* img, int w, int h, double pyr_scale, int k, int poly_n,
                                 double poly_sigma, float* out, int* lw, int* lh)
{
    double scale, sigma;
    int ksize;
    level_geometry(w, h, pyr_scale, k, &scale, &sigma, &ksize, lw, lh);
    const size_t n = (size_t)w * h;
    float* f = (float*)malloc(n * sizeof(float));
    float* b = (float*)malloc(n * sizeof(float));
    float* I = (float*)malloc((size_t)(*lw) * (*lh) * sizeof(float));
    for (size_t i = 0; i < n; ++i) f[i] = (float)img[i];
    gaussian_blur_f32(f, w, h, ksize, sigma, b);
    resize_linear_f32(b, w, h, I, *lw, *lh, 1);
    poly_exp(I, *lw, *lh, poly_n, poly_sigma, out);
    free(f);
    free(b);
    free(I);
}

void gdo_farneback(const uint8_t* prev, const uint8_t* next, int w, int h, double pyr_scale, int levels, int winsize,
                   int iterations, int poly_n, double poly_sigma, float* flow_out)
{
    const int min_size = 32;
    int k;
    double scale = 1;
    for (k = 0; k < levels; k++) {
        scale *= pyr_scale;
        if (w * scale < min_size || h * scale < min_size) break;
    }
    levels = k;
    float* prev_flow = NULL;
    int pw = 0, ph = 0;
    for (k = levels; k >= 0; k--) {
        double sc, sigma;
        int ksize, lw, lh;
        level_geometry(w, h, pyr_scale, k, &sc, &sigma, &ksize, &lw, &lh);
        const size_t ln = (size_t)lw * lh;
        float* flow = (float*)calloc(ln * 2, sizeof(float));
        if (prev_flow) {
            resize_linear_f32(prev_flow, pw, ph, flow, lw, lh, 2);
            const float mul = (float)(1. / pyr_scale);
            for (size_t i = 0; i < ln * 2; ++i) flow[i] *= mul;
        }
        float* R0 = (float*)malloc(ln * 5 * sizeof(float));
        float* R1 = (float*)malloc(ln * 5 * sizeof(float));
        float* M = (float*)malloc(ln * 5 * sizeof(float));
        int a, b;
        gdo_farneback_polyexp_level(prev, w, h, pyr_scale, k, poly_n, poly_sigma, R0, &a, &b);
        gdo_farneback_polyexp_level(next, w, h, pyr_scale, k, poly_n, poly_sigma, R1, &a, &b);
        update_matrices(R0, R1, flow, M, lw, lh);
        for (int i = 0; i < iterations; i++) {
            update_flow_blur(M, flow, lw, lh, winsize);
            if (i < iterations - 1) update_matrices(R0, R1, flow, M, lw, lh);
        }
        free(R0);
        free(R1);
        free(M);
        free(prev_flow);
        prev_flow = flow;
        pw = lw;
        ph = lh;
    }
    memcpy(flow_out, prev_flow, (size_t)w * h * 2 * sizeof(float));
    free(prev_flow);
}

/* GetNoGMMmask for one pair with the pose given (GeoMaskMaker.cc:167-277, 405-407) */
void gdo_geomask_pair(const uint8_t* bgr_ref, const uint8_t* bgr_cur, const float* depth_ref, const float* depth_cur,
                      int w, int h, const float* K, const float* R, const float* T, uint8_t* mask, float* flow_out,
                      float* dist_out)
{
    const size_t n = (size_t)w * h;
    uint8_t* g0 = (uint8_t*)malloc(n);
    uint8_t* g1 = (uint8_t*)malloc(n);
    uint8_t* e0 = (uint8_t*)malloc(n);
    uint8_t* e1 = (uint8_t*)malloc(n);
    float* flow = (float*)malloc(n * 2 * sizeof(float));
    float* dist = (float*)malloc(n * sizeof(float));
    gdo_gray_u8(bgr_ref, (size_t)w * 3, w, h, 0, g0);
    gdo_gray_u8(bgr_cur, (size_t)w * 3, w, h, 0, g1);
    gdo_farneback(g0, g1, w, h, 0.5, 3, 15, 3, 5, 1.2, flow);
    gdo_depth_edge(depth_ref, w, h, K, e0);
    gdo_depth_edge(depth_cur, w, h, K, e1);
    gdo_mahalanobis(flow, depth_ref, depth_cur, e0, e1, NULL, w, h, K, R, T, dist, NULL, NULL);
    gdo_normalize_threshold(dist, w, h, mask, NULL, NULL);
    if (flow_out) memcpy(flow_out, flow, n * 2 * sizeof(float));
    if (dist_out) memcpy(dist_out, dist, n * sizeof(float));
    free(g0); free(g1); free(e0); free(e1); free(flow); free(dist);
}
