/*
 * geomask_oracle.c — CPU restatement of GeoMaskMaker's active path (TEST INFRASTRUCTURE, see gd_oracle.h).
 *
 * Follows /root/reference/src/GeoMaskMaker.cc:
 *   GetFlow gray conversion        :158-166   -> gdo_gray_u8
 *   GetEdge                        :854-964   -> gdo_depth_edge
 *   GetNoGMMmask per-pixel loop    :190-272   -> gdo_mahalanobis
 *   normalize / convertTo / <20    :276-277, 405-407 -> gdo_normalize_threshold
 *   depth2std                      :1386-1391 -> gdo_depth2std
 * OpenCV arithmetic pinned by probes against cv2 4.13 (tests/golden/make_golden.py):
 *   - Mat::inv() 3x3: closed form, determinant and cofactors in f64, cast to the Mat type;
 *   - gemm with flags==0 and inner length 3 (3x3*3x1, 3x3*3x3): f32 left-to-right accumulation;
 *   - every other gemm on the path (3x6*6x6, A*Bt, At*B, 1x3*3x1): f64 accumulation, cast to f32;
 *   - scaleAdd (Mat*s + Mat) and convertTo(scale,shift): fused multiply-add in f32;
 *   - convertTo(8U): round-half-even + saturate.
 * Compile with -ffp-contract=off.
 */
#include "gd_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ---- cvtColor ---------------------------------------------------------------- */
void gdo_gray_u8(const uint8_t* src, size_t src_step, int w, int h, int order, uint8_t* dst)
{
    const int k0 = order == 0 ? 3735 : 9798; /* weight of byte 0 */
    const int k2 = order == 0 ? 9798 : 3735; /* weight of byte 2 */
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + (size_t)y * src_step;
        for (int x = 0; x < w; ++x)
            dst[(size_t)y * w + x] = (uint8_t)((s[3 * x] * k0 + s[3 * x + 1] * 19235 + s[3 * x + 2] * k2 + 16384) >> 15);
    }
}

/* ---- cv::invert 3x3 ---------------------------------------------------------------------- */
static int inv3_core(const double* m, double* t)
{
    double d = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
    if (d == 0.0) {
        for (int i = 0; i < 9; ++i) t[i] = 0.0;
        return 0;
    }
    d = 1.0 / d;
    t[0] = (m[4] * m[8] - m[5] * m[7]) * d;
    t[1] = (m[2] * m[7] - m[1] * m[8]) * d;
    t[2] = (m[1] * m[5] - m[2] * m[4]) * d;
    t[3] = (m[5] * m[6] - m[3] * m[8]) * d;
    t[4] = (m[0] * m[8] - m[2] * m[6]) * d;
    t[5] = (m[2] * m[3] - m[0] * m[5]) * d;
    t[6] = (m[3] * m[7] - m[4] * m[6]) * d;
    t[7] = (m[1] * m[6] - m[0] * m[7]) * d;
    t[8] = (m[0] * m[4] - m[1] * m[3]) * d;
    return 1;
}

int gdo_inv3_f64(const double* m, double* out) { return inv3_core(m, out); }

int gdo_inv3_f32(const float* m, float* out)
{
    double md[9], t[9];
    for (int i = 0; i < 9; ++i) md[i] = (double)m[i];
    int ok = inv3_core(md, t);
    for (int i = 0; i < 9; ++i) out[i] = (float)t[i];
    return ok;
}

/* ---- GetEdge (GeoMaskMaker.cc:854-964) ----------------------------------------------------
 * The reference sweeps column-major and clamps d>3.5 -> 0 in place for interior pixels only;
 * every read of a neighbour (y-1,x) / (y,x-1) happens after that neighbour was clamped (or it is a
 * border pixel, never clamped) => identical to working on a pre-clamped map (SURVEY B-8). */
void gdo_depth_edge(const float* depth, int w, int h, const float* K, uint8_t* edge)
{
    const size_t n = (size_t)w * h;
    double* d = (double*)malloc(n * sizeof(double));
    double* nrm = (double*)calloc(n * 3, sizeof(double));
    double* vtx = (double*)calloc(n * 3, sizeof(double));
    double Kd[9], Ki[9];
    for (int i = 0; i < 9; ++i) Kd[i] = (double)K[i];
    gdo_inv3_f64(Kd, Ki); /* inst_param_d.inv() :888 */
    for (size_t i = 0; i < n; ++i) d[i] = (double)depth[i];
    for (int y = 1; y < h - 1; ++y)
        for (int x = 1; x < w - 1; ++x)
            if (d[(size_t)y * w + x] > 3.5) d[(size_t)y * w + x] = 0.0; /* :870-874 */

    for (int y = 1; y < h - 1; ++y) {
        for (int x = 1; x < w - 1; ++x) {
            const size_t i = (size_t)y * w + x;
            const double dc = d[i], dt = d[i - w], dl = d[i - 1];
            if (dc == 0.0 || dt == 0.0 || dl == 0.0) continue; /* :870-878 (clamped centre also skips) */
            /* (l-c) x (t-c) with l-c=(-1,0,dl-dc), t-c=(0,-1,dt-dc)  :879-882 */
            const double a0 = -1.0, a1 = 0.0, a2 = dl - dc;
            const double b0 = 0.0, b1 = -1.0, b2 = dt - dc;
            const double c0 = a1 * b2 - a2 * b1;
            const double c1 = a2 * b0 - a0 * b2;
            const double c2 = a0 * b1 - a1 * b0;
            /* cv::normalize(Vec3d): v * (1/norm), norm = sqrt(((0+x*x)+y*y)+z*z)  :883 */
            double s = 0.0;
            s += c0 * c0;
            s += c1 * c1;
            s += c2 * c2;
            const double nv = sqrt(s);
            const double inv = nv != 0.0 ? 1.0 / nv : 0.0;
            nrm[3 * i + 0] = c0 * inv;
            nrm[3 * i + 1] = c1 * inv;
            nrm[3 * i + 2] = c2 * inv;
            /* inst_param_d.inv() * (x,y,1) then * depth  :886-891 ; f64 gemm, inner length 3, left-to-right */
            const double px = (double)x, py = (double)y;
            const double h0 = Ki[0] * px + Ki[1] * py + Ki[2] * 1.0;
            const double h1 = Ki[3] * px + Ki[4] * py + Ki[5] * 1.0;
            const double h2 = Ki[6] * px + Ki[7] * py + Ki[8] * 1.0;
            vtx[3 * i + 0] = h0 * dc;
            vtx[3 * i + 1] = h1 * dc;
            vtx[3 * i + 2] = h2 * dc;
        }
    }
    static const int nx[8] = {-1, -1, 0, 1, 1, 1, 0, -1}; /* :894-895 */
    static const int ny[8] = {0, -1, -1, -1, 0, 1, 1, 1};
    memset(edge, 0, n);
    for (int y = 1; y < h - 1; ++y) {
        for (int x = 1; x < w - 1; ++x) {
            const size_t i = (size_t)y * w + x;
            if (d[i] == 0.0) continue; /* :901 */
            int zero_nb = 0;
            double max_phi_d = -1.0, max_phi_c = -1.0;
            for (int k = 0; k < 8; ++k) {
                const size_t j = (size_t)(y + ny[k]) * w + (x + nx[k]);
                if (vtx[3 * j + 2] == 0.0) { /* :913-922 */
                    zero_nb = 1;
                    continue;
                }
                /* Matx<1,3>*Vec3: s=0; s+=a_k*b_k  :923-924 */
                double phi_d = 0.0;
                phi_d += (vtx[3 * j + 0] - vtx[3 * i + 0]) * nrm[3 * i + 0];
                phi_d += (vtx[3 * j + 1] - vtx[3 * i + 1]) * nrm[3 * i + 1];
                phi_d += (vtx[3 * j + 2] - vtx[3 * i + 2]) * nrm[3 * i + 2];
                if (max_phi_d < fabs(phi_d)) max_phi_d = fabs(phi_d);
                double phi_c = 0.0;
                if (phi_d < 0.0) {
                    if (max_phi_c < phi_c) max_phi_c = phi_c;
                } else {
                    double dot = 0.0;
                    dot += nrm[3 * j + 0] * nrm[3 * i + 0];
                    dot += nrm[3 * j + 1] * nrm[3 * i + 1];
                    dot += nrm[3 * j + 2] * nrm[3 * i + 2];
                    phi_c = 1.0 - dot;
                    if (phi_c > max_phi_c) max_phi_c = phi_c;
                }
            }
            if (zero_nb) { /* :946-950 */
                edge[i] = 255;
                continue;
            }
            if (max_phi_c == -1.0 || max_phi_d == -1.0) continue; /* :951 (unreachable: 8 zero neighbours set the flag) */
            const double thres_edge = max_phi_d + 0.05 * max_phi_c; /* :955 */
            if (thres_edge > 0.04) edge[i] = 255;
        }
    }
    free(d);
    free(nrm);
    free(vtx);
}

/* ---- depth2std (GeoMaskMaker.cc:1386-1391): left-to-right f32 products ---------------------- */
float gdo_depth2std(float depth, float fu)
{
    const float sigma_norm = 0.5f;
    const float inv = 1 / fu;
    float r = inv * inv;
    r = r * sigma_norm;
    r = r * sigma_norm;
    r = r * depth;
    r = r * depth;
    r = r * depth;
    r = r * depth;
    return r;
}

/* f32 gemm, inner length 3, flags==0: t = a0*b0 + a1*b1 + a2*b2 in f32, left to right */
static inline float dot3f(float a0, float b0, float a1, float b1, float a2, float b2)
{
    float t = a0 * b0;
    t = t + a1 * b1;
    t = t + a2 * b2;
    return t;
}

/* ---- GetNoGMMmask main loop (GeoMaskMaker.cc:190-272) ------------------------------------ */
void gdo_mahalanobis(const float* flow, const float* depth_ref, const float* depth_cur, const uint8_t* edge_ref,
                     const uint8_t* edge_cur, const float* lut, int w, int h, const float* K, const float* R,
                     const float* T, float* dist, uint8_t* written, int32_t* src_index)
{
    const size_t n = (size_t)w * h;
    memset(dist, 0, n * sizeof(float));
    if (written) memset(written, 0, n);
    if (src_index)
        for (size_t i = 0; i < n; ++i) src_index[i] = -1;

    const float fu = K[0], fv = K[4], cu = K[2]; /* :44-47 (cv unused by the reference loop) */
    float Ki[9];
    gdo_inv3_f32(K, Ki); /* :200 */
    /* R*inv_inst_param is evaluated to a temporary first (MatExpr left-assoc), f32 3x3*3x3 */
    float RK[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            RK[3 * i + j] = dot3f(R[3 * i], Ki[j], R[3 * i + 1], Ki[3 + j], R[3 * i + 2], Ki[6 + j]);

    for (int y = 0; y < h; ++y) {
        for (int x = 0; x < w; ++x) {
            const size_t i = (size_t)y * w + x;
            const float cur_x = (float)x + flow[2 * i];     /* :213 */
            const float cur_y = (float)y + flow[2 * i + 1]; /* :214 */
            if (!(cur_x == cur_x) || !(cur_y == cur_y)) continue; /* NaN: UB in the reference (SURVEY B-2) -> skip */
            if (cur_x < 0 || cur_y < 0 || cur_x > (float)(w - 1) || cur_y > (float)(h - 1)) continue; /* :215 */
            const int icx = (int)cur_x, icy = (int)cur_y;
            float rx, ry, cx, cy;
            if (lut) { /* :219-220 */
                rx = lut[2 * i];
                ry = lut[2 * i + 1];
                cx = lut[2 * ((size_t)icy * w + icx)];
                cy = lut[2 * ((size_t)icy * w + icx) + 1];
            } else {
                rx = (float)x;
                ry = (float)y;
                cx = (float)icx;
                cy = (float)icy;
            }
            const int rix = (int)rx, riy = (int)ry, cix = (int)cx, ciy = (int)cy;
            /* the reference indexes without a guard; out-of-image LUT entries are UB there -> skip */
            if (rix < 0 || riy < 0 || rix >= w || riy >= h || cix < 0 || ciy < 0 || cix >= w || ciy >= h) continue;
            const float ref_depth = depth_ref[(size_t)riy * w + rix]; /* :222 */
            const float cur_depth = depth_cur[(size_t)ciy * w + cix]; /* :223 */
            if (edge_ref[(size_t)riy * w + rix] == 255 || edge_cur[(size_t)ciy * w + cix] == 255) continue; /* :224-228 */
            if (cur_depth == 0 || (double)cur_depth > 3.5 || ref_depth == 0 || (double)ref_depth > 3.5) continue; /* :229 */

            /* U_t = (R*Kinv)*homoRefPixel  :241  (f32, inner 3) */
            const float U0 = dot3f(RK[0], rx, RK[1], ry, RK[2], 1.0f);
            const float U1 = dot3f(RK[3], rx, RK[4], ry, RK[5], 1.0f);
            const float U2 = dot3f(RK[6], rx, RK[7], ry, RK[8], 1.0f);
            /* CurPoint3D = gemm(Kinv, homoCur, alpha=cur_depth)  :242 ; (float)(t*alpha) == f32 product */
            const float C0 = dot3f(Ki[0], cx, Ki[1], cy, Ki[2], 1.0f) * cur_depth;
            const float C1 = dot3f(Ki[3], cx, Ki[4], cy, Ki[5], 1.0f) * cur_depth;
            const float C2 = dot3f(Ki[6], cx, Ki[7], cy, Ki[8], 1.0f) * cur_depth;
            /* RefPoint3D = scaleAdd(U_t, ref_depth, T)  :243 ; fused in OpenCV 4.13 */
            const float P0 = fmaf(U0, ref_depth, T[0]);
            const float P1 = fmaf(U1, ref_depth, T[1]);
            const float P2 = fmaf(U2, ref_depth, T[2]);
            const float e0 = C0 - P0, e1 = C1 - P1, e2 = C2 - P2; /* :244 */

            /* S diag (1,1,s_ref,1,1,s_cur)  :246-247 */
            const float s2 = gdo_depth2std(ref_depth, fu);
            const float s5 = gdo_depth2std(cur_depth, fu);
            /* J (3x6) :250-265, index slips kept verbatim */
            float J[3][6];
            memset(J, 0, sizeof(J));
            J[0][0] = cur_depth / fu;
            J[0][2] = (cx - cu) / fu;
            J[0][3] = -R[0] * ref_depth / fu;
            J[0][4] = -R[1] * ref_depth / fv;
            J[0][5] = -U0;
            J[1][1] = ref_depth / fv;
            J[1][2] = (cx - cu) / fv;
            J[1][3] = -R[3] * ref_depth / fu;
            J[1][4] = -R[4] * ref_depth / fv;
            J[1][5] = -U1;
            J[2][2] = 1.0f;
            J[2][3] = -R[6] * ref_depth / fu;
            J[2][4] = -R[7] * ref_depth / fv;
            J[2][5] = -U2;
            /* J*S : f64-accumulated gemm; S diagonal => (float)((double)J*(double)S) == f32 product */
            float JS[3][6];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 6; ++c) JS[r][c] = J[r][c];
                JS[r][2] = J[r][2] * s2;
                JS[r][5] = J[r][5] * s5;
            }
            /* (J*S)*J^T : f64 accumulation over k=0..5, cast to f32 */
            float Cm[9];
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 3; ++c) {
                    double s = 0.0;
                    for (int k = 0; k < 6; ++k) s += (double)JS[r][k] * (double)J[c][k];
                    Cm[3 * r + c] = (float)s;
                }
            float Ci[9];
            gdo_inv3_f32(Cm, Ci); /* .inv()  :267 */
            /* dist^T * Cinv : f64 accumulation -> f32 (1x3) */
            float q[3];
            for (int c = 0; c < 3; ++c) {
                double s = 0.0;
                s += (double)e0 * (double)Ci[c];
                s += (double)e1 * (double)Ci[3 + c];
                s += (double)e2 * (double)Ci[6 + c];
                q[c] = (float)s;
            }
            double l = 0.0;
            l += (double)q[0] * (double)e0;
            l += (double)q[1] * (double)e1;
            l += (double)q[2] * (double)e2;
            const float lik = (float)l;
            const float value = sqrtf(lik); /* :268 */
            const size_t t = (size_t)icy * w + icx;
            dist[t] = value; /* :269 scatter, last raster-order writer wins */
            if (written) written[t] = 1;
            if (src_index) src_index[t] = (int32_t)i;
        }
    }
}

/* ---- normalize + convertTo + threshold (GeoMaskMaker.cc:276-277,405-407) ------- */
void gdo_normalize_threshold(const float* dist, int w, int h, uint8_t* mask, uint8_t* d8, float* minmax)
{
    const size_t n = (size_t)w * h;
    float mn = dist[0], mx = dist[0];
    for (size_t i = 1; i < n; ++i) {
        if (dist[i] < mn) mn = dist[i];
        if (dist[i] > mx) mx = dist[i];
    }
    if (minmax) {
        minmax[0] = mn;
        minmax[1] = mx;
    }
    const double smin = (double)mn, smax = (double)mx;
    const double scale = 255.0 * (smax - smin > 2.220446049250313e-16 ? 1.0 / (smax - smin) : 0.0);
    const double shift = 0.0 - smin * scale;
    const float a = (float)scale, b = (float)shift;
    for (size_t i = 0; i < n; ++i) {
        const float v = fmaf(dist[i], a, b); /* convertTo(scale,shift) f32, fused */
        /* convertTo(CV_8U): cvRound (round half even) + saturate */
        long r = lrintf(v);
        if (!(v == v)) r = 0;
        if (r < 0) r = 0;
        if (r > 255) r = 255;
        if (d8) d8[i] = (uint8_t)r;
        mask[i] = (uint8_t)(r < 20 ? 1 : 0); /* (dist<20)/255  :405-406 */
    }
}

/* ---- "next" row (f)-2: Frame ctor mask erosion + keypoint filter (src/Frame.cc:258-282) ----------------
 * cv::getStructuringElement(MORPH_ELLIPSE, 31x31, anchor (15,15)) + cv::erode with the default border (pixels outside
 * the image do not take part in the minimum), then keep keypoint i iff eroded((int)pt.y, (int)pt.x) == 1. */
void gdo_ellipse31(int* j1, int* j2)
{
    const int r = 15, c = 15;
    const double inv_r2 = 1.0 / ((double)r * r);
    for (int i = 0; i < 31; ++i) {
        const int dy = i - r;
        const int dx = (int)lrint(c * sqrt((r * r - dy * dy) * inv_r2));
        j1[i] = c - dx > 0 ? c - dx : 0;
        j2[i] = c + dx + 1 < 31 ? c + dx + 1 : 31; /* exclusive */
    }
}

void gdo_erode31(const uint8_t* mask, int w, int h, uint8_t* out)
{
    int j1[31], j2[31];
    gdo_ellipse31(j1, j2);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            int mn = 255;
            for (int i = 0; i < 31; ++i) {
                const int yy = y + i - 15;
                if (yy < 0 || yy >= h) continue;
                for (int j = j1[i]; j < j2[i]; ++j) {
                    const int xx = x + j - 15;
                    if (xx < 0 || xx >= w) continue;
                    const int v = mask[(size_t)yy * w + xx];
                    if (v < mn) mn = v;
                }
            }
            out[(size_t)y * w + x] = (uint8_t)mn;
        }
}

/* kps: n records of 7 x 4 bytes (x, y, size, angle, response, octave, class_id); keep: n flags */
int gdo_erode_filter(const uint8_t* mask, int w, int h, const float* kps, int n, uint8_t* keep)
{
    int j1[31], j2[31], kept = 0;
    gdo_ellipse31(j1, j2);
    for (int k = 0; k < n; ++k) {
        const int x = (int)kps[7 * k], y = (int)kps[7 * k + 1];
        int mn = 255;
        for (int i = 0; i < 31; ++i) {
            const int yy = y + i - 15;
            if (yy < 0 || yy >= h) continue;
            for (int j = j1[i]; j < j2[i]; ++j) {
                const int xx = x + j - 15;
                if (xx < 0 || xx >= w) continue;
                const int v = mask[(size_t)yy * w + xx];
                if (v < mn) mn = v;
            }
        }
        keep[k] = (uint8_t)(mn == 1);
        kept += keep[k];
    }
    return kept;
}

/* ---- "next" row (f)-3: Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:576-636), ComputeStereoFromRGBD
 * (:815-837) and AssignFeaturesToGrid / PosInGrid (:402-417, :553-565).  K: 3x3 f32; D: k1 k2 p1 p2 k3 (f32, NULL = none).
 * The reference tests mDistCoef(0) == 0 to decide whether anything is distorted (:578, :610).
 * cell = col * 48 + row (mGrid[col][row]); cell_start has 64*48+1 entries; items are keypoint indices in increasing order.
 * Pinned by tests/golden/stereo_grid.npz (literal Python transcription with cv2.undistortPoints). */
void gdo_undistort_point(const float* K, const float* D, float u, float v, float* ou, float* ov)
{
    /* cv::undistortPoints(src, dst, K, D, noArray(), K): f64, 5 fixed-point iterations (default criteria), then K again */
    const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
    const double ifx = 1.0 / fx, ify = 1.0 / fy;
    double k[5] = {D[0], D[1], D[2], D[3], D[4]};
    double x = ((double)u - cx) * ifx, y = ((double)v - cy) * ify;
    const double x0 = x, y0 = y;
    for (int it = 0; it < 5; ++it) {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + ((k[4] * r2 + k[1]) * r2 + k[0]) * r2);
        if (icdist < 0) {
            x = ((double)u - cx) * ifx;
            y = ((double)v - cy) * ify;
            break;
        }
        const double dX = 2 * k[2] * x * y + k[3] * (r2 + 2 * x * x);
        const double dY = k[2] * (r2 + 2 * y * y) + 2 * k[3] * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ou = (float)(fx * x + cx);
    *ov = (float)(fy * y + cy);
}

void gdo_stereo_grid(const float* depth_m, int w, int h, const float* kps, int n, float bf, const float* K, const float* D,
                     float* depth_out, float* uright, int* cell_start, int* cell_items, float* un_out, float* bounds_out)
{
    enum { COLS = 64, ROWS = 48 };
    const int distorted = D && D[0] != 0.0f;
    float minx = 0.f, maxx = (float)w, miny = 0.f, maxy = (float)h;
    if (distorted) { /* ComputeImageBounds: the four undistorted image corners */
        float c[4][2];
        gdo_undistort_point(K, D, 0.f, 0.f, &c[0][0], &c[0][1]);
        gdo_undistort_point(K, D, (float)w, 0.f, &c[1][0], &c[1][1]);
        gdo_undistort_point(K, D, 0.f, (float)h, &c[2][0], &c[2][1]);
        gdo_undistort_point(K, D, (float)w, (float)h, &c[3][0], &c[3][1]);
        minx = c[0][0] < c[2][0] ? c[0][0] : c[2][0];
        maxx = c[1][0] > c[3][0] ? c[1][0] : c[3][0];
        miny = c[0][1] < c[1][1] ? c[0][1] : c[1][1];
        maxy = c[2][1] > c[3][1] ? c[2][1] : c[3][1];
    }
    if (bounds_out) {
        bounds_out[0] = minx; bounds_out[1] = maxx; bounds_out[2] = miny; bounds_out[3] = maxy;
    }
    const float inv_w = (float)COLS / (maxx - minx), inv_h = (float)ROWS / (maxy - miny);
    int* cell = (int*)malloc((size_t)(n > 0 ? n : 1) * sizeof(int));
    for (int i = 0; i < n; ++i) {
        const float u = kps[7 * i], v = kps[7 * i + 1];
        float uu = u, vu = v; /* mvKeysUn */
        if (distorted) gdo_undistort_point(K, D, u, v, &uu, &vu);
        if (un_out) {
            un_out[2 * i] = uu;
            un_out[2 * i + 1] = vu;
        }
        const float d = depth_m[(size_t)(int)v * w + (int)u]; /* imDepth.at<float>(v,u) at the DISTORTED keypoint */
        depth_out[i] = -1.f;
        uright[i] = -1.f;
        if (d > 0) {
            depth_out[i] = d;
            uright[i] = uu - bf / d;
        }
        const int px = (int)roundf((uu - minx) * inv_w), py = (int)roundf((vu - miny) * inv_h);
        cell[i] = (px < 0 || px >= COLS || py < 0 || py >= ROWS) ? -1 : px * ROWS + py;
    }
    int pos = 0;
    for (int c = 0; c < COLS * ROWS; ++c) {
        cell_start[c] = pos;
        for (int i = 0; i < n; ++i)
            if (cell[i] == c) cell_items[pos++] = i;
    }
    cell_start[COLS * ROWS] = pos;
    free(cell);
}
