"""Numpy restatement of the feature / matching front half of GeoMaskMaker::GetRt (GD-SLAM src/GeoMaskMaker.cc:77-141) —
SURVEY 8(f)-1, the next row of the scope table.  TEST INFRASTRUCTURE ONLY (like everything under oracle/): no product
path exists for this row yet; this file records the OpenCV 4.13 semantics that a GPU implementation has to reproduce,
each pinned against cv2 by tests/test_oracle_getrt.py.

    cv::ORB::create(2000, 1.2, 8, 31, 0, 2)->detectAndCompute     GeoMaskMaker.cc:82-90
    BFMatcher(NORM_HAMMING, crossCheck = true)->match              :92-94
    sort + first 100, undistortPoints, depth lookup, back-projection :95-141

Findings (probed against cv2 4.13.0, bit-exact unless stated):
  * level l has scale (float)pow((double)1.2f, l) and size cvRound(cols / scale) x cvRound(rows / scale); every level is
    resized from the PREVIOUS level with INTER_LINEAR_EXACT = 8.8 fixed-point weights round(frac * 256), horizontal then
    vertical, one final (v + 32768) >> 16;
  * per level: cv::FAST(20, nonmax) on the whole level, keep 31 <= x < w - 31 (same for y), retainBest(2 N_l) on the FAST
    response (everything tied with the N-th value is kept), Harris response (7x7 block of Sobel-like sums in int, then
    (a b - c^2 - 0.04 (a + b)^2) * scale^4 in individually rounded f32, scale = 1 / (4 * 7 * 255)), retainBest(N_l);
    N_l = cvRound of the geometric series 2000 (1 - f) / (1 - f^8), f = 1 / 1.2f, last level takes the remainder;
  * angle = IC_Angle exactly as ORB_SLAM2's copy (fastAtan2 of the integer patch moments), pt *= scale afterwards;
  * descriptors read a level blurred with the FLOAT separable Gaussian (7 taps, sigma 2, REFLECT_101, cvRound at the end):
    cv::ORB blurs a submatrix of its pyramid buffer, and cv::GaussianBlur only takes the 8-bit fixed-point path
    ([18 34 48 56 48 34 18] / 256, what ORBextractor's blur of a cloned level gets) for non-submatrix inputs — 2.5 % of the
    pixels differ by one grey level between the two; the pattern and the rotation arithmetic equal ORB_SLAM2's;
  * the ORDER of the returned keypoints is the permutation std::nth_element + std::partition leave behind inside the two
    retainBest calls, applied to the raster-ordered FAST output: calling the same libstdc++ routines on the same sequence
    (oracle gdo_retain_best_order) reproduces cv2's order exactly;
  * BFMatcher cross-check: query q pairs with its nearest train t (smallest index on ties) iff q is the nearest query of
    t (smallest index on ties); results come ordered by query index;
  * sort(matches) is an unstable std::sort on the distance: which of several equal-distance matches make the first 100 is
    implementation defined; oracle gdo_sort_matches_order runs libstdc++'s std::sort on the same sequence
    (GeoMaskMaker.cc:96-97 also reads past the end when fewer than 100 matches exist);
  * solvePnPRansac + Rodrigues (:143-150) have no restatement: cv2 itself is the oracle for that step (fixed-seed RNG, EPnP
    on 5-point samples, LM refinement on the inliers).
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def cv_round(x) -> int:
    return int(np.rint(x))


def resize_linear_exact(src: np.ndarray, dw: int, dh: int) -> np.ndarray:
    sh, sw = src.shape

    def axis(dn, sn):
        d = np.arange(dn)
        f = (d + 0.5) * (sn / dn) - 0.5
        i = np.floor(f).astype(np.int64)
        fr = f - i
        lo, hi = i < 0, i >= sn - 1
        i[lo], fr[lo] = 0, 0.0
        i[hi], fr[hi] = sn - 1, 0.0
        return i, np.floor(fr * 256 + 0.5).astype(np.int64)

    xi, xa = axis(dw, sw)
    yi, ya = axis(dh, sh)
    xi1, yi1 = np.minimum(xi + 1, sw - 1), np.minimum(yi + 1, sh - 1)
    s = src.astype(np.int64)
    h = s[:, xi] * (256 - xa) + s[:, xi1] * xa
    v = h[yi, :] * (256 - ya)[:, None] + h[yi1, :] * ya[:, None]
    return ((v + 32768) >> 16).clip(0, 255).astype(np.uint8)


def gaussian7_float(im: np.ndarray) -> np.ndarray:
    """cv::GaussianBlur(7x7, sigma 2, REFLECT_101) on a SUBMATRIX: float separable filter, symmetric form, cvRound."""
    x = np.arange(7) - 3
    g = np.exp(-(x * x) / (2.0 * 2.0 * 2.0))
    g = (g / g.sum()).astype(f32)
    h, w = im.shape
    p = np.pad(im, 3, mode="reflect").astype(f32)
    row = g[3] * p[:, 3:3 + w]
    for k in (1, 2, 3):
        row = row + g[3 + k] * (p[:, 3 + k:3 + k + w] + p[:, 3 - k:3 - k + w])
    col = g[3] * row[3:3 + h, :]
    for k in (1, 2, 3):
        col = col + g[3 + k] * (row[3 + k:3 + k + h, :] + row[3 - k:3 - k + h, :])
    return np.rint(col).clip(0, 255).astype(np.uint8)


def retain_best(resp: np.ndarray, n: int) -> np.ndarray:
    """KeyPointsFilter::retainBest as a set: indices with response >= the n-th largest response."""
    if len(resp) <= n:
        return np.arange(len(resp))
    if n <= 0:
        return np.arange(0)
    thr = np.sort(resp)[::-1][n - 1]
    return np.nonzero(resp >= thr)[0]


def harris_responses(img: np.ndarray, xs, ys) -> np.ndarray:
    bs, r = 7, 3
    scale = f32(1.0) / f32((1 << 2) * bs * 255.0)
    s4 = f32(f32(f32(scale * scale) * scale) * scale)
    I = img.astype(np.int64)
    out = np.zeros(len(xs), f32)
    for n, (x0, y0) in enumerate(zip(xs, ys)):
        P = I[y0 - r - 1:y0 + r + 2, x0 - r - 1:x0 + r + 2]
        Ix = (P[1:-1, 2:] - P[1:-1, :-2]) * 2 + (P[:-2, 2:] - P[:-2, :-2]) + (P[2:, 2:] - P[2:, :-2])
        Iy = (P[2:, 1:-1] - P[:-2, 1:-1]) * 2 + (P[2:, :-2] - P[:-2, :-2]) + (P[2:, 2:] - P[:-2, 2:])
        fa, fb, fc = f32(int((Ix * Ix).sum())), f32(int((Iy * Iy).sum())), f32(int((Ix * Iy).sum()))
        t = f32(f32(fa * fb) - f32(fc * fc))
        sab = f32(fa + fb)
        out[n] = f32(f32(t - f32(f32(f32(0.04) * sab) * sab)) * s4)
    return out


def features_per_level(nfeatures: int, scale_factor: float, nlevels: int):
    factor = 1.0 / float(f32(scale_factor))
    nd = nfeatures * (1 - factor) / (1 - factor ** nlevels)
    out = []
    for _ in range(nlevels - 1):
        out.append(cv_round(nd))
        nd *= factor
    out.append(max(nfeatures - sum(out), 0))
    return out


def cv_orb_detect_and_compute(img, fast_detect, ic_angle, orb_descriptor, nfeatures=2000, scale_factor=1.2, nlevels=8, edge=31,
                              retain_best_order=None):
    """cv::ORB (HARRIS_SCORE, WTA_K 2, patch 31, FAST 20): list of (octave, x, y, response, angle, descriptor).
    fast_detect(img, th) -> [(x, y, response)] in raster order, ic_angle(img, x, y), orb_descriptor(blurred, x, y, angle):
    the cv2-pinned primitives of oracle/pyoracle.py.  With retain_best_order = pyoracle.retain_best_order the list has
    cv2's own order, otherwise only the set is meaningful."""
    rb = retain_best_order if retain_best_order is not None else (lambda r, n: retain_best(np.asarray(r), n))
    nper = features_per_level(nfeatures, scale_factor, nlevels)
    levels, scales = [img], [f32(1.0)]
    for l in range(1, nlevels):
        sc = f32(np.power(float(f32(scale_factor)), float(l)))
        scales.append(sc)
        levels.append(resize_linear_exact(levels[-1], cv_round(img.shape[1] / sc), cv_round(img.shape[0] / sc)))
    out = []
    for l, im in enumerate(levels):
        h, w = im.shape
        if min(h, w) <= 2 * edge:
            continue
        pts = np.asarray(fast_detect(im, 20), np.float32).reshape(-1, 3)
        keep = (pts[:, 0] >= edge) & (pts[:, 0] < w - edge) & (pts[:, 1] >= edge) & (pts[:, 1] < h - edge)
        pts = pts[keep]
        pts = pts[rb(pts[:, 2], 2 * nper[l])]
        hr = harris_responses(im, pts[:, 0].astype(int), pts[:, 1].astype(int))
        blurred = gaussian7_float(im)
        for i in rb(hr, nper[l]):
            x, y = int(pts[i, 0]), int(pts[i, 1])
            ang = f32(ic_angle(im, x, y))
            out.append((l, f32(f32(x) * scales[l]), f32(f32(y) * scales[l]), hr[i], ang, orb_descriptor(blurred, x, y, ang)))
    return out


def bf_match_hamming_crosscheck(d1: np.ndarray, d2: np.ndarray):
    """BFMatcher(NORM_HAMMING, crossCheck=True).match(d1, d2): [(queryIdx, trainIdx, distance)] ordered by queryIdx."""
    lut = np.array([bin(i).count("1") for i in range(256)], np.int32)
    dist = lut[d1[:, None, :] ^ d2[None, :, :]].sum(-1)
    nn12 = dist.argmin(1)   # first minimum = smallest train index on ties
    nn21 = dist.argmin(0)
    return [(q, int(t), int(dist[q, t])) for q, t in enumerate(nn12) if nn21[t] == q]


def back_project_matches(matches, kp1_xy, kp2_xy, depth1, K, ntop=100, sort_order=None, undistort=None):
    """GeoMaskMaker.cc:95-141: the `ntop` best matches (sort_order = pyoracle.sort_matches_order gives the reference's
    std::sort order, otherwise a stable order), undistortPoints on the first image's points (undistort: callable on an
    (n, 2) f32 array = cv2.undistortPoints(pts, K, D, None, K); None = undistorted camera, where it is the identity in f32),
    depth of the first image at the truncated pixel, K^-1 [x y 1] * d in f32."""
    if sort_order is not None:
        order = [int(i) for i in sort_order(np.asarray([m[2] for m in matches], np.float32))][:ntop]
    else:
        order = sorted(range(len(matches)), key=lambda i: (matches[i][2], i))[:ntop]
    Ki = np.linalg.inv(K.astype(np.float64)).astype(f32)  # cv::Mat::inv of a 3x3 f32: f64 cofactors (oracle gdo_inv3_f32)
    obj, pix = [], []
    pts = np.array([[kp1_xy[matches[i][0]][0], kp1_xy[matches[i][0]][1]] for i in order], f32).reshape(-1, 2)
    if undistort is not None and len(pts):
        pts = np.asarray(undistort(pts), f32).reshape(-1, 2)
    for r, i in enumerate(order):
        q, t, _ = matches[i]
        x, y = f32(pts[r][0]), f32(pts[r][1])
        if not (0 <= int(x) < depth1.shape[1] and 0 <= int(y) < depth1.shape[0]):
            continue  # outside the image after undistortion: the reference would read out of bounds
        d = depth1[int(y), int(x)]
        if d == 0:
            continue
        v = np.array([x, y, f32(1.0)], f32)
        P = np.array([f32(f32(f32(Ki[r, 0] * v[0]) + f32(Ki[r, 1] * v[1])) + f32(Ki[r, 2] * v[2])) for r in range(3)], f32)
        obj.append(P * f32(d))
        pix.append(kp2_xy[t])
    return np.asarray(obj, f32).reshape(-1, 3), np.asarray(pix, f32).reshape(-1, 2)


def get_rt(img1, img2, depth1, K, prims, solve_pnp_ransac):
    """GeoMaskMaker::GetRt (GeoMaskMaker.cc:77-156) for an undistorted camera: (ok, R 3x3 f32, T 3 f32).
    prims: oracle.pyoracle (fast_detect in raster order via cv2-pinned gdo_fast_detect, ic_angle, orb_descriptor,
    retain_best_order, sort_matches_order); solve_pnp_ransac(obj, pix, K) -> (rvec, tvec) = cv2.solvePnPRansac (the oracle
    of that step) followed by cv2.Rodrigues."""
    def fast(im, th):
        return prims.fast_detect(im, th)
    feats = [cv_orb_detect_and_compute(im, fast, prims.ic_angle, prims.orb_descriptor, retain_best_order=prims.retain_best_order)
             for im in (img1, img2)]
    desc = [np.stack([f[5] for f in fs]) for fs in feats]
    xy = [[(f[1], f[2]) for f in fs] for fs in feats]
    matches = bf_match_hamming_crosscheck(desc[0], desc[1])
    obj, pix = back_project_matches(matches, xy[0], xy[1], depth1, K, 100, prims.sort_matches_order)
    if len(obj) < 20:
        return False, None, None
    R, T = solve_pnp_ransac(obj, pix, K)
    return True, np.asarray(R, np.float32), np.asarray(T, np.float32).reshape(3)
