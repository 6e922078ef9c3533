#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ (run in the build container, where cv2 4.13 is importable).

The reference (GD-SLAM) has no tests or golden vectors, and its arithmetic on this
path lives in OpenCV, which it neither vendors nor pins.  These fixtures pin the CPU oracle (oracle/) to
OpenCV 4.13.0 semantics:

  geomask_small.npz   literal, sequential, per-pixel transcription of GeoMaskMaker::GetEdge and of the
                      GetNoGMMmask loop (GeoMaskMaker.cc:190-277,405-407,854-964) in Python, calling
                      cv2.invert / cv2.gemm / cv2.scaleAdd / cv2.normalize for every Mat operation the
                      reference performs (so the accumulation widths and fused operations are OpenCV's own).
  farneback_*.npz     cv2.calcOpticalFlowFarneback(prev,next,None,0.5,3,15,3,5,1.2,0) (GeoMaskMaker.cc:165).
  prims.npz           known-answer vectors of the OpenCV primitives on the path: cvtColor (both orders),
                      fastAtan2, FAST-9/16 + NMS on random cells, resize(INTER_LINEAR, 8U) for the seven
                      pyramid transitions, GaussianBlur 7x7 sigma 2 (8U).
  orb_*.npz           (written by make_golden_orb.py)

Usage:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import importlib
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
synth = importlib.import_module("gd-slam_b200.synth")

f32 = np.float32


def literal_get_edge(depth_m: np.ndarray, K: np.ndarray) -> np.ndarray:
    """GeoMaskMaker::GetEdge, GeoMaskMaker.cc:854-964, transcribed loop by loop (column-major, in-place clamp)."""
    h, w = depth_m.shape
    md = depth_m.astype(np.float64).copy()
    Kd = K.astype(np.float64)
    normals = np.zeros((h, w, 3), np.float64)
    vertex = np.zeros((h, w, 3), np.float64)
    _, Kinv = cv2.invert(Kd)
    for x in range(1, w - 1):
        for y in range(1, h - 1):
            if md[y, x] > 3.5:
                md[y, x] = 0.0
                continue
            if md[y - 1, x] == 0.0 or md[y, x] == 0.0 or md[y, x - 1] == 0.0:
                continue
            t = np.array([x, y - 1, md[y - 1, x]], np.float64)
            l = np.array([x - 1, y, md[y, x - 1]], np.float64)
            c = np.array([x, y, md[y, x]], np.float64)
            a, b = l - c, t - c
            d = np.array([a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]])
            s = np.float64(0.0)
            for v in d:  # cv::norm(Vec3d): sequential sum of squares
                s = s + v * v
            nv = np.sqrt(s)
            n = d * (1.0 / nv if nv != 0 else 0.0)
            normals[y, x] = n
            hp = cv2.gemm(Kinv, np.array([[x], [y], [1.0]], np.float64), 1.0, None, 0.0)
            hp = hp * md[y, x]
            vertex[y, x] = hp[:, 0]
    nx = [-1, -1, 0, 1, 1, 1, 0, -1]
    ny = [0, -1, -1, -1, 0, 1, 1, 1]
    edge = np.zeros((h, w), np.uint8)
    for x in range(1, w - 1):
        for y in range(1, h - 1):
            if md[y, x] == 0.0:
                continue
            zero_nb = False
            max_phi_d = -1.0
            max_phi_c = -1.0
            for i in range(8):
                px, py = x + nx[i], y + ny[i]
                if vertex[py, px, 2] == 0:
                    zero_nb = True
                    continue
                diff = vertex[py, px] - vertex[y, x]
                phi_d = np.float64(0.0)
                for k in range(3):
                    phi_d = phi_d + diff[k] * normals[y, x, k]
                if max_phi_d < abs(phi_d):
                    max_phi_d = abs(phi_d)
                phi_c = 0.0
                if phi_d < 0:
                    if max_phi_c < phi_c:
                        max_phi_c = phi_c
                else:
                    dot = np.float64(0.0)
                    for k in range(3):
                        dot = dot + normals[py, px, k] * normals[y, x, k]
                    phi_c = 1 - dot
                    if phi_c > max_phi_c:
                        max_phi_c = phi_c
            if zero_nb:
                edge[y, x] = 255
                continue
            if max_phi_c == -1 or max_phi_d == -1:
                continue
            if max_phi_d + 0.05 * max_phi_c > 0.04:
                edge[y, x] = 255
    return edge


def literal_mahalanobis(flow, d_ref, d_cur, e_ref, e_cur, K, R, T):
    """GetNoGMMmask loop, GeoMaskMaker.cc:190-272, one cv2 call per cv::Mat operation of the reference."""
    h, w = d_ref.shape
    K = K.astype(f32)
    R = R.astype(f32).reshape(3, 3)
    T = T.astype(f32).reshape(3, 1)
    fu, fv, cu = K[0, 0], K[1, 1], K[0, 2]
    _, Kinv = cv2.invert(K)
    dist = np.zeros((h, w), f32)
    J = np.zeros((3, 6), f32)
    S = np.eye(6, dtype=f32)

    def d2s(d):
        r = f32(1) / fu
        r = f32(r * (f32(1) / fu))
        for m in (f32(0.5), f32(0.5), d, d, d, d):
            r = f32(r * m)
        return r

    RK = cv2.gemm(R, Kinv, 1.0, None, 0.0)
    for y in range(h):
        for x in range(w):
            cur_x = f32(f32(x) + flow[y, x, 0])
            cur_y = f32(f32(y) + flow[y, x, 1])
            if cur_x < 0 or cur_y < 0 or cur_x > w - 1 or cur_y > h - 1:
                continue
            icx, icy = int(cur_x), int(cur_y)
            ref_depth = d_ref[y, x]
            cur_depth = d_cur[icy, icx]
            if e_ref[y, x] == 255 or e_cur[icy, icx] == 255:
                continue
            if cur_depth == 0 or cur_depth > 3.5 or ref_depth == 0 or ref_depth > 3.5:
                continue
            hc = np.array([[icx], [icy], [1]], f32)
            hr = np.array([[x], [y], [1]], f32)
            U = cv2.gemm(RK, hr, 1.0, None, 0.0)
            Cp = cv2.gemm(Kinv, hc, float(cur_depth), None, 0.0)
            Rp = cv2.scaleAdd(U, float(ref_depth), T)
            e = (Cp - Rp).astype(f32)
            S[2, 2] = d2s(ref_depth)
            S[5, 5] = d2s(cur_depth)
            J[0, 0] = cur_depth / fu
            J[0, 2] = f32(f32(icx) - cu) / fu
            J[0, 3] = f32(-R[0, 0] * ref_depth) / fu
            J[0, 4] = f32(-R[0, 1] * ref_depth) / fv
            J[0, 5] = -U[0, 0]
            J[1, 1] = ref_depth / fv
            J[1, 2] = f32(f32(icx) - cu) / fv
            J[1, 3] = f32(-R[1, 0] * ref_depth) / fu
            J[1, 4] = f32(-R[1, 1] * ref_depth) / fv
            J[1, 5] = -U[1, 0]
            J[2, 2] = 1
            J[2, 3] = f32(-R[2, 0] * ref_depth) / fu
            J[2, 4] = f32(-R[2, 1] * ref_depth) / fv
            J[2, 5] = -U[2, 0]
            JS = cv2.gemm(J, S, 1.0, None, 0.0)
            Cm = cv2.gemm(JS, J, 1.0, None, 0.0, flags=cv2.GEMM_2_T)
            _, Ci = cv2.invert(Cm)
            q = cv2.gemm(e, Ci, 1.0, None, 0.0, flags=cv2.GEMM_1_T)
            lik = cv2.gemm(q, e, 1.0, None, 0.0)
            dist[icy, icx] = np.sqrt(lik[0, 0])
    n = cv2.normalize(dist, None, 0.0, 255.0, cv2.NORM_MINMAX)
    d8 = np.clip(np.rint(n), 0, 255).astype(np.uint8)  # convertTo(CV_8UC1): cvRound + saturate
    mask = ((d8 < 20).astype(np.uint8) * 255) // 255
    return dist, d8, mask


def make_geomask_small():
    w, h = 160, 120
    s = synth.SyntheticStream(3, w, h, roll_deg_per_frame=0.04)
    f0, f5 = s.frame(2), s.frame(7)
    # put a > 3.5 m value on the border too (never clamped there, SURVEY B-8)
    f0.depth_m[0, 10:30] = f32(3.9)
    f0.depth_m[40:60, 0] = f32(3.7)
    K = synth.intrinsics(w, h)
    R, T = s.pair_pose(2, 7)
    g0 = cv2.cvtColor(f0.bgr, cv2.COLOR_BGR2GRAY)
    g5 = cv2.cvtColor(f5.bgr, cv2.COLOR_BGR2GRAY)
    flow = cv2.calcOpticalFlowFarneback(g0, g5, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    e0 = literal_get_edge(f0.depth_m, K)
    e5 = literal_get_edge(f5.depth_m, K)
    dist, d8, mask = literal_mahalanobis(flow, f0.depth_m, f5.depth_m, e0, e5, K, R, T)
    np.savez_compressed(os.path.join(HERE, "geomask_small.npz"), K=K, R=R, T=T, bgr_ref=f0.bgr, bgr_cur=f5.bgr,
                        gray_ref=g0, gray_cur=g5, depth_ref=f0.depth_m, depth_cur=f5.depth_m, flow=flow,
                        edge_ref=e0, edge_cur=e5, dist=dist, d8=d8, mask=mask)
    print("geomask_small: edges", (e0 == 255).mean(), "written", (dist > 0).mean(), "dynamic", (mask == 0).mean(),
          "max", dist.max())


def make_geomask_640():
    """The per-pixel loop (GeoMaskMaker.cc:208-272) + normalise/threshold at the BASELINE size 640x480: the literal cv2
    transcription over every source pixel; the fixture keeps every 4th row of dist, the packed mask and min/max.  Inputs are
    reproducible from the seeded generator: flow = the oracle's Farneback of frames (0, 5) of stream 0 (pinned against
    cv2 by farneback.npz), edges = the oracle's GetEdge (pinned bit-exact against literal_get_edge by geomask_small.npz)."""
    from oracle import pyoracle as po

    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    K = synth.intrinsics()
    R, T = s.pair_pose(0, 5)
    flow = po.farneback(po.gray(f0.bgr), po.gray(f5.bgr))
    e0, e5 = po.depth_edge(f0.depth_m, K), po.depth_edge(f5.depth_m, K)
    dist, d8, mask = literal_mahalanobis(flow, f0.depth_m, f5.depth_m, e0, e5, K, R, T)
    np.savez_compressed(os.path.join(HERE, "geomask_640.npz"), K=K, R=R, T=T,
                        crc=np.array([synth.frame_crc(f0), synth.frame_crc(f5)], np.uint64),
                        flow_crc=np.uint64(__import__("zlib").crc32(flow.tobytes())), dist_rows4=dist[::4].copy(),
                        mask=np.packbits(mask), minmax=np.array([dist.min(), dist.max()], np.float32))
    print("geomask_640: written", (dist > 0).mean(), "dynamic", (mask == 0).mean(), "max", dist.max())


def make_farneback():
    out = {}
    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    g0 = cv2.cvtColor(f0.bgr, cv2.COLOR_BGR2GRAY)
    g5 = cv2.cvtColor(f5.bgr, cv2.COLOR_BGR2GRAY)
    flow = cv2.calcOpticalFlowFarneback(g0, g5, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    out["crc_640"] = np.array([synth.frame_crc(f0), synth.frame_crc(f5)], np.uint64)
    out["flow_640_rows8"] = flow[::8].copy()  # every 8th row (fixture size)
    out["flow_640_absmax"] = np.abs(flow).max()
    # quarter-size, full field, roll variant
    s2 = synth.SyntheticStream(1, 320, 240, roll_deg_per_frame=0.04)
    a, b = s2.frame(3), s2.frame(8)
    ga = cv2.cvtColor(a.bgr, cv2.COLOR_BGR2GRAY)
    gb = cv2.cvtColor(b.bgr, cv2.COLOR_BGR2GRAY)
    out["crc_320"] = np.array([synth.frame_crc(a), synth.frame_crc(b)], np.uint64)
    out["gray_320_a"], out["gray_320_b"] = ga, gb
    out["flow_320"] = cv2.calcOpticalFlowFarneback(ga, gb, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    # ragged size where the pyramid is cut short (a level would be < 32 px): 150x100 -> levels 1
    rs = np.random.RandomState(5)
    base = cv2.GaussianBlur(rs.randint(0, 256, (140, 200)).astype(np.uint8), (0, 0), 2.0)
    base = cv2.normalize(base, None, 0, 255, cv2.NORM_MINMAX)
    pa = np.ascontiguousarray(base[10:110, 10:160])
    pb = np.ascontiguousarray(base[12:112, 13:163])
    out["gray_150_a"], out["gray_150_b"] = pa, pb
    out["flow_150"] = cv2.calcOpticalFlowFarneback(pa, pb, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    np.savez_compressed(os.path.join(HERE, "farneback.npz"), **out)
    print("farneback: absmax", out["flow_640_absmax"], "150x100 median", np.median(out["flow_150"][..., 0]))


def make_erode():
    """Frame ctor filter (src/Frame.cc:258-282): cv2.erode with the 31x31 ellipse on a real mask of the pipeline, and the
    keep flags of the reference's loop on the reference ORB keypoints of the newest frame."""
    from oracle import pyoracle as po

    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    K = synth.intrinsics()
    R, T = s.pair_pose(0, 5)
    g0 = cv2.cvtColor(f0.bgr, cv2.COLOR_BGR2GRAY)
    g5 = cv2.cvtColor(f5.bgr, cv2.COLOR_BGR2GRAY)
    flow = cv2.calcOpticalFlowFarneback(g0, g5, None, 0.5, 3, 15, 3, 5, 1.2, 0)
    e0, e5 = literal_get_edge_fast(f0.depth_m, K, po), literal_get_edge_fast(f5.depth_m, K, po)
    dist, _, _ = po.mahalanobis(flow, f0.depth_m, f5.depth_m, e0, e5, K, R, T)
    n = cv2.normalize(dist, None, 0.0, 255.0, cv2.NORM_MINMAX)
    mask = (np.clip(np.rint(n), 0, 255).astype(np.uint8) < 20).astype(np.uint8)
    kernel = cv2.getStructuringElement(cv2.MORPH_ELLIPSE, (31, 31), (15, 15))
    er = cv2.erode(mask, kernel)
    kp, _, _ = po.orbref_extract(cv2.cvtColor(f5.bgr, cv2.COLOR_RGB2GRAY)) if po.have_ref() else po.orb_extract(
        cv2.cvtColor(f5.bgr, cv2.COLOR_RGB2GRAY))
    keep = np.array([er[int(k["y"]), int(k["x"])] == 1 for k in kp], np.uint8)
    np.savez_compressed(os.path.join(HERE, "erode.npz"), mask=np.packbits(mask), eroded=np.packbits(er), kp=kp, keep=keep,
                        se_rows=np.array([(int(np.nonzero(r)[0].min()), int(np.nonzero(r)[0].max()) + 1) for r in kernel], np.int32))
    print("erode: static", mask.mean(), "eroded", er.mean(), "kept", int(keep.sum()), "of", len(kp))


def make_undistort():
    """undistorted-pixel LUT of the GeoMaskMaker ctor (GeoMaskMaker.cc:56-69) for TUM1.yaml's distortion."""
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    D = np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    w, h = 640, 480
    pts = np.stack(np.meshgrid(np.arange(w, dtype=np.float32), np.arange(h, dtype=np.float32)), -1).reshape(-1, 1, 2)
    u = cv2.undistortPoints(pts, K, D, None, K).reshape(h, w, 2)
    np.savez_compressed(os.path.join(HERE, "undistort_tum1.npz"), K=K, D=D, lut_rows16=u[::16].copy(),
                        lut_sum=np.float64(u.astype(np.float64).sum()))


def make_stereo_grid():
    """Row (f)-3: Frame::UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD / AssignFeaturesToGrid + PosInGrid
    (src/Frame.cc:576-636, 815-837, 402-417, 553-565) transcribed literally (numpy f32 + cv2.undistortPoints) on the
    reference-extracted keypoints of a synthetic frame, for TUM3 (no distortion) and TUM1 (distorted)."""
    from oracle import pyoracle as po

    s = synth.SyntheticStream(0)
    fr = s.frame(5)
    g = cv2.cvtColor(fr.bgr, cv2.COLOR_RGB2GRAY)
    kp, _, _ = po.orbref_extract(g) if po.have_ref() else po.orb_extract(g)
    out = {"kp": kp, "crc": np.uint64(synth.frame_crc(fr))}
    bf = f32(40.0)
    cams = {"tum3": (synth.intrinsics(), np.zeros(5, f32)),
            "tum1": (np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], f32),
                     np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], f32))}
    h, w = fr.depth_m.shape
    for name, (K, D) in cams.items():
        N = len(kp)
        # UndistortKeyPoints (:576-606)
        if D[0] == 0.0:
            un = np.stack([kp["x"], kp["y"]], 1).astype(f32)
        else:
            mat = np.stack([kp["x"], kp["y"]], 1).astype(f32).reshape(-1, 1, 2)
            un = cv2.undistortPoints(mat, K, D, None, K).reshape(-1, 2).astype(f32)
        # ComputeImageBounds (:608-636)
        if D[0] != 0.0:
            c = np.array([[0, 0], [w, 0], [0, h], [w, h]], f32).reshape(-1, 1, 2)
            c = cv2.undistortPoints(c, K, D, None, K).reshape(-1, 2).astype(f32)
            mnMinX, mnMaxX = min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0])
            mnMinY, mnMaxY = min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])
        else:
            mnMinX, mnMaxX, mnMinY, mnMaxY = f32(0), f32(w), f32(0), f32(h)
        winv = f32(f32(64) / f32(mnMaxX - mnMinX))
        hinv = f32(f32(48) / f32(mnMaxY - mnMinY))
        # ComputeStereoFromRGBD (:815-837): depth at the DISTORTED keypoint, uRight from the undistorted x
        depth = np.full(N, -1, f32)
        uright = np.full(N, -1, f32)
        for i in range(N):
            d = fr.depth_m[int(kp["y"][i]), int(kp["x"][i])]
            if d > 0:
                depth[i] = d
                uright[i] = f32(un[i, 0] - f32(bf / d))
        # AssignFeaturesToGrid + PosInGrid (:402-417, 553-565): C round() = half away from zero
        grid = [[[] for _ in range(48)] for _ in range(64)]
        for i in range(N):
            fx_, fy_ = float(f32(f32(un[i, 0] - mnMinX) * winv)), float(f32(f32(un[i, 1] - mnMinY) * hinv))
            px = int(np.floor(abs(fx_) + 0.5) * np.sign(fx_))
            py = int(np.floor(abs(fy_) + 0.5) * np.sign(fy_))
            if px < 0 or px >= 64 or py < 0 or py >= 48:
                continue
            grid[px][py].append(i)
        start, items = [], []
        for px in range(64):
            for py in range(48):
                start.append(len(items))
                items += grid[px][py]
        start.append(len(items))
        out[name + "_K"], out[name + "_D"] = K, D
        out[name + "_un"] = un
        out[name + "_bounds"] = np.array([mnMinX, mnMaxX, mnMinY, mnMaxY], f32)
        out[name + "_depth"], out[name + "_uright"] = depth, uright
        out[name + "_cell_start"], out[name + "_cell_items"] = np.array(start, np.int32), np.array(items, np.int32)
        print("stereo_grid", name, "keypoints", N, "in grid", len(items), "with depth", int((depth > 0).sum()),
              "bounds", mnMinX, mnMaxX, mnMinY, mnMaxY)
    np.savez_compressed(os.path.join(HERE, "stereo_grid.npz"), **out)


def literal_get_edge_fast(depth, K, po):
    return po.depth_edge(depth, K)  # pinned bit-exact against literal_get_edge by geomask_small.npz


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "erode":
        make_erode()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "stereo_grid":
        make_stereo_grid()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "geomask_640":
        make_geomask_640()
        sys.exit(0)
    make_geomask_small()
    make_geomask_640()
    make_farneback()
    make_erode()
    make_undistort()
    make_stereo_grid()
    print("cv2", cv2.__version__)
