"""ctypes loader for the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE ONLY — importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (gd-slam_b200) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build() -> None:
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def build_native() -> str | None:
    """-O3 -march=native build for bench.py's CPU timing legs, compiled on the machine it runs on; None if that fails."""
    try:
        subprocess.run(["make", "-s", "-C", _HERE, "native"], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    except Exception:
        return None
    p = os.path.join(_HERE, "_native", "liboracle_native.so")
    return p if os.path.exists(p) else None


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        # GD_ORACLE_LIB: bench.py's CPU legs point this at the native timing build; the parity tests never set it
        path = os.environ.get("GD_ORACLE_LIB") or os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        L = _LIB
        L.gdo_gray_u8.argtypes = [u8p, C.c_size_t, C.c_int, C.c_int, C.c_int, u8p]
        L.gdo_inv3_f32.argtypes = [f32p, f32p]
        L.gdo_inv3_f64.argtypes = [f64p, f64p]
        L.gdo_depth_edge.argtypes = [f32p, C.c_int, C.c_int, f32p, u8p]
        L.gdo_mahalanobis.argtypes = [f32p, f32p, f32p, u8p, u8p, C.c_void_p, C.c_int, C.c_int, f32p, f32p, f32p,
                                      f32p, C.c_void_p, C.c_void_p]
        L.gdo_normalize_threshold.argtypes = [f32p, C.c_int, C.c_int, u8p, C.c_void_p, C.c_void_p]
        L.gdo_depth2std.argtypes = [C.c_float, C.c_float]
        L.gdo_depth2std.restype = C.c_float
        L.gdo_farneback.argtypes = [u8p, u8p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int,
                                    C.c_double, f32p]
        L.gdo_farneback_polyexp_level.argtypes = [u8p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_double,
                                                  f32p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.gdo_geomask_pair.argtypes = [u8p, u8p, f32p, f32p, C.c_int, C.c_int, f32p, f32p, f32p, u8p, C.c_void_p,
                                       C.c_void_p]
    return _LIB


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


def gray(bgr: np.ndarray, order: int = 0) -> np.ndarray:
    bgr = _c(bgr, np.uint8)
    h, w = bgr.shape[:2]
    out = np.empty((h, w), np.uint8)
    lib().gdo_gray_u8(bgr.reshape(-1), w * 3, w, h, order, out.reshape(-1))
    return out


def inv3(m: np.ndarray) -> np.ndarray:
    if m.dtype == np.float64:
        out = np.empty(9, np.float64)
        lib().gdo_inv3_f64(_c(m, np.float64).reshape(-1), out)
    else:
        out = np.empty(9, np.float32)
        lib().gdo_inv3_f32(_c(m, np.float32).reshape(-1), out)
    return out.reshape(3, 3)


def depth_edge(depth: np.ndarray, K: np.ndarray) -> np.ndarray:
    depth = _c(depth, np.float32)
    h, w = depth.shape
    out = np.empty((h, w), np.uint8)
    lib().gdo_depth_edge(depth.reshape(-1), w, h, _c(K, np.float32).reshape(-1), out.reshape(-1))
    return out


def mahalanobis(flow, depth_ref, depth_cur, edge_ref, edge_cur, K, R, T, lut=None):
    flow = _c(flow, np.float32)
    h, w = flow.shape[:2]
    dist = np.empty((h, w), np.float32)
    written = np.empty((h, w), np.uint8)
    src = np.empty((h, w), np.int32)
    lut_p = None
    if lut is not None:
        lut = _c(lut, np.float32)
        lut_p = lut.ctypes.data_as(C.c_void_p)
    lib().gdo_mahalanobis(flow.reshape(-1), _c(depth_ref, np.float32).reshape(-1),
                          _c(depth_cur, np.float32).reshape(-1), _c(edge_ref, np.uint8).reshape(-1),
                          _c(edge_cur, np.uint8).reshape(-1), lut_p, w, h, _c(K, np.float32).reshape(-1),
                          _c(R, np.float32).reshape(-1), _c(T, np.float32).reshape(-1), dist.reshape(-1),
                          written.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p))
    return dist, written, src


def normalize_threshold(dist):
    dist = _c(dist, np.float32)
    h, w = dist.shape
    mask = np.empty((h, w), np.uint8)
    d8 = np.empty((h, w), np.uint8)
    mm = np.empty(2, np.float32)
    lib().gdo_normalize_threshold(dist.reshape(-1), w, h, mask.reshape(-1), d8.ctypes.data_as(C.c_void_p),
                                  mm.ctypes.data_as(C.c_void_p))
    return mask, d8, mm


def farneback(prev, nxt, pyr_scale=0.5, levels=3, winsize=15, iterations=3, poly_n=5, poly_sigma=1.2):
    prev = _c(prev, np.uint8)
    nxt = _c(nxt, np.uint8)
    h, w = prev.shape
    flow = np.empty((h, w, 2), np.float32)
    lib().gdo_farneback(prev.reshape(-1), nxt.reshape(-1), w, h, pyr_scale, levels, winsize, iterations, poly_n,
                        poly_sigma, flow.reshape(-1))
    return flow


def polyexp_level(img, k, pyr_scale=0.5, poly_n=5, poly_sigma=1.2):
    img = _c(img, np.uint8)
    h, w = img.shape
    out = np.empty(h * w * 5, np.float32)
    lw, lh = C.c_int(0), C.c_int(0)
    lib().gdo_farneback_polyexp_level(img.reshape(-1), w, h, pyr_scale, k, poly_n, poly_sigma, out, C.byref(lw),
                                      C.byref(lh))
    return out[: lw.value * lh.value * 5].reshape(lh.value, lw.value, 5).copy()


def geomask_pair(bgr_ref, bgr_cur, depth_ref, depth_cur, K, R, T, want_debug=False):
    bgr_ref = _c(bgr_ref, np.uint8)
    h, w = bgr_ref.shape[:2]
    mask = np.empty((h, w), np.uint8)
    flow = np.empty((h, w, 2), np.float32) if want_debug else None
    dist = np.empty((h, w), np.float32) if want_debug else None
    lib().gdo_geomask_pair(bgr_ref.reshape(-1), _c(bgr_cur, np.uint8).reshape(-1),
                           _c(depth_ref, np.float32).reshape(-1), _c(depth_cur, np.float32).reshape(-1), w, h,
                           _c(K, np.float32).reshape(-1), _c(R, np.float32).reshape(-1),
                           _c(T, np.float32).reshape(-1), mask.reshape(-1),
                           flow.ctypes.data_as(C.c_void_p) if want_debug else None,
                           dist.ctypes.data_as(C.c_void_p) if want_debug else None)
    return (mask, flow, dist) if want_debug else mask


# ---------------------------------------------------------------------------------------------- ORB
KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
_ORB_BOUND = False
_REF = None


def _orb():
    global _ORB_BOUND
    L = lib()
    if not _ORB_BOUND:
        ip = C.POINTER(C.c_int)
        L.gdo_orb_config.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p]
        L.gdo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.gdo_fast_atan2.restype = C.c_float
        L.gdo_resize_u8.argtypes = [u8p, C.c_int, C.c_int, u8p, C.c_int, C.c_int]
        L.gdo_gaussian7_u8.argtypes = [u8p, C.c_int, C.c_int, u8p]
        L.gdo_fast_detect.argtypes = [u8p, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
        L.gdo_orb_candidates.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_int, f32p, C.c_int]
        L.gdo_orb_distribute.argtypes = [f32p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i32p, C.c_int]
        L.gdo_ic_angle.argtypes = [u8p, C.c_int, C.c_int, C.c_int]
        L.gdo_ic_angle.restype = C.c_float
        L.gdo_orb_descriptor.argtypes = [u8p, C.c_int, C.c_int, C.c_int, C.c_float, u8p]
        L.gdo_orb_descriptor.restype = None
        L.gdo_retain_best_order.argtypes = [f32p, C.c_int, C.c_int, i32p]
        L.gdo_retain_best_order.restype = C.c_int
        L.gdo_sort_matches_order.argtypes = [f32p, C.c_int, i32p]
        L.gdo_sort_matches_order.restype = None
        L.gdo_orb_extract.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        _ORB_BOUND = True
    return L


def orb_config(w=640, h=480, nfeatures=1500, scale=1.2, nlevels=8):
    n = np.zeros(nlevels, np.int32)
    sz = np.zeros((nlevels, 2), np.int32)
    sc = np.zeros(nlevels, np.float32)
    um = np.zeros(16, np.int32)
    _orb().gdo_orb_config(nfeatures, scale, nlevels, w, h, n.ctypes.data, sz.ctypes.data, sc.ctypes.data, um.ctypes.data)
    return dict(n_per_level=n, level_sizes=sz, scales=sc, umax=um)


def fast_atan2(y, x):
    return _orb().gdo_fast_atan2(float(y), float(x))


def resize_u8(src, dw, dh):
    src = _c(src, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    _orb().gdo_resize_u8(src.reshape(-1), src.shape[1], src.shape[0], out.reshape(-1), dw, dh)
    return out


def gaussian7(src):
    src = _c(src, np.uint8)
    out = np.empty_like(src)
    _orb().gdo_gaussian7_u8(src.reshape(-1), src.shape[1], src.shape[0], out.reshape(-1))
    return out


def fast_detect(img, threshold):
    img = _c(img, np.uint8)
    cap = img.size
    out = np.empty((cap, 3), np.int32)
    n = _orb().gdo_fast_detect(img.reshape(-1), img.shape[1], img.shape[0], threshold, out.reshape(-1), cap)
    return out[:n].copy()


def orb_candidates(img, ini_th=20, min_th=7):
    img = _c(img, np.uint8)
    cap = img.size // 4
    out = np.empty((cap, 3), np.float32)
    n = _orb().gdo_orb_candidates(img.reshape(-1), img.shape[1], img.shape[0], ini_th, min_th, out.reshape(-1), cap)
    return out[:n].copy()


def orb_distribute(cand, minX, maxX, minY, maxY, N):
    cand = _c(cand, np.float32)
    kept = np.empty(N + 64, np.int32)
    n = _orb().gdo_orb_distribute(cand.reshape(-1), cand.shape[0], minX, maxX, minY, maxY, N, kept, kept.size)
    return kept[:n].copy()


def ic_angle(img, x, y):
    img = _c(img, np.uint8)
    return _orb().gdo_ic_angle(img.reshape(-1), img.shape[1], int(x), int(y))


def retain_best_order(resp, n_points):
    """cv::KeyPointsFilter::retainBest: indices kept, in the order std::nth_element + std::partition leave them."""
    resp = _c(resp, np.float32)
    perm = np.zeros(len(resp), np.int32)
    n = _orb().gdo_retain_best_order(resp, len(resp), int(n_points), perm)
    return perm[:n].copy()


def sort_matches_order(dist):
    """Order std::sort leaves a vector<DMatch> in (comparison on the distance only)."""
    dist = _c(dist, np.float32)
    perm = np.zeros(len(dist), np.int32)
    _orb().gdo_sort_matches_order(dist, len(dist), perm)
    return perm


def orb_descriptor(blurred, x, y, angle_deg):
    """rBRIEF descriptor of one keypoint at integer (x, y) of an already blurred level image (32 bytes)."""
    blurred = _c(blurred, np.uint8)
    out = np.zeros(32, np.uint8)
    _orb().gdo_orb_descriptor(blurred.reshape(-1), blurred.shape[1], int(x), int(y), float(angle_deg), out)
    return out


def _extract(fn, gray, nfeatures, scale, nlevels, ini_th, min_th, want_pyramid, has_nlevel):
    gray = _c(gray, np.uint8)
    h, w = gray.shape
    cap = nfeatures + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    cfg = orb_config(w, h, nfeatures, scale, nlevels)
    # the reference has no guards here: a level narrower than one 30-px cell divides by zero at ORBextractor.cc:783-786 and a
    # portrait level with round(width / height) == 0 indexes an empty node vector at :543-566.  Both sides reject such input.
    for lw, lh in cfg["level_sizes"]:
        bw, bh = float(int(lw) - 32 + 3), float(int(lh) - 32 + 3)
        if int(bw / 30.0) < 1 or int(bh / 30.0) < 1:
            raise ValueError(f"level {lw}x{lh} too small for one FAST cell")
        if int(np.round(np.float32(bw) / np.float32(bh))) < 1:
            raise ValueError(f"level {lw}x{lh}: aspect ratio not supported by DistributeOctTree (nIni = 0)")
    tot = int(sum(int(a) * int(b) for a, b in cfg["level_sizes"]))
    pyr = np.zeros(tot, np.uint8) if want_pyramid else None
    aux = np.zeros(nlevels * 2, np.int32)
    n = fn(gray.reshape(-1), w, h, w, nfeatures, scale, nlevels, ini_th, min_th, kps.ctypes.data, desc.ctypes.data, cap,
           pyr.ctypes.data if want_pyramid else None, aux.ctypes.data)
    assert n <= cap
    levels = None
    if want_pyramid:
        levels, off = [], 0
        for lw, lh in cfg["level_sizes"]:
            levels.append(pyr[off: off + lw * lh].reshape(lh, lw))
            off += lw * lh
    return kps[:n].copy(), desc[:n].copy(), levels


def orb_extract(gray, nfeatures=1500, scale=1.2, nlevels=8, ini_th=20, min_th=7, want_pyramid=False):
    """Restated ORBextractor::operator() (oracle/orb_oracle.cpp)."""
    return _extract(_orb().gdo_orb_extract, gray, nfeatures, scale, nlevels, ini_th, min_th, want_pyramid, True)


def have_ref() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "liborbref.so"))


def orbref_extract(gray, nfeatures=1500, scale=1.2, nlevels=8, ini_th=20, min_th=7, want_pyramid=False):
    """The reference's own src/ORBextractor.cc (compiled verbatim, oracle/_ref/liborbref.so)."""
    global _REF
    if _REF is None:
        _REF = C.CDLL(os.path.join(_HERE, "_ref", "liborbref.so"))
        _REF.orbref_extract.argtypes = [u8p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    return _extract(_REF.orbref_extract, gray, nfeatures, scale, nlevels, ini_th, min_th, want_pyramid, False)


# ---------------------------------------------------------------------------------------------- Frame ctor filter (8f-2)
def erode31(mask):
    mask = _c(mask, np.uint8)
    out = np.empty_like(mask)
    L = lib()
    L.gdo_erode31.argtypes = [u8p, C.c_int, C.c_int, u8p]
    L.gdo_erode31(mask.reshape(-1), mask.shape[1], mask.shape[0], out.reshape(-1))
    return out


def erode_filter(mask, kps):
    """keep flags of Frame.cc:267-277 for keypoints given as KP_DTYPE records."""
    mask = _c(mask, np.uint8)
    kps = np.ascontiguousarray(kps)
    keep = np.zeros(len(kps), np.uint8)
    L = lib()
    L.gdo_erode_filter.argtypes = [u8p, C.c_int, C.c_int, C.c_void_p, C.c_int, u8p]
    L.gdo_erode_filter.restype = C.c_int
    n = L.gdo_erode_filter(mask.reshape(-1), mask.shape[1], mask.shape[0], kps.ctypes.data, len(kps), keep)
    assert n == int(keep.sum())
    return keep


def stereo_grid(depth_m, kps, bf, K=None, D=None, want_undistorted=False):
    """Frame::UndistortKeyPoints / ComputeImageBounds / ComputeStereoFromRGBD / AssignFeaturesToGrid:
    (depth, uright, cell_start[3073], cell_items) [+ (mvKeysUn xy, bounds)].  D = None / zeros: undistorted camera."""
    depth_m = _c(depth_m, np.float32)
    kps = np.ascontiguousarray(kps)
    n = len(kps)
    d = np.empty(n, np.float32)
    ur = np.empty(n, np.float32)
    cs = np.empty(64 * 48 + 1, np.int32)
    ci = np.empty(max(n, 1), np.int32)
    un = np.empty((max(n, 1), 2), np.float32)
    bounds = np.empty(4, np.float32)
    Kf = _c(np.eye(3) if K is None else K, np.float32).reshape(-1)
    Df = np.zeros(5, np.float32)
    if D is not None:
        Df[: len(np.ravel(D))] = np.ravel(D)
    L = lib()
    L.gdo_stereo_grid.argtypes = [f32p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_float, f32p, f32p, f32p, f32p, i32p, i32p,
                                  f32p, f32p]
    L.gdo_stereo_grid(depth_m.reshape(-1), depth_m.shape[1], depth_m.shape[0], kps.ctypes.data, n, bf, Kf, Df, d, ur, cs, ci,
                      un.reshape(-1), bounds)
    if want_undistorted:
        return d, ur, cs, ci[: cs[-1]].copy(), un[:n].copy(), bounds
    return d, ur, cs, ci[: cs[-1]].copy()
