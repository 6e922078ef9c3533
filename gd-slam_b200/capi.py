"""ctypes binding of include/gdslam_cuda.h (libgdslam_cuda.so).

Used by the parity tests and bench.py to drive the C ABI exactly the way the C++ shim does.  There is no
fallback of any kind: if the shared library is missing or a call fails, GdError is raised.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgdslam_cuda.so")

GD_OK, GD_EINVAL, GD_ENODEVICE, GD_ECUDA, GD_ENOMEM, GD_ECAPACITY, GD_EINTERNAL = 0, -1, -2, -3, -4, -5, -6
DBG_FLOW, DBG_DIST, DBG_EDGE_REF, DBG_EDGE_CUR, DBG_GRAY_CUR, DBG_MINMAX, DBG_LUT = range(7)


class GdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libgdslam_cuda error {code}: {msg}")
        self.code = code


class Keypoint(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float), ("size", C.c_float), ("angle", C.c_float),
                ("response", C.c_float), ("octave", C.c_int32), ("class_id", C.c_int32)]


KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])


class FrontendConfig(C.Structure):
    _fields_ = [("K", C.c_float * 9), ("dist", C.c_float * 5), ("ndist", C.c_int), ("depth_factor", C.c_float),
                ("width", C.c_int), ("height", C.c_int), ("device", C.c_int), ("batch", C.c_int),
                ("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("ini_th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("orb_gray_order", C.c_int), ("kp_capacity", C.c_int),
                ("staged_slots", C.c_int), ("getrt", C.c_int)]


# int hook(void* user, int stream, const float* obj, const float* pix, int n, float R[9], float T[3])
POSE_HOOK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float),
                        C.POINTER(C.c_float))


def build_library() -> None:
    subprocess.run(["make", "-s", "-j8", "-C", os.path.join(_HERE, "csrc")], check=True)


_lib = None
vp = C.c_void_p
fp = C.POINTER(C.c_float)
ip = C.POINTER(C.c_int)

# every symbol include/gdslam_cuda.h declares: name -> (restype, argtypes)
SYMBOLS = {
    "gd_last_error": (C.c_char_p, []),
    "gd_abi_version": (C.c_int, []),
    "gd_device_count": (C.c_int, [ip]),
    "gd_device_info": (C.c_int, [C.c_int, C.c_char_p, C.c_int, ip, C.POINTER(C.c_size_t)]),
    "gd_host_alloc": (C.c_int, [C.POINTER(vp), C.c_size_t]),
    "gd_host_free": (C.c_int, [vp]),
    "gd_probe_copy": (C.c_int, [C.c_int, C.c_size_t, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gd_geomask_create": (C.c_int, [C.POINTER(vp), fp, fp, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int]),
    "gd_geomask_destroy": (None, [vp]),
    "gd_geomask_push": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.c_size_t]),
    "gd_geomask_mask": (C.c_int, [vp, fp, fp, ip, C.POINTER(vp), C.c_size_t]),
    "gd_geomask_frames": (C.c_int, [vp]),
    "gd_geomask_enable_getrt": (C.c_int, [vp]),
    "gd_geomask_getrt_points": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_frontend_set_pose_hook": (C.c_int, [vp, POSE_HOOK, vp]),
    "gd_frontend_fetch_getrt": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_geomask_debug_fetch": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
    "gd_orb_create": (C.c_int, [C.POINTER(vp), C.c_int, C.c_float, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int]),
    "gd_orb_destroy": (None, [vp]),
    "gd_orb_extract": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.c_int, C.c_int, C.POINTER(vp), C.POINTER(vp),
                                 C.c_int, ip]),
    "gd_orb_fetch_level": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t, ip, ip]),
    "gd_orb_level_size": (C.c_int, [vp, C.c_int, ip, ip]),
    "gd_orb_features_per_level": (C.c_int, [vp, ip]),
    "gd_frontend_create": (C.c_int, [C.POINTER(vp), C.POINTER(FrontendConfig)]),
    "gd_frontend_destroy": (None, [vp]),
    "gd_frontend_step": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.c_size_t, fp, fp, ip,
                                   C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_frontend_step_u16": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.c_size_t, fp, fp, ip,
                                       C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_frontend_stage": (C.c_int, [vp, C.c_int, C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.c_size_t]),
    "gd_frontend_step_staged": (C.c_int, [vp, C.c_int, fp, fp, ip]),
    "gd_frontend_fetch": (C.c_int, [vp, C.POINTER(vp), C.c_size_t, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_frontend_fetch_filtered": (C.c_int, [vp, C.POINTER(vp), C.POINTER(vp), ip]),
    "gd_stage_erode_filter": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "gd_frontend_fetch_stereo_grid": (C.c_int, [vp, C.c_float, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)]),
    "gd_frontend_fetch_stereo_grid_un": (C.c_int, [vp, C.c_float, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                                   C.POINTER(vp)]),
    "gd_frontend_sync": (C.c_int, [vp]),
    "gd_frontend_timer_begin": (C.c_int, [vp]),
    "gd_frontend_timer_end": (C.c_int, [vp, fp]),
    "gd_frontend_launch_count": (C.c_int, [vp, C.POINTER(C.c_longlong)]),
    "gd_frontend_profile": (C.c_int, [vp, C.c_int]),
    "gd_frontend_profile_read": (C.c_int, [vp, C.c_int, C.POINTER(C.c_char_p), fp, C.POINTER(C.c_longlong), ip]),
    "gd_frontend_debug_fetch": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_size_t]),
    "gd_frontend_flush_l2": (C.c_int, [vp]),
    "gd_stage_gray": (C.c_int, [C.c_int, vp, C.c_size_t, C.c_int, C.c_int, C.c_int, vp]),
    "gd_stage_depth_edge": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, fp, vp]),
    "gd_stage_mahalanobis": (C.c_int, [C.c_int, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, fp, fp, fp, vp, vp, vp]),
    "gd_stage_farneback": (C.c_int, [C.c_int, vp, vp, C.c_int, C.c_int, vp]),
    "gd_stage_polyexp": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, ip, ip]),
    "gd_stage_orb_pyramid": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_float, vp, ip]),
    "gd_stage_fast_cells": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_int, ip]),
    "gd_stage_gaussian7": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp]),
    "gd_getrt_points": (C.c_int, [C.c_int, vp, vp, C.c_int, C.c_int, vp, fp, fp, C.c_int, vp, vp, ip]),
    "gd_stage_cvorb_detect_and_compute": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, ip]),
    "gd_stage_fast_whole": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, C.c_int, vp]),
    "gd_stage_resize_linear_exact": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp, C.c_int, C.c_int]),
    "gd_stage_gaussian7_float": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp]),
    "gd_stage_harris": (C.c_int, [C.c_int, vp, C.c_int, C.c_int, vp, vp, C.c_int, vp]),
    "gd_stage_hamming_crosscheck": (C.c_int, [C.c_int, vp, C.c_int, vp, C.c_int, vp, vp, vp, C.c_int, ip]),
}


def lib() -> C.CDLL:
    """Load libgdslam_cuda.so; raise (never fall back) when it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GdError(GD_ENODEVICE, f"{LIB_PATH} not built (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                                        "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name, None)
            if fn is None:  # reported by missing_symbols(); tests/test_capi_symbols.py fails on any
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def missing_symbols():
    L = lib()
    return [n for n in SYMBOLS if not hasattr(L, n)]


def check(code: int) -> None:
    if code != GD_OK:
        raise GdError(code, lib().gd_last_error().decode(errors="replace"))


def _fptr(a):
    return None if a is None else a.ctypes.data_as(fp)


def _vptr(a):
    return None if a is None else a.ctypes.data_as(vp)


def _ptr_array(arrs):
    return (vp * len(arrs))(*[a.ctypes.data for a in arrs])


def device_count() -> int:
    n = C.c_int(0)
    code = lib().gd_device_count(C.byref(n))
    return n.value if code == GD_OK else 0


def device_info(device=0):
    name = C.create_string_buffer(256)
    sm = C.c_int(0)
    mem = C.c_size_t(0)
    check(lib().gd_device_info(device, name, 256, C.byref(sm), C.byref(mem)))
    return name.value.decode(), sm.value, mem.value


def probe_copy(device=0, nbytes=256 << 20, iters=8, to_device=True) -> float:
    """GB/s of bare pinned-memory copies on one device (gd_probe_copy)."""
    g = C.c_double(0)
    check(lib().gd_probe_copy(device, nbytes, iters, 1 if to_device else 0, C.byref(g)))
    return g.value


def pinned_empty(shape, dtype):
    """numpy array backed by page-locked memory from gd_host_alloc (kept alive by the returned array)."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    p = vp()
    check(lib().gd_host_alloc(C.byref(p), nbytes))
    buf = (C.c_char * nbytes).from_address(p.value)
    arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
    _PINNED[id(buf)] = (buf, p)
    return arr


_PINNED = {}


# ---------------------------------------------------------------------------------------------- stages
def stage_gray(bgr, order=0, device=0):
    bgr = np.ascontiguousarray(bgr, np.uint8)
    h, w = bgr.shape[:2]
    out = np.empty((h, w), np.uint8)
    check(lib().gd_stage_gray(device, _vptr(bgr), w * 3, w, h, order, _vptr(out)))
    return out


def stage_depth_edge(depth, K, device=0):
    depth = np.ascontiguousarray(depth, np.float32)
    K = np.ascontiguousarray(K, np.float32)
    h, w = depth.shape
    out = np.empty((h, w), np.uint8)
    check(lib().gd_stage_depth_edge(device, _vptr(depth), w, h, _fptr(K), _vptr(out)))
    return out


def stage_mahalanobis(flow, d_ref, d_cur, e_ref, e_cur, K, R, T, lut=None, device=0):
    flow = np.ascontiguousarray(flow, np.float32)
    h, w = flow.shape[:2]
    d_ref = np.ascontiguousarray(d_ref, np.float32)
    d_cur = np.ascontiguousarray(d_cur, np.float32)
    e_ref = np.ascontiguousarray(e_ref, np.uint8)
    e_cur = np.ascontiguousarray(e_cur, np.uint8)
    K = np.ascontiguousarray(K, np.float32)
    R = np.ascontiguousarray(R, np.float32)
    T = np.ascontiguousarray(T, np.float32)
    if lut is not None:
        lut = np.ascontiguousarray(lut, np.float32)
    dist = np.empty((h, w), np.float32)
    mask = np.empty((h, w), np.uint8)
    mm = np.empty(2, np.float32)
    check(lib().gd_stage_mahalanobis(device, _vptr(flow), _vptr(d_ref), _vptr(d_cur), _vptr(e_ref), _vptr(e_cur),
                                     _vptr(lut), w, h, _fptr(K), _fptr(R), _fptr(T), _vptr(dist), _vptr(mask),
                                     _vptr(mm)))
    return dist, mask, mm


def stage_erode_filter(mask, kps, device=0):
    mask = np.ascontiguousarray(mask, np.uint8)
    kps = np.ascontiguousarray(kps)
    keep = np.zeros(len(kps), np.uint8)
    check(lib().gd_stage_erode_filter(device, _vptr(mask), mask.shape[1], mask.shape[0], _vptr(kps), len(kps), _vptr(keep)))
    return keep


def stage_farneback(prev, nxt, device=0):
    prev = np.ascontiguousarray(prev, np.uint8)
    nxt = np.ascontiguousarray(nxt, np.uint8)
    h, w = prev.shape
    flow = np.empty((h, w, 2), np.float32)
    check(lib().gd_stage_farneback(device, _vptr(prev), _vptr(nxt), w, h, _vptr(flow)))
    return flow


def stage_polyexp(gray, k, device=0):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.empty(5 * h * w, np.float32)
    lw, lh = C.c_int(0), C.c_int(0)
    check(lib().gd_stage_polyexp(device, _vptr(gray), w, h, k, _vptr(out), C.byref(lw), C.byref(lh)))
    n = lw.value * lh.value
    planes = out[: 5 * n].reshape(5, lh.value, lw.value)
    return np.ascontiguousarray(np.moveaxis(planes, 0, 2))  # (lh, lw, 5) like OpenCV's CV_32FC5


# ---------------------------------------------------------------------------------------------- GeoMaskMaker
class GeoMask:
    """Mirror of the reference's GeoMaskMaker for `batch` lockstep streams (include/GeoMaskMaker.h:52-116)."""

    def __init__(self, K, dist=None, depth_factor=5000.0, width=640, height=480, device=0, batch=1):
        self.w, self.h, self.batch = width, height, batch
        K = np.ascontiguousarray(K, np.float32)
        d = None if dist is None else np.ascontiguousarray(dist, np.float32)
        self._h = vp()
        check(lib().gd_geomask_create(C.byref(self._h), _fptr(K), _fptr(d), 0 if d is None else d.size, depth_factor,
                                      width, height, device, batch))

    def close(self):
        if self._h:
            lib().gd_geomask_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def add_new_image(self, bgrs, depths):
        """AddNewImage (GeoMaskMaker.cc:409-429): one (bgr, depth_m) per stream."""
        bgrs = [np.ascontiguousarray(b, np.uint8) for b in bgrs]
        depths = [np.ascontiguousarray(d, np.float32) for d in depths]
        assert len(bgrs) == self.batch and len(depths) == self.batch
        check(lib().gd_geomask_push(self._h, _ptr_array(bgrs), self.w * 3, _ptr_array(depths), self.w * 4))

    def get_no_gmm_mask(self, R=None, T=None, pose_valid=None):
        """GetNoGMMmask (GeoMaskMaker.cc:167-408) with the pose as input; returns one {0,1} mask per stream."""
        B = self.batch
        R = np.ascontiguousarray(np.tile(np.eye(3, dtype=np.float32), (B, 1, 1)) if R is None else R, np.float32)
        T = np.ascontiguousarray(np.zeros((B, 3), np.float32) if T is None else T, np.float32)
        pv = np.ascontiguousarray(np.ones(B, np.int32) if pose_valid is None else pose_valid, np.int32)
        masks = [np.empty((self.h, self.w), np.uint8) for _ in range(B)]
        check(lib().gd_geomask_mask(self._h, _fptr(R), _fptr(T), pv.ctypes.data_as(ip), _ptr_array(masks), self.w))
        return masks

    def frames(self):
        return lib().gd_geomask_frames(self._h)

    def enable_getrt(self):
        check(lib().gd_geomask_enable_getrt(self._h))

    def getrt_points(self):
        """GetRt up to solvePnPRansac for the buffered pair: list of (object_points [n,3], image_pixels [n,2]) per stream."""
        B = self.batch
        obj = [np.zeros((100, 3), np.float32) for _ in range(B)]
        pix = [np.zeros((100, 2), np.float32) for _ in range(B)]
        n = np.zeros(B, np.int32)
        check(lib().gd_geomask_getrt_points(self._h, _ptr_array(obj), _ptr_array(pix), n.ctypes.data_as(ip)))
        return [(obj[b][: n[b]].copy(), pix[b][: n[b]].copy()) for b in range(B)]

    def debug(self, what, stream=0):
        shapes = {DBG_FLOW: ((self.h, self.w, 2), np.float32), DBG_DIST: ((self.h, self.w), np.float32),
                  DBG_EDGE_REF: ((self.h, self.w), np.uint8), DBG_EDGE_CUR: ((self.h, self.w), np.uint8),
                  DBG_GRAY_CUR: ((self.h, self.w), np.uint8), DBG_MINMAX: ((2,), np.float32),
                  DBG_LUT: ((self.h, self.w, 2), np.float32)}
        shp, dt = shapes[what]
        out = np.empty(shp, dt)
        check(lib().gd_geomask_debug_fetch(self._h, what, stream, _vptr(out), out.nbytes))
        return out


# ---------------------------------------------------------------------------------------------- ORB stages
def stage_orb_pyramid(gray, nlevels=8, scale=1.2, device=0):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    out = np.empty(w * h * 4, np.uint8)
    sizes = np.zeros((nlevels, 2), np.int32)
    check(lib().gd_stage_orb_pyramid(device, _vptr(gray), w, h, nlevels, scale, _vptr(out), sizes.ctypes.data_as(ip)))
    levels, off = [], 0
    for lw, lh in sizes:
        levels.append(out[off: off + lw * lh].reshape(lh, lw).copy())
        off += lw * lh
    return levels


def stage_fast_cells(gray, ini_th=20, min_th=7, device=0):
    gray = np.ascontiguousarray(gray, np.uint8)
    h, w = gray.shape
    cap = w * h // 4
    out = np.empty((cap, 3), np.float32)
    n = C.c_int(0)
    check(lib().gd_stage_fast_cells(device, _vptr(gray), w, h, ini_th, min_th, _vptr(out), cap, C.byref(n)))
    return out[: n.value].copy()


def stage_gaussian7(gray, device=0):
    gray = np.ascontiguousarray(gray, np.uint8)
    out = np.empty_like(gray)
    check(lib().gd_stage_gaussian7(device, _vptr(gray), gray.shape[1], gray.shape[0], _vptr(out)))
    return out


# ---------------------------------------------------------------------------------------------- ORBextractor
# ---- GetRt building blocks
def getrt_points(gray_first, gray_second, depth_first_m, K, dist=None, device=0):  # dist: k1 k2 p1 p2 [k3] (may be zero)
    """GeoMaskMaker::GetRt up to (not including) solvePnPRansac: (object_points [n,3], image_pixels [n,2]) f32."""
    g1, g2 = np.ascontiguousarray(gray_first, np.uint8), np.ascontiguousarray(gray_second, np.uint8)
    dep = np.ascontiguousarray(depth_first_m, np.float32)
    Kf = np.ascontiguousarray(K, np.float32).reshape(9)
    d = None if dist is None else np.ascontiguousarray(dist, np.float32).reshape(-1)
    obj, pix = np.zeros((100, 3), np.float32), np.zeros((100, 2), np.float32)
    n = C.c_int(0)
    check(lib().gd_getrt_points(device, _vptr(g1), _vptr(g2), g1.shape[1], g1.shape[0], _vptr(dep), _fptr(Kf),
                                _fptr(d) if d is not None else None, 0 if d is None else len(d), _vptr(obj), _vptr(pix), C.byref(n)))
    return obj[: n.value].copy(), pix[: n.value].copy()


def stage_cvorb_detect_and_compute(gray, nfeatures=2000, device=0):
    """cv::ORB(nfeatures, 1.2, 8, 31, 0, 2).detectAndCompute -> (keypoints[KP_DTYPE], descriptors[n, 32]) in cv2's order."""
    gray = np.ascontiguousarray(gray, np.uint8)
    cap = nfeatures + 64
    kps = np.zeros(cap, KP_DTYPE)
    desc = np.zeros((cap, 32), np.uint8)
    n = C.c_int(0)
    check(lib().gd_stage_cvorb_detect_and_compute(device, _vptr(gray), gray.shape[1], gray.shape[0], nfeatures, _vptr(kps), _vptr(desc),
                                                  cap, C.byref(n)))
    return kps[: n.value].copy(), desc[: n.value].copy()


def stage_fast_whole(gray, threshold=20, device=0):
    """cv::FAST(threshold, nonmax) on the whole image -> int array of (x, y, S') in raster order."""
    gray = np.ascontiguousarray(gray, np.uint8)
    kept = np.empty_like(gray)
    check(lib().gd_stage_fast_whole(device, _vptr(gray), gray.shape[1], gray.shape[0], threshold, _vptr(kept)))
    ys, xs = np.nonzero(kept)
    return np.stack([xs, ys, kept[ys, xs].astype(np.int64)], 1).astype(np.int32)


def stage_resize_linear_exact(src, dw, dh, device=0):
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty((dh, dw), np.uint8)
    check(lib().gd_stage_resize_linear_exact(device, _vptr(src), src.shape[1], src.shape[0], _vptr(out), dw, dh))
    return out


def stage_gaussian7_float(src, device=0):
    src = np.ascontiguousarray(src, np.uint8)
    out = np.empty_like(src)
    check(lib().gd_stage_gaussian7_float(device, _vptr(src), src.shape[1], src.shape[0], _vptr(out)))
    return out


def stage_harris(img, xs, ys, device=0):
    img = np.ascontiguousarray(img, np.uint8)
    xs, ys = np.ascontiguousarray(xs, np.int32), np.ascontiguousarray(ys, np.int32)
    out = np.empty(len(xs), np.float32)
    check(lib().gd_stage_harris(device, _vptr(img), img.shape[1], img.shape[0], _vptr(xs), _vptr(ys), len(xs), _vptr(out)))
    return out


def stage_hamming_crosscheck(d1, d2, device=0):
    d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
    cap = len(d1)
    q, t, d = (np.empty(cap, np.int32) for _ in range(3))
    n = C.c_int(0)
    check(lib().gd_stage_hamming_crosscheck(device, _vptr(d1), len(d1), _vptr(d2), len(d2), _vptr(q), _vptr(t), _vptr(d), cap, C.byref(n)))
    return [(int(q[i]), int(t[i]), int(d[i])) for i in range(n.value)]


class Orb:
    """Mirror of ORB_SLAM2::ORBextractor (include/ORBextractor.h:45-111) for `batch` images per call."""

    def __init__(self, nfeatures=1500, scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7, max_width=640,
                 max_height=480, device=0, batch=1):
        self.batch, self.nlevels, self.nfeatures = batch, nlevels, nfeatures
        self._h = vp()
        check(lib().gd_orb_create(C.byref(self._h), nfeatures, scale_factor, nlevels, ini_th_fast, min_th_fast, max_width,
                                  max_height, device, batch))

    def close(self):
        if self._h:
            lib().gd_orb_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __call__(self, grays):
        """operator(): list of 8UC1 images (same size) -> list of (keypoints[KP_DTYPE], descriptors[n,32])."""
        grays = [np.ascontiguousarray(g, np.uint8) for g in grays]
        assert len(grays) == self.batch
        h, w = grays[0].shape
        cap = self.nfeatures + 3 * self.nlevels + 8
        kps = [np.zeros(cap, KP_DTYPE) for _ in grays]
        desc = [np.zeros((cap, 32), np.uint8) for _ in grays]
        n = (C.c_int * self.batch)()
        check(lib().gd_orb_extract(self._h, _ptr_array(grays), w, w, h, _ptr_array(kps), _ptr_array(desc), cap, n))
        return [(kps[b][: n[b]].copy(), desc[b][: n[b]].copy()) for b in range(self.batch)]

    def level(self, level, stream=0):
        w, h = C.c_int(0), C.c_int(0)
        check(lib().gd_orb_level_size(self._h, level, C.byref(w), C.byref(h)))
        out = np.empty((h.value, w.value), np.uint8)
        check(lib().gd_orb_fetch_level(self._h, stream, level, _vptr(out), w.value, None, None))
        return out

    def features_per_level(self):
        n = (C.c_int * self.nlevels)()
        check(lib().gd_orb_features_per_level(self._h, n))
        return list(n)


def smoke_orb(po):
    """Tiny ORB extraction on cuda:0 checked bit-exactly against the oracle (used by __graft_entry__.smoke)."""
    import importlib

    synth = importlib.import_module("gd-slam_b200.synth")
    s = synth.SyntheticStream(0, 320, 240)
    gray = po.gray(s.frame(0).bgr, 1)
    orb = Orb(500, 1.2, 6, 20, 7, 320, 240, 0, 1)
    kp, desc = orb([gray])[0]
    rkp, rdesc, _ = po.orb_extract(gray, nfeatures=500, nlevels=6)
    ok = len(kp) == len(rkp) and all(np.array_equal(kp[f], rkp[f]) for f in kp.dtype.names) and np.array_equal(desc, rdesc)
    print(f"smoke: ORB 320x240 {len(kp)} keypoints, bit-exact vs oracle = {ok}")
    assert ok
    orb.close()


# ---------------------------------------------------------------------------------------------- batched front-end
class Frontend:
    """gd_frontend_*: GrabImageRGBD_GD's per-frame sequence (gray, ORB, AddNewImage, GetNoGMMmask) for `batch` streams."""

    def __init__(self, K, width=640, height=480, batch=1, device=0, dist=None, depth_factor=5000.0, nfeatures=1500,
                 scale_factor=1.2, nlevels=8, ini_th_fast=20, min_th_fast=7, orb_gray_order=1, staged_slots=0, getrt=False):
        cfg = FrontendConfig()
        K = np.ascontiguousarray(K, np.float32).reshape(-1)
        for i in range(9):
            cfg.K[i] = float(K[i])
        d = np.zeros(5, np.float32) if dist is None else np.ascontiguousarray(dist, np.float32).reshape(-1)
        for i in range(min(5, d.size)):
            cfg.dist[i] = float(d[i])
        cfg.ndist = 0 if dist is None else int(d.size)
        cfg.depth_factor = depth_factor
        cfg.width, cfg.height, cfg.device, cfg.batch = width, height, device, batch
        cfg.nfeatures, cfg.scale_factor, cfg.nlevels = nfeatures, scale_factor, nlevels
        cfg.ini_th_fast, cfg.min_th_fast, cfg.orb_gray_order = ini_th_fast, min_th_fast, orb_gray_order
        cfg.kp_capacity = nfeatures + 3 * nlevels
        cfg.staged_slots = staged_slots
        cfg.getrt = 1 if getrt else 0
        self._hook = None
        self.cfg = cfg
        self.w, self.h, self.batch, self.cap = width, height, batch, cfg.kp_capacity
        self._h = vp()
        check(lib().gd_frontend_create(C.byref(self._h), C.byref(cfg)))
        B = batch
        # result buffers: page-locked, densely packed per batch (one D2H copy for the masks)
        self.masks = pinned_empty((B, height, width), np.uint8)
        self.kps = pinned_empty((B, self.cap), KP_DTYPE)
        self.desc = pinned_empty((B, self.cap, 32), np.uint8)
        self.n_kp = np.zeros(B, np.int32)
        self._mask_ptrs = _ptr_array([self.masks[b] for b in range(B)])
        self._kp_ptrs = _ptr_array([self.kps[b] for b in range(B)])
        self._desc_ptrs = _ptr_array([self.desc[b] for b in range(B)])

    def close(self):
        if self._h:
            lib().gd_frontend_destroy(self._h)
            self._h = vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @staticmethod
    def _pose(R, T, pv, B):
        R = np.ascontiguousarray(np.tile(np.eye(3, dtype=np.float32), (B, 1, 1)) if R is None else R, np.float32)
        T = np.ascontiguousarray(np.zeros((B, 3), np.float32) if T is None else T, np.float32)
        pv = np.ascontiguousarray(np.ones(B, np.int32) if pv is None else pv, np.int32)
        return R, T, pv

    def set_pose_hook(self, fn):
        """fn(stream, object_points [n,3], image_pixels [n,2]) -> (R 3x3, T 3) or None: the caller's solvePnPRansac + Rodrigues
        (GeoMaskMaker.cc:148-150).  Used by step(..., use_hook=True)."""
        if fn is None:
            self._hook = None
            check(lib().gd_frontend_set_pose_hook(self._h, C.cast(None, POSE_HOOK), None))
            return

        def tramp(user, stream, obj, pix, n, Rp, Tp):
            o = np.ctypeslib.as_array(obj, shape=(n, 3)).copy()
            p = np.ctypeslib.as_array(pix, shape=(n, 2)).copy()
            res = fn(stream, o, p)
            if res is None:
                return 0
            Rm = np.asarray(res[0], np.float32).reshape(9)
            Tm = np.asarray(res[1], np.float32).reshape(3)
            for i in range(9):
                Rp[i] = float(Rm[i])
            for i in range(3):
                Tp[i] = float(Tm[i])
            return 1

        self._hook = POSE_HOOK(tramp)  # kept alive by the object
        check(lib().gd_frontend_set_pose_hook(self._h, self._hook, None))

    def fetch_getrt(self):
        """Points of the last step's GetRt stage: list of (object_points [n,3], image_pixels [n,2]) per stream."""
        B = self.batch
        obj = [np.zeros((100, 3), np.float32) for _ in range(B)]
        pix = [np.zeros((100, 2), np.float32) for _ in range(B)]
        n = np.zeros(B, np.int32)
        check(lib().gd_frontend_fetch_getrt(self._h, _ptr_array(obj), _ptr_array(pix), n.ctypes.data_as(ip)))
        return [(obj[b][: n[b]].copy(), pix[b][: n[b]].copy()) for b in range(B)]

    def step(self, bgr, depth, R=None, T=None, pose_valid=None, fetch=True, use_hook=False):
        """bgr: (B,H,W,3) u8 array or list of (H,W,3); depth: (B,H,W) f32 or list.  Host buffers in, results out.
        use_hook: pass no pose at all -> the pose comes from the GetRt stage + the installed pose hook."""
        B = self.batch
        bp = _ptr_array([bgr[b] for b in range(B)])
        dp = _ptr_array([depth[b] for b in range(B)])
        if use_hook:
            check(lib().gd_frontend_step(self._h, bp, self.w * 3, dp, self.w * 4, None, None, None,
                                         self._mask_ptrs if fetch else None, self.w, self._kp_ptrs if fetch else None,
                                         self._desc_ptrs if fetch else None, self.n_kp.ctypes.data_as(ip)))
            return self.results() if fetch else None
        R, T, pv = self._pose(R, T, pose_valid, B)
        check(lib().gd_frontend_step(self._h, bp, self.w * 3, dp, self.w * 4, _fptr(R), _fptr(T), pv.ctypes.data_as(ip),
                                     self._mask_ptrs if fetch else None, self.w, self._kp_ptrs if fetch else None,
                                     self._desc_ptrs if fetch else None, self.n_kp.ctypes.data_as(ip)))
        return self.results() if fetch else None

    def step_u16(self, bgr, depth_u16, R=None, T=None, pose_valid=None):
        """Same as step() with the raw 16-bit TUM depth (row f-4: converted on the device like Tracking.cc:234-235)."""
        B = self.batch
        R, T, pv = self._pose(R, T, pose_valid, B)
        bl = [np.ascontiguousarray(bgr[b], np.uint8) for b in range(B)]  # views of a packed (pinned) batch stay views
        dl = [np.ascontiguousarray(depth_u16[b], np.uint16) for b in range(B)]
        check(lib().gd_frontend_step_u16(self._h, _ptr_array(bl), self.w * 3, _ptr_array(dl), self.w * 2, _fptr(R), _fptr(T),
                                         pv.ctypes.data_as(ip), self._mask_ptrs, self.w, self._kp_ptrs, self._desc_ptrs,
                                         self.n_kp.ctypes.data_as(ip)))
        return self.results()

    def stage(self, slot, bgr, depth):
        B = self.batch
        bp = _ptr_array([bgr[b] for b in range(B)])
        dp = _ptr_array([depth[b] for b in range(B)])
        check(lib().gd_frontend_stage(self._h, slot, bp, self.w * 3, dp, self.w * 4))

    def step_staged(self, slot, R=None, T=None, pose_valid=None):
        R, T, pv = self._pose(R, T, pose_valid, self.batch)
        check(lib().gd_frontend_step_staged(self._h, slot, _fptr(R), _fptr(T), pv.ctypes.data_as(ip)))

    def fetch(self):
        check(lib().gd_frontend_fetch(self._h, self._mask_ptrs, self.w, self._kp_ptrs, self._desc_ptrs,
                                      self.n_kp.ctypes.data_as(ip)))
        return self.results()

    def fetch_stereo_grid(self, bf):
        """Row f-3 for the filtered keypoints: list of (mvDepth, mvuRight, cell_start[3073], cell_items) per stream."""
        B = self.batch
        d = [np.zeros(self.cap, np.float32) for _ in range(B)]
        ur = [np.zeros(self.cap, np.float32) for _ in range(B)]
        cs = [np.zeros(64 * 48 + 1, np.int32) for _ in range(B)]
        ci = [np.zeros(self.cap, np.int32) for _ in range(B)]
        un = [np.zeros((self.cap, 2), np.float32) for _ in range(B)]
        check(lib().gd_frontend_fetch_stereo_grid_un(self._h, bf, _ptr_array(d), _ptr_array(ur), _ptr_array(cs), _ptr_array(ci),
                                                     _ptr_array(un)))
        self.keys_un = un  # mvKeysUn positions of the filtered keypoints (rows beyond the count are unused)
        return [(d[b], ur[b], cs[b], ci[b][: cs[b][-1]].copy()) for b in range(B)]

    def results(self):
        return [(self.masks[b], self.kps[b][: self.n_kp[b]], self.desc[b][: self.n_kp[b]]) for b in range(self.batch)]

    def fetch_filtered(self):
        """Frame ctor filter (Frame.cc:258-282) with the new mask: list of (keypoints, descriptors) per stream."""
        B = self.batch
        kps = [np.zeros(self.cap, KP_DTYPE) for _ in range(B)]
        desc = [np.zeros((self.cap, 32), np.uint8) for _ in range(B)]
        n = np.zeros(B, np.int32)
        check(lib().gd_frontend_fetch_filtered(self._h, _ptr_array(kps), _ptr_array(desc), n.ctypes.data_as(ip)))
        return [(kps[b][: n[b]].copy(), desc[b][: n[b]].copy()) for b in range(B)]

    def sync(self):
        check(lib().gd_frontend_sync(self._h))

    def timer_begin(self):
        check(lib().gd_frontend_timer_begin(self._h))

    def timer_end(self):
        ms = C.c_float(0)
        check(lib().gd_frontend_timer_end(self._h, C.byref(ms)))
        return ms.value

    def launch_count(self):
        n = C.c_longlong(0)
        check(lib().gd_frontend_launch_count(self._h, C.byref(n)))
        return n.value

    def profile(self, enable):
        check(lib().gd_frontend_profile(self._h, 1 if enable else 0))

    def profile_read(self):
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        ln = (C.c_longlong * 64)()
        n = C.c_int(0)
        check(lib().gd_frontend_profile_read(self._h, 64, names, ms, ln, C.byref(n)))
        return [(names[i].decode(), ms[i], ln[i]) for i in range(min(n.value, 64))]

    def flush_l2(self):
        check(lib().gd_frontend_flush_l2(self._h))

    def debug(self, what, stream=0):
        shapes = {DBG_FLOW: ((self.h, self.w, 2), np.float32), DBG_DIST: ((self.h, self.w), np.float32),
                  DBG_EDGE_REF: ((self.h, self.w), np.uint8), DBG_EDGE_CUR: ((self.h, self.w), np.uint8),
                  DBG_GRAY_CUR: ((self.h, self.w), np.uint8), DBG_MINMAX: ((2,), np.float32)}
        shp, dt = shapes[what]
        out = np.empty(shp, dt)
        check(lib().gd_frontend_debug_fetch(self._h, what, stream, _vptr(out), out.nbytes))
        return out
