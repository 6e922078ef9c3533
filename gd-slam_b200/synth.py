"""Deterministic synthetic RGB-D streams for the GeoMaskMaker + ORB hot path.

Generator contract: seeded white noise -> Gaussian sigma=2.5 -> stretch
texture on a slanted plane, one textured "dynamic object" in front of the plane with its
own image motion, camera translating (and optionally rolling) between frames, depth holes
(discs) and a far band (> 3.5 m), depth quantised to uint16 * 1/5000 exactly like
Tracking.cc:234-235 of the reference.

Everything is built from element-wise IEEE operations in a fixed order (no library
reductions, no library filters) so the same seed yields the same bytes on every machine.
"""
from __future__ import annotations

import zlib
from dataclasses import dataclass

import numpy as np

TUM3 = dict(fx=535.4, fy=539.2, cx=320.1, cy=247.6)


def intrinsics(width: int = 640, height: int = 480) -> np.ndarray:
    """TUM3.yaml intrinsics (Examples/RGB-D/TUM3.yaml:8-16) scaled by width/640."""
    s = width / 640.0
    K = np.array(
        [[TUM3["fx"] * s, 0.0, TUM3["cx"] * s], [0.0, TUM3["fy"] * s, TUM3["cy"] * s], [0.0, 0.0, 1.0]],
        dtype=np.float32,
    )
    return K


def _gauss_taps(sigma: float, radius: int) -> np.ndarray:
    x = np.arange(-radius, radius + 1, dtype=np.float64)
    g = np.exp(-(x * x) / (2.0 * sigma * sigma))
    s = 0.0
    for v in g:  # fixed-order sum
        s += float(v)
    return g / s


def _blur_wrap(img: np.ndarray, sigma: float) -> np.ndarray:
    """Separable circular Gaussian, explicit tap loop (deterministic)."""
    r = int(np.ceil(3.0 * sigma))
    g = _gauss_taps(sigma, r)
    out = np.zeros_like(img, dtype=np.float64)
    for k in range(-r, r + 1):
        out = out + g[k + r] * np.roll(img, k, axis=1)
    out2 = np.zeros_like(out)
    for k in range(-r, r + 1):
        out2 = out2 + g[k + r] * np.roll(out, k, axis=0)
    return out2


def texture(seed: int, height: int, width: int, channels: int = 3) -> np.ndarray:
    """Periodic texture, float64 in [0,255], shape (height, width, channels)."""
    rs = np.random.RandomState(seed & 0x7FFFFFFF)
    out = np.empty((height, width, channels), dtype=np.float64)
    for c in range(channels):
        n = rs.randint(0, 256, size=(height, width)).astype(np.float64)
        b = _blur_wrap(n, 2.5)
        # fixed gain instead of a data-dependent stretch: no library reduction, so the bytes are
        # machine independent.  White noise (std 73.9) blurred with sigma=2.5 has std ~8.3; x7 -> ~58.
        out[:, :, c] = np.clip(127.5 + (b - 127.5) * 7.0, 0.0, 255.0)
    return out


def _bilinear_wrap(tex: np.ndarray, u: np.ndarray, v: np.ndarray) -> np.ndarray:
    th, tw = tex.shape[:2]
    u0 = np.floor(u)
    v0 = np.floor(v)
    fu = (u - u0)[..., None]
    fv = (v - v0)[..., None]
    iu0 = np.mod(u0.astype(np.int64), tw)
    iv0 = np.mod(v0.astype(np.int64), th)
    iu1 = np.mod(iu0 + 1, tw)
    iv1 = np.mod(iv0 + 1, th)
    a = tex[iv0, iu0] * (1.0 - fu) + tex[iv0, iu1] * fu
    b = tex[iv1, iu0] * (1.0 - fu) + tex[iv1, iu1] * fu
    return a * (1.0 - fv) + b * fv


@dataclass
class Frame:
    bgr: np.ndarray  # (H, W, 3) uint8, BGR byte order as imread delivers (rgbd_tum.cc:118)
    depth_u16: np.ndarray  # (H, W) uint16, TUM raw depth (metres * 5000)
    depth_m: np.ndarray  # (H, W) float32 metres, = u16 * (1/5000.f)  (Tracking.cc:234-235)
    R_w: np.ndarray  # world->camera rotation of this frame (float64 3x3)
    T_w: np.ndarray  # world->camera translation of this frame (float64 3)


class SyntheticStream:
    """One synthetic RGB-D sequence; frame f is a pure function of (stream, f, options)."""

    def __init__(self, stream: int = 0, width: int = 640, height: int = 480, roll_deg_per_frame: float = 0.0,
                 object_gap_m: float = 0.05, t_per_frame=(0.002, -0.0012, 0.0)):
        self.stream = stream
        self.w, self.h = width, height
        self.K = intrinsics(width, height).astype(np.float64)
        self.s = width / 640.0
        self.roll = roll_deg_per_frame
        self.gap = object_gap_m
        self.t = np.asarray(t_per_frame, dtype=np.float64)
        # periodic textures: background and object
        self.tex_bg = texture(1000 * stream + 17, 512, 1024)
        self.tex_obj = texture(1000 * stream + 91, 256, 256)

    # plane in frame-0 camera pixel coordinates (scaled so every resolution sees the same slope)
    def _plane(self, u0, v0):
        return 1.2 + 0.0015 * (u0 / self.s) + 0.001 * (v0 / self.s)

    def pose(self, f: int):
        """P_cam(f) = R_w P_world + T_w (world = camera of frame 0)."""
        th = np.deg2rad(self.roll * f)
        c, s = np.cos(th), np.sin(th)
        R = np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])
        T = self.t * f
        return R, T

    def pair_pose(self, f_ref: int, f_cur: int):
        """R, T (float32) with P_cur = R P_ref + T — the pose GetRt would estimate (GeoMaskMaker.cc:77-156)."""
        R0, T0 = self.pose(f_ref)
        R1, T1 = self.pose(f_cur)
        R = R1 @ R0.T
        T = T1 - R @ T0
        return R.astype(np.float32), T.astype(np.float32)

    def frame(self, f: int) -> Frame:
        w, h = self.w, self.h
        fx, fy, cx, cy = self.K[0, 0], self.K[1, 1], self.K[0, 2], self.K[1, 2]
        R, T = self.pose(f)
        uu, vv = np.meshgrid(np.arange(w, dtype=np.float64), np.arange(h, dtype=np.float64))
        xn = (uu - cx) / fx
        yn = (vv - cy) / fy
        z = np.full((h, w), 1.2 + 0.0015 * 320 + 0.001 * 240, dtype=np.float64)
        u0 = uu
        v0 = vv
        for _ in range(4):  # fixed-point: depth of the plane seen through this pixel
            px = xn * z - T[0]
            py = yn * z - T[1]
            pz = z - T[2]
            # P0 = R^T (P - T)
            qx = R[0, 0] * px + R[1, 0] * py + R[2, 0] * pz
            qy = R[0, 1] * px + R[1, 1] * py + R[2, 1] * pz
            qz = R[0, 2] * px + R[1, 2] * py + R[2, 2] * pz
            u0 = fx * qx / qz + cx
            v0 = fy * qy / qz + cy
            z0 = self._plane(u0, v0)
            z = z0 + (z - qz)  # keep the (tiny) depth change of the rigid motion consistent
        col = _bilinear_wrap(self.tex_bg, u0 / self.s, v0 / self.s)
        depth = z.copy()

        # far band (> 3.5 m cut of GeoMaskMaker.cc:229,870), fixed in frame-0 coordinates
        band_lo = 520.0 * self.s
        band = (u0 >= band_lo) & (u0 < band_lo + 40.0 * self.s)
        depth[band] = 4.0

        # dynamic object: 160x120 px (at 640x480), own texture, in front of the plane, +3 px/frame in x
        ow, oh = int(160 * self.s), int(120 * self.s)
        ox = 180.0 * self.s + 3.0 * self.s * f + fx * T[0] / 1.4
        oy = 170.0 * self.s + fy * T[1] / 1.4
        inside = (uu >= ox) & (uu < ox + ow) & (vv >= oy) & (vv < oy + oh)
        ocol = _bilinear_wrap(self.tex_obj, (uu - ox) / self.s, (vv - oy) / self.s)
        col = np.where(inside[..., None], ocol, col)
        depth = np.where(inside, z - self.gap, depth)

        # zero-depth disc holes (sensor dropouts), 20 per frame
        rs = np.random.RandomState((1000 * self.stream + f) & 0x7FFFFFFF)
        for _ in range(20):
            hx = rs.randint(20, w - 20)
            hy = rs.randint(20, h - 20)
            hr = rs.randint(4, 13) * self.s
            depth[(uu - hx) ** 2 + (vv - hy) ** 2 <= hr * hr] = 0.0

        bgr = np.clip(np.rint(col), 0, 255).astype(np.uint8)
        d16 = np.clip(np.rint(depth * 5000.0), 0, 65535).astype(np.uint16)
        dm = d16.astype(np.float32) * np.float32(1.0 / 5000.0)
        return Frame(np.ascontiguousarray(bgr), d16, dm, R, T)


def frame_crc(fr: Frame) -> int:
    return zlib.crc32(fr.depth_u16.tobytes(), zlib.crc32(fr.bgr.tobytes()))
