#!/usr/bin/env python
"""Golden fixtures for the ORBextractor half of the path (run in the build container: needs cv2 4.13 and
oracle/_ref/liborbref.so = the reference's src/ORBextractor.cc compiled verbatim, see oracle/Makefile).

  prims.npz  known-answer vectors of the OpenCV primitives ORBextractor calls, produced by cv2 itself:
             cvtColor (both byte orders), fastAtan2, FAST(th, nonmax) on random cells, resize(INTER_LINEAR) for the
             seven pyramid transitions of a 640x480 frame, GaussianBlur 7x7 sigma 2 REFLECT_101.
  orb.npz    keypoints + descriptors of the reference's own ORBextractor (verbatim compile) on seeded inputs:
             640x480 synthetic frame (TUM3 ORB settings 1500/1.2/8/20/7), a 320x240 one, a ragged 333x247 one and a
             low-contrast frame that exercises the minThFAST fallback and levels with fewer candidates than requested.
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
from oracle import pyoracle as po  # noqa: E402


def low_contrast(w, h, seed):
    rs = np.random.RandomState(seed)
    base = cv2.GaussianBlur(rs.randint(0, 256, (h, w)).astype(np.uint8), (0, 0), 3.0)
    img = (base.astype(np.float32) - 128) * 2.2 + 128
    img[:, : w // 3] = (img[:, : w // 3] - 128) * 0.25 + 128  # left third: only the threshold-7 pass fires
    img[: h // 3, w // 2:] = 90  # a flat block: no corners at all
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def make_prims():
    out = {}
    s = synth.SyntheticStream(0)
    f0 = s.frame(0)
    small = np.ascontiguousarray(f0.bgr[100:196, 200:328])
    out["bgr_small"] = small
    out["gray_bgr2gray"] = cv2.cvtColor(small, cv2.COLOR_BGR2GRAY)
    out["gray_rgb2gray"] = cv2.cvtColor(small, cv2.COLOR_RGB2GRAY)
    rs = np.random.RandomState(7)
    yx = rs.randint(-200000, 200000, size=(4000, 2)).astype(np.float32)
    yx[:50] = rs.randint(-3, 4, size=(50, 2))
    out["atan2_yx"] = yx
    out["atan2_deg"] = np.array([cv2.fastAtan2(float(y), float(x)) for y, x in yx], np.float32)
    g = cv2.cvtColor(f0.bgr, cv2.COLOR_RGB2GRAY)
    out["pyr_src_crc"] = np.array([synth.frame_crc(f0)], np.uint64)
    cfg = po.orb_config()
    prev = g
    for l in range(1, 8):
        lw, lh = [int(v) for v in cfg["level_sizes"][l]]
        prev = cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_LINEAR)
        out[f"pyr_L{l}"] = prev
    out["blur_L3"] = cv2.GaussianBlur(out["pyr_L3"], (7, 7), 2, 2, borderType=cv2.BORDER_REFLECT_101)
    # FAST on 40 random cells (both thresholds), cv2 keypoints as (x, y, response)
    cells, res20, res7, idx20, idx7 = [], [], [], [0], [0]
    for i in range(40):
        cw, ch = rs.randint(20, 44), rs.randint(20, 44)
        x0, y0 = rs.randint(0, 640 - cw), rs.randint(0, 480 - ch)
        cell = np.ascontiguousarray(g[y0:y0 + ch, x0:x0 + cw])
        if i % 5 == 0:
            cell = (cell // 6 + 100).astype(np.uint8)  # low contrast: only threshold 7 finds corners
        cells.append(np.pad(cell, ((0, 44 - ch), (0, 44 - cw))))
        out.setdefault("fast_cell_sizes", []).append((cw, ch))
        for th, res, idx in ((20, res20, idx20), (7, res7, idx7)):
            k = cv2.FastFeatureDetector_create(th, True).detect(cell)
            res.extend([(int(p.pt[0]), int(p.pt[1]), int(p.response)) for p in k])
            idx.append(len(res))
    out["fast_cells"] = np.stack(cells)
    out["fast_cell_sizes"] = np.array(out["fast_cell_sizes"], np.int32)
    out["fast_kp20"] = np.array(res20, np.int32).reshape(-1, 3)
    out["fast_kp7"] = np.array(res7, np.int32).reshape(-1, 3)
    out["fast_idx20"] = np.array(idx20, np.int32)
    out["fast_idx7"] = np.array(idx7, np.int32)
    np.savez_compressed(os.path.join(HERE, "prims.npz"), **out)
    print("prims: fast kps", len(res20), len(res7))


def make_orb():
    assert po.have_ref(), "build oracle/_ref first (make -C oracle ref)"
    out = {}
    s = synth.SyntheticStream(0)
    g640 = cv2.cvtColor(s.frame(0).bgr, cv2.COLOR_RGB2GRAY)
    kp, d, _ = po.orbref_extract(g640)
    out["kp_640"], out["desc_640"] = kp, d
    s2 = synth.SyntheticStream(1, 320, 240)
    g320 = cv2.cvtColor(s2.frame(4).bgr, cv2.COLOR_RGB2GRAY)
    kp, d, _ = po.orbref_extract(g320)
    out["gray_320"], out["kp_320"], out["desc_320"] = g320, kp, d
    rag = np.ascontiguousarray(g640[50:297, 100:433])
    kp, d, _ = po.orbref_extract(rag, nfeatures=1000)
    out["kp_rag"], out["desc_rag"] = kp, d
    lc = low_contrast(640, 480, 11)
    kp, d, _ = po.orbref_extract(lc)
    out["gray_lc"], out["kp_lc"], out["desc_lc"] = lc, kp, d
    np.savez_compressed(os.path.join(HERE, "orb.npz"), **out)
    print("orb: n640", len(out["kp_640"]), "n320", len(out["kp_320"]), "rag", len(out["kp_rag"]), "lc", len(out["kp_lc"]),
          np.bincount(out["kp_lc"]["octave"], minlength=8))


if __name__ == "__main__":
    make_prims()
    make_orb()
