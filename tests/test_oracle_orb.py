"""ORB oracle (oracle/orb_oracle.cpp, orb_prims.hpp) vs cv2 known-answer vectors and vs the reference's own
ORBextractor.cc compiled verbatim (fixtures in tests/golden/, CPU only)."""
import numpy as np
import pytest


def test_fast_atan2_kat(oracle, golden):
    g = golden("prims.npz")
    mine = np.array([oracle.fast_atan2(y, x) for y, x in g["atan2_yx"]], np.float32)
    assert np.array_equal(mine, g["atan2_deg"])


def test_gray_kat(oracle, golden):
    g = golden("prims.npz")
    assert np.array_equal(oracle.gray(g["bgr_small"], 0), g["gray_bgr2gray"])
    assert np.array_equal(oracle.gray(g["bgr_small"], 1), g["gray_rgb2gray"])


def test_pyramid_resize_and_blur_kat(oracle, synth, golden):
    g = golden("prims.npz")
    f0 = synth.SyntheticStream(0).frame(0)
    assert synth.frame_crc(f0) == int(g["pyr_src_crc"][0])
    prev = oracle.gray(f0.bgr, 1)
    for l in range(1, 8):
        ref = g[f"pyr_L{l}"]
        prev = oracle.resize_u8(prev, ref.shape[1], ref.shape[0])
        assert np.array_equal(prev, ref), l
    assert np.array_equal(oracle.gaussian7(g["pyr_L3"]), g["blur_L3"])


def test_fast_cells_kat(oracle, golden):
    g = golden("prims.npz")
    for th, kp, idx in ((20, g["fast_kp20"], g["fast_idx20"]), (7, g["fast_kp7"], g["fast_idx7"])):
        for i, (cw, ch) in enumerate(g["fast_cell_sizes"]):
            cell = np.ascontiguousarray(g["fast_cells"][i][:ch, :cw])
            assert np.array_equal(oracle.fast_detect(cell, th), kp[idx[i]:idx[i + 1]]), (th, i)


def test_config_tables(oracle):
    cfg = oracle.orb_config()
    assert list(cfg["n_per_level"]) == [326, 271, 226, 189, 157, 131, 109, 91]
    assert list(cfg["umax"]) == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [tuple(v) for v in cfg["level_sizes"]] == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                      (257, 193), (214, 161), (179, 134)]


@pytest.mark.parametrize("case", ["640", "320", "rag", "lc"])
def test_extract_equals_verbatim_reference(oracle, synth, golden, case):
    g = golden("orb.npz")
    if case == "640":
        gray = oracle.gray(synth.SyntheticStream(0).frame(0).bgr, 1)
        nf = 1500
    elif case == "320":
        gray, nf = g["gray_320"], 1500
    elif case == "rag":
        gray = np.ascontiguousarray(oracle.gray(synth.SyntheticStream(0).frame(0).bgr, 1)[50:297, 100:433])
        nf = 1000
    else:
        gray, nf = g["gray_lc"], 1500
    kp, desc, _ = oracle.orb_extract(gray, nfeatures=nf)
    ref_kp, ref_desc = g[f"kp_{case}"], g[f"desc_{case}"]
    assert len(kp) == len(ref_kp)
    for f in kp.dtype.names:
        assert np.array_equal(kp[f], ref_kp[f]), f
    assert np.array_equal(desc, ref_desc)


def test_live_reference_when_built(oracle, synth):
    """Where oracle/_ref exists (build container, or shipped to the GPU box) compare live on a fresh frame."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/liborbref.so not built")
    gray = oracle.gray(synth.SyntheticStream(5).frame(3).bgr, 1)
    kp, desc, pyr = oracle.orb_extract(gray, want_pyramid=True)
    rkp, rdesc, rpyr = oracle.orbref_extract(gray, want_pyramid=True)
    assert np.array_equal(kp, rkp) and np.array_equal(desc, rdesc)
    for a, b in zip(pyr, rpyr):
        assert np.array_equal(a, b)


def test_sincos_contract_vs_libm_sweep():
    """Contract: a = (float)cos((double)angle).  Report (do not require) agreement with this box's cosf/sinf (SURVEY B-5)."""
    import ctypes
    import ctypes.util

    m = ctypes.CDLL(ctypes.util.find_library("m"))
    m.cosf.restype = ctypes.c_float
    m.cosf.argtypes = [ctypes.c_float]
    ang = (np.arange(0, 36000, dtype=np.float32) * np.float32(0.01)) * np.float32(np.pi / 180)
    contract = np.cos(ang.astype(np.float64)).astype(np.float32)
    libm = np.array([m.cosf(float(a)) for a in ang], np.float32)
    diff = np.abs(contract.view(np.int32).astype(np.int64) - libm.view(np.int32).astype(np.int64))
    assert diff.max() <= 1  # never more than one ulp apart
    print("cosf vs correctly-rounded: %d / %d differ by 1 ulp" % ((diff > 0).sum(), diff.size))


def test_live_reference_random_sizes_and_parameters(oracle):
    """Randomised pin of the restated ORBextractor against the verbatim-compiled reference: image sizes that are not
    multiples of anything, different feature budgets / level counts / thresholds, textures from sparse to dense."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref/liborbref.so not built")
    rs = np.random.RandomState(20240517)
    checked = 0
    for case in range(10):
        w, h = int(rs.randint(150, 700)), int(rs.randint(120, 520))
        nlevels = int(rs.randint(2, 6)) if min(w, h) < 250 else int(rs.randint(4, 9))
        nf = int(rs.choice([300, 800, 1500, 2500]))
        ini, mn = [(20, 7), (30, 10), (12, 5)][int(rs.randint(0, 3))]
        img = rs.rand(h, w).astype(np.float32)
        k = int(rs.choice([1, 2, 4]))  # box-smooth: controls corner density
        if k > 1:
            pad = np.pad(img, k, mode="edge")
            acc = np.zeros_like(img)
            for dy in range(2 * k + 1):
                for dx in range(2 * k + 1):
                    acc += pad[dy:dy + h, dx:dx + w]
            img = acc
        img = ((img - img.min()) / (img.max() - img.min()) * rs.choice([255.0, 90.0])).astype(np.uint8)
        try:
            kp, desc, _ = oracle.orb_extract(img, nfeatures=nf, nlevels=nlevels, ini_th=ini, min_th=mn)
        except ValueError:
            continue  # level smaller than one cell / portrait level with nIni == 0: the reference has undefined behaviour
        rkp, rdesc, _ = oracle.orbref_extract(img, nfeatures=nf, nlevels=nlevels, ini_th=ini, min_th=mn)
        assert len(kp) == len(rkp), (case, w, h, nlevels, nf, len(kp), len(rkp))
        assert np.array_equal(kp, rkp) and np.array_equal(desc, rdesc), (case, w, h, nlevels, nf)
        checked += 1
    assert checked >= 6
