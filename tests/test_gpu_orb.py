"""GPU parity (-m gpu): ORBextractor kernels through the C ABI vs the CPU oracle / verbatim-reference goldens. Bit-exact."""
import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    c = load_pkg("capi")
    c.lib()
    assert c.device_count() >= 1, "no CUDA device: the GPU tests must run on the B200 box"
    return c


@pytest.fixture(scope="module")
def gray640(synth, oracle):
    return oracle.gray(synth.SyntheticStream(0).frame(0).bgr, 1)


def _assert_same(kp, desc, rkp, rdesc):
    assert len(kp) == len(rkp), (len(kp), len(rkp))
    for f in kp.dtype.names:
        bad = np.nonzero(kp[f] != rkp[f])[0]
        assert bad.size == 0, (f, bad[:5], kp[f][bad[:5]], rkp[f][bad[:5]])
    assert np.array_equal(desc, rdesc)


def test_pyramid_bit_exact(capi, oracle, gray640, golden):
    levels = capi.stage_orb_pyramid(gray640)
    g = golden("prims.npz")
    assert np.array_equal(levels[0], gray640)
    for l in range(1, 8):
        assert np.array_equal(levels[l], g[f"pyr_L{l}"]), l  # cv2.resize chain


def test_gaussian7_bit_exact(capi, oracle, gray640, golden):
    g = golden("prims.npz")
    assert np.array_equal(capi.stage_gaussian7(g["pyr_L3"]), g["blur_L3"])
    assert np.array_equal(capi.stage_gaussian7(gray640), oracle.gaussian7(gray640))
    rag = np.ascontiguousarray(gray640[:133, :211])
    assert np.array_equal(capi.stage_gaussian7(rag), oracle.gaussian7(rag))


def test_fast_cells_candidates_bit_exact(capi, oracle, gray640, golden):
    for img in (gray640, golden("orb.npz")["gray_lc"], np.ascontiguousarray(gray640[20:267, 40:373])):
        mine = capi.stage_fast_cells(img)
        ref = oracle.orb_candidates(img)
        assert mine.shape == ref.shape, (mine.shape, ref.shape)
        assert np.array_equal(mine, ref)
    assert len(ref) > 500


@pytest.mark.parametrize("case", ["640", "320", "rag", "lc"])
def test_extract_vs_verbatim_reference_golden(capi, oracle, gray640, golden, case):
    g = golden("orb.npz")
    if case == "640":
        gray, nf = gray640, 1500
    elif case == "320":
        gray, nf = g["gray_320"], 1500
    elif case == "rag":
        gray, nf = np.ascontiguousarray(gray640[50:297, 100:433]), 1000
    else:
        gray, nf = g["gray_lc"], 1500
    orb = capi.Orb(nf, 1.2, 8, 20, 7, gray.shape[1], gray.shape[0], 0, 1)
    kp, desc = orb([gray])[0]
    _assert_same(kp, desc, g[f"kp_{case}"], g[f"desc_{case}"])
    assert orb.features_per_level() == list(oracle.orb_config(gray.shape[1], gray.shape[0], nf)["n_per_level"])
    orb.close()


def test_extract_batch_of_streams_vs_oracle(capi, oracle, synth):
    """Four different streams in one handle, several frames: keypoints, order, angles and descriptors bit-exact."""
    streams = [synth.SyntheticStream(s, roll_deg_per_frame=0.04 if s % 2 else 0.0) for s in range(4)]
    orb = capi.Orb(1500, 1.2, 8, 20, 7, 640, 480, 0, batch=4)
    for f in (0, 3):
        grays = [oracle.gray(s.frame(f).bgr, 1) for s in streams]
        res = orb(grays)
        for b in range(4):
            rkp, rdesc, rpyr = oracle.orb_extract(grays[b], want_pyramid=True)
            _assert_same(res[b][0], res[b][1], rkp, rdesc)
            assert np.array_equal(orb.level(7, b), rpyr[7])
    orb.close()


def test_extract_other_settings_and_resolution(capi, oracle, synth):
    gray = oracle.gray(synth.SyntheticStream(0, 1280, 720).frame(1).bgr, 1)
    orb = capi.Orb(2000, 1.2, 8, 20, 7, 1280, 720, 0, 1)
    kp, desc = orb([gray])[0]
    rkp, rdesc, _ = oracle.orb_extract(gray, nfeatures=2000)
    _assert_same(kp, desc, rkp, rdesc)
    # a smaller image through the same handle (re-plan), fewer levels worth of features
    small = np.ascontiguousarray(gray[:480, :640])
    kp, desc = orb([small])[0]
    rkp, rdesc, _ = oracle.orb_extract(small, nfeatures=2000)
    _assert_same(kp, desc, rkp, rdesc)
    orb.close()


def test_capacity_error_is_reported(capi, gray640):
    import ctypes as C

    orb = capi.Orb(1500, 1.2, 8, 20, 7, 640, 480, 0, 1)
    kps = np.zeros(10, capi.KP_DTYPE)
    desc = np.zeros((10, 32), np.uint8)
    n = (C.c_int * 1)()
    code = capi.lib().gd_orb_extract(orb._h, capi._ptr_array([gray640]), 640, 640, 480, capi._ptr_array([kps]),
                                     capi._ptr_array([desc]), 10, n)
    assert code == capi.GD_ECAPACITY and n[0] > 1000
    orb.close()


def test_random_sizes_parameters_and_rejected_inputs(capi, oracle):
    """Randomised sizes / feature budgets / level counts / thresholds / corner densities: GPU == oracle bit for bit, and the
    inputs on which the reference has undefined behaviour (portrait level with nIni == 0, level smaller than one cell) are
    rejected with an error by both."""
    rs = np.random.RandomState(20240517)
    checked = rejected = 0
    for case in range(10):
        w, h = int(rs.randint(150, 700)), int(rs.randint(120, 520))
        nlevels = int(rs.randint(2, 6)) if min(w, h) < 250 else int(rs.randint(4, 9))
        nf = int(rs.choice([300, 800, 1500, 2500]))
        ini, mn = [(20, 7), (30, 10), (12, 5)][int(rs.randint(0, 3))]
        img = rs.rand(h, w).astype(np.float32)
        k = int(rs.choice([1, 2, 4]))
        if k > 1:
            pad = np.pad(img, k, mode="edge")
            acc = np.zeros_like(img)
            for dy in range(2 * k + 1):
                for dx in range(2 * k + 1):
                    acc += pad[dy:dy + h, dx:dx + w]
            img = acc
        img = ((img - img.min()) / (img.max() - img.min()) * rs.choice([255.0, 90.0])).astype(np.uint8)
        try:
            rkp, rdesc, _ = oracle.orb_extract(img, nfeatures=nf, nlevels=nlevels, ini_th=ini, min_th=mn)
        except ValueError:
            with pytest.raises(capi.GdError):
                capi.Orb(nf, 1.2, nlevels, ini, mn, w, h, 0, 1)
            rejected += 1
            continue
        orb = capi.Orb(nf, 1.2, nlevels, ini, mn, w, h, 0, 1)
        for _ in range(3):  # third call replays the CUDA graph of the launch sequence
            kp, desc = orb([img])[0]
            _assert_same(kp, desc, rkp, rdesc)
        orb.close()
        checked += 1
    assert checked >= 6 and rejected >= 1
