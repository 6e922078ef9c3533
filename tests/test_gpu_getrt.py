"""GPU parity (-m gpu): building blocks of GeoMaskMaker::GetRt through the C ABI vs oracle/getrt_proto.py
(which tests/test_oracle_getrt.py pins against cv2).  Bit-exact."""
import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def capi():
    c = load_pkg("capi")
    c.lib()
    assert c.device_count() >= 1
    return c


@pytest.fixture(scope="module")
def proto():
    from oracle import getrt_proto
    return getrt_proto


@pytest.fixture(scope="module")
def gray(oracle, synth):
    return oracle.gray(synth.SyntheticStream(0, 320, 240).frame(0).bgr, 0)


def test_resize_linear_exact(capi, proto, gray):
    prev = gray
    for l in range(1, 6):
        sc = np.float32(np.power(float(np.float32(1.2)), float(l)))
        dw, dh = int(np.rint(320 / sc)), int(np.rint(240 / sc))
        ref = proto.resize_linear_exact(prev, dw, dh)
        assert np.array_equal(capi.stage_resize_linear_exact(prev, dw, dh), ref), l
        prev = ref
    rs = np.random.RandomState(3)
    src = rs.randint(0, 256, (97, 131), np.uint8)
    for dw, dh in ((131, 97), (77, 58), (200, 150), (1, 1)):
        assert np.array_equal(capi.stage_resize_linear_exact(src, dw, dh), proto.resize_linear_exact(src, dw, dh)), (dw, dh)


def test_gaussian7_float_path(capi, proto, gray, oracle):
    out = capi.stage_gaussian7_float(gray)
    assert np.array_equal(out, proto.gaussian7_float(gray))
    # it is NOT the 8-bit fixed-point blur ORBextractor gets (cv::GaussianBlur picks that one only for non-submatrix inputs)
    assert (out != oracle.gaussian7(gray)).mean() > 0.005
    rag = np.ascontiguousarray(gray[5:106, 7:150])
    assert np.array_equal(capi.stage_gaussian7_float(rag), proto.gaussian7_float(rag))


def test_harris_responses(capi, proto, gray):
    rs = np.random.RandomState(5)
    xs = rs.randint(4, 320 - 4, 3000).astype(np.int32)
    ys = rs.randint(4, 240 - 4, 3000).astype(np.int32)
    ref = proto.harris_responses(gray, xs, ys)
    out = capi.stage_harris(gray, xs, ys)
    assert np.array_equal(out.view(np.uint32), ref.view(np.uint32))
    assert (ref != 0).mean() > 0.9


def test_hamming_crosscheck_matcher(capi, proto):
    rs = np.random.RandomState(7)
    d2 = rs.randint(0, 256, (1500, 32), np.uint8)
    d1 = d2[rs.permutation(1500)[:1200]].copy()
    flip = rs.rand(*d1.shape) < 0.04  # a few flipped bytes -> near matches, many distance ties among the rest
    d1[flip] ^= rs.randint(1, 256, flip.sum()).astype(np.uint8)
    d1[::50] = rs.randint(0, 256, (len(d1[::50]), 32), np.uint8)  # some queries without a partner
    d2[1::97] = d2[0]  # duplicated train descriptors: ties resolved to the smallest index
    ref = proto.bf_match_hamming_crosscheck(d1, d2)
    assert capi.stage_hamming_crosscheck(d1, d2) == ref and 800 < len(ref) < 1200


def test_fast_whole_level_with_nms(capi, oracle, gray):
    """cv::FAST(20, nonmax) on a whole level: same corners, same order, response = S' - 1 (oracle.fast_detect == cv2)."""
    for img in (gray, np.ascontiguousarray(gray[3:200, 5:278]), np.ascontiguousarray(gray[::2, ::2])):
        ref = oracle.fast_detect(img, 20)
        out = capi.stage_fast_whole(img, 20)
        out[:, 2] -= 1
        assert len(ref) > 200 and np.array_equal(out, ref)
    flat = np.full((64, 80), 90, np.uint8)
    assert len(capi.stage_fast_whole(flat, 20)) == 0
