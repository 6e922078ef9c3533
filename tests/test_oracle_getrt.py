"""SURVEY 8(f)-1 groundwork (CPU only): the numpy restatement of the feature / matching half of GeoMaskMaker::GetRt
(oracle/getrt_proto.py) pinned against cv2 4.13 live.  No product path exists for this row yet."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")


@pytest.fixture(scope="module")
def proto():
    from oracle import getrt_proto
    return getrt_proto


def _cv_fast(img, th):
    k = cv2.FastFeatureDetector_create(th, True).detect(img, None)
    return [(p.pt[0], p.pt[1], p.response) for p in k]


def test_resize_linear_exact(proto):
    rs = np.random.RandomState(1)
    src = rs.randint(0, 256, (120, 160), np.uint8)
    for dw, dh in ((133, 100), (80, 60), (159, 119), (97, 73)):
        assert np.array_equal(proto.resize_linear_exact(src, dw, dh), cv2.resize(src, (dw, dh), interpolation=cv2.INTER_LINEAR_EXACT))


def test_cv_orb_features_as_a_set_and_matcher(proto, oracle, synth):
    s = synth.SyntheticStream(0, 320, 240)
    imgs = [cv2.cvtColor(s.frame(f).bgr, cv2.COLOR_BGR2GRAY) for f in (0, 5)]
    orb = cv2.ORB_create(500, 1.2, 8, 31, 0, 2)
    descs = []
    for img in imgs:
        kps, desc = orb.detectAndCompute(img, None)
        mine = proto.cv_orb_detect_and_compute(img, _cv_fast, oracle.ic_angle, oracle.orb_descriptor, nfeatures=500)
        assert len(mine) == len(kps) > 300
        ref = {(k.octave, np.float32(k.pt[0]).tobytes(), np.float32(k.pt[1]).tobytes()): (np.float32(k.response), np.float32(k.angle), d)
               for k, d in zip(kps, desc)}
        for l, x, y, r, a, d in mine:
            key = (l, x.tobytes(), y.tobytes())
            assert key in ref
            rr, ra, rd = ref[key]
            assert rr == r and ra == a and np.array_equal(rd, d), key
        descs.append(desc)
    m_cv = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(descs[0], descs[1])
    m_me = proto.bf_match_hamming_crosscheck(descs[0], descs[1])
    assert [(m.queryIdx, m.trainIdx, int(m.distance)) for m in m_cv] == m_me and len(m_me) > 100
