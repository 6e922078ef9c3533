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


def test_cv_orb_order_and_full_getrt_equal_the_cv2_transcription(proto, oracle, synth):
    """The keypoint ORDER of cv::ORB (std::nth_element inside retainBest) and the first-100 selection (std::sort) are
    reproduced by running the same libstdc++ routines on the same sequences; with cv2.solvePnPRansac as the oracle of the
    last step the restated GetRt returns the same R, T as a line-by-line cv2 transcription of GeoMaskMaker.cc:77-156."""
    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    K = synth.intrinsics()
    g = [cv2.cvtColor(f.bgr, cv2.COLOR_BGR2GRAY) for f in (f0, f5)]  # detectAndCompute converts BGR input like this
    orb = cv2.ORB_create(2000, 1.2, 8, 31, 0, 2)
    k1, d1 = orb.detectAndCompute(f0.bgr, None)
    k2, d2 = orb.detectAndCompute(f5.bgr, None)
    # (a) order of the features
    mine = proto.cv_orb_detect_and_compute(g[0], oracle.fast_detect, oracle.ic_angle, oracle.orb_descriptor,
                                           retain_best_order=oracle.retain_best_order)
    assert [(m[0], float(m[1]), float(m[2])) for m in mine] == [(k.octave, k.pt[0], k.pt[1]) for k in k1]
    assert np.array_equal(np.stack([m[5] for m in mine]), d1)

    # (b) the reference, transcribed with cv2 calls (C++ std::sort replaced by the oracle's call of the same routine)
    def solve(obj, pix, Kf):
        ok, rvec, tvec, _ = cv2.solvePnPRansac(obj, pix, Kf, np.zeros((4, 1), np.float32))
        R, _ = cv2.Rodrigues(rvec)
        return R, tvec

    m = cv2.BFMatcher(cv2.NORM_HAMMING, True).match(d1, d2)
    order = oracle.sort_matches_order(np.array([x.distance for x in m], np.float32))[:100]
    Ki = cv2.invert(K)[1]
    obj, pix = [], []
    for i in order:
        x, y = k1[m[i].queryIdx].pt
        d = f0.depth_m[int(y), int(x)]
        if d == 0:
            continue
        P = cv2.gemm(Ki, np.array([[x], [y], [1.0]], np.float32), 1.0, None, 0.0) * np.float32(d)
        obj.append(P.reshape(3))
        pix.append(k2[m[i].trainIdx].pt)
    R_ref, T_ref = solve(np.asarray(obj, np.float32), np.asarray(pix, np.float32), K)
    ok, R, T = proto.get_rt(g[0], g[1], f0.depth_m, K, oracle, solve)
    assert ok
    assert np.array_equal(R, R_ref.astype(np.float32)) and np.array_equal(T, T_ref.astype(np.float32).reshape(3))
