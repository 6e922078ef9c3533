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


def _proto_features(proto, oracle, img, nf):
    feats = proto.cv_orb_detect_and_compute(img, oracle.fast_detect, oracle.ic_angle, oracle.orb_descriptor, nfeatures=nf,
                                            retain_best_order=oracle.retain_best_order)
    return feats


def test_cvorb_detect_and_compute_in_opencv_order(capi, proto, oracle, synth):
    """cv::ORB::detectAndCompute of GetRt on the GPU blocks (+ host retainBest ordering): keypoints, responses, angles,
    descriptors and their ORDER equal the cv2-pinned restatement."""
    for (w, h, nf, stream) in ((320, 240, 500, 0), (640, 480, 2000, 1)):
        img = oracle.gray(synth.SyntheticStream(stream, w, h).frame(2).bgr, 0)
        ref = _proto_features(proto, oracle, img, nf)
        kp, desc = capi.stage_cvorb_detect_and_compute(img, nf)
        assert len(kp) == len(ref) > 0.8 * nf
        assert [int(o) for o in kp["octave"]] == [r[0] for r in ref]
        for name, col in (("x", 1), ("y", 2), ("response", 3), ("angle", 4)):
            assert np.array_equal(kp[name].view(np.uint32), np.array([r[col] for r in ref], np.float32).view(np.uint32)), name
        assert np.array_equal(desc, np.stack([r[5] for r in ref]))
        sc = np.power(float(np.float32(1.2)), kp["octave"].astype(np.float64)).astype(np.float32)
        assert np.array_equal(kp["size"], np.float32(31.0) * sc) and (kp["class_id"] == -1).all()


def test_getrt_front_half_gives_the_oracle_point_sets(capi, proto, oracle, synth):
    """Features of both frames and the cross-check matcher on the GPU, then the reference's first-100 selection and
    back-projection: the object / image points handed to solvePnPRansac equal the all-CPU restatement bit for bit."""
    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    K = synth.intrinsics()
    g0, g5 = oracle.gray(f0.bgr, 0), oracle.gray(f5.bgr, 0)
    k1, d1 = capi.stage_cvorb_detect_and_compute(g0, 2000)
    k2, d2 = capi.stage_cvorb_detect_and_compute(g5, 2000)
    m_gpu = capi.stage_hamming_crosscheck(d1, d2)
    obj, pix = proto.back_project_matches(m_gpu, list(zip(k1["x"], k1["y"])), list(zip(k2["x"], k2["y"])), f0.depth_m, K, 100,
                                          oracle.sort_matches_order)
    feats = [_proto_features(proto, oracle, g, 2000) for g in (g0, g5)]
    dd = [np.stack([f[5] for f in fs]) for fs in feats]
    xy = [[(f[1], f[2]) for f in fs] for fs in feats]
    m_ref = proto.bf_match_hamming_crosscheck(dd[0], dd[1])
    obj_r, pix_r = proto.back_project_matches(m_ref, xy[0], xy[1], f0.depth_m, K, 100, oracle.sort_matches_order)
    assert m_gpu == m_ref and len(obj_r) >= 20
    assert np.array_equal(obj, obj_r) and np.array_equal(pix, pix_r)


def test_getrt_points_entry_point_and_final_pose(capi, proto, oracle, synth):
    """gd_getrt_points = GeoMaskMaker::GetRt up to solvePnPRansac: the point sets equal the all-CPU restatement, hence the
    pose cv2.solvePnPRansac + Rodrigues returns for them is the reference's (checked where cv2 is installed)."""
    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    K = synth.intrinsics()
    g0, g5 = oracle.gray(f0.bgr, 0), oracle.gray(f5.bgr, 0)
    obj, pix = capi.getrt_points(g0, g5, f0.depth_m, K, np.zeros(4, np.float32))
    seen = {}

    def solve(o, p, Kf):
        seen["obj"], seen["pix"] = o.copy(), p.copy()
        try:
            import cv2
        except ImportError:
            return np.eye(3), np.zeros(3)
        ok, rvec, tvec, _ = cv2.solvePnPRansac(o, p, Kf, np.zeros((4, 1), np.float32))
        return cv2.Rodrigues(rvec)[0], tvec

    ok, R_ref, T_ref = proto.get_rt(g0, g5, f0.depth_m, K, oracle, solve)
    assert ok and len(obj) >= 20
    assert np.array_equal(obj, seen["obj"]) and np.array_equal(pix, seen["pix"])
    R, T = solve(obj, pix, K)
    assert np.array_equal(np.asarray(R, np.float32), R_ref) and np.array_equal(np.asarray(T, np.float32).reshape(3), T_ref)


def test_getrt_points_distorted_camera(capi, proto, oracle, synth):
    """TUM1-like distortion: undistortPoints on the 100 matched points of the first image (GeoMaskMaker.cc:104-110), then the
    depth look-up at the undistorted pixel.  Reference = the GPU features / matches (pinned above) + cv2.undistortPoints."""
    import cv2

    s = synth.SyntheticStream(1)
    f0, f5 = s.frame(0), s.frame(5)
    K = np.array([[517.3, 0, 318.6], [0, 516.5, 255.3], [0, 0, 1]], np.float32)
    D = np.array([0.2624, -0.9531, -0.0054, 0.0026, 1.1633], np.float32)
    g0, g5 = oracle.gray(f0.bgr, 0), oracle.gray(f5.bgr, 0)
    k1, d1 = capi.stage_cvorb_detect_and_compute(g0, 2000)
    k2, d2 = capi.stage_cvorb_detect_and_compute(g5, 2000)
    m = capi.stage_hamming_crosscheck(d1, d2)

    def und(p):
        return cv2.undistortPoints(p.reshape(-1, 1, 2), K, D, None, K).reshape(-1, 2)

    obj_r, pix_r = proto.back_project_matches(m, list(zip(k1["x"], k1["y"])), list(zip(k2["x"], k2["y"])), f0.depth_m, K, 100,
                                              oracle.sort_matches_order, undistort=und)
    obj_p, _ = proto.back_project_matches(m, list(zip(k1["x"], k1["y"])), list(zip(k2["x"], k2["y"])), f0.depth_m, K, 100,
                                          oracle.sort_matches_order)
    for dist in (D, D[:4]):
        obj, pix = capi.getrt_points(g0, g5, f0.depth_m, K, dist)
        if len(dist) == 5:
            assert len(obj) >= 20 and np.array_equal(obj, obj_r) and np.array_equal(pix, pix_r)
            assert not np.array_equal(obj, obj_p[: len(obj)])  # the distortion does move the points
        else:
            und4 = lambda p: cv2.undistortPoints(p.reshape(-1, 1, 2), K, D[:4], None, K).reshape(-1, 2)  # noqa: E731
            o4, p4 = proto.back_project_matches(m, list(zip(k1["x"], k1["y"])), list(zip(k2["x"], k2["y"])), f0.depth_m, K, 100,
                                                oracle.sort_matches_order, undistort=und4)
            assert np.array_equal(obj, o4) and np.array_equal(pix, p4)


def test_getrt_resident_stage_in_the_frontend(capi, oracle, synth):
    """config.getrt: every step computes the new frame's cv::ORB features once (ring-slot cache), matches against the frame five
    steps back and returns the solvePnPRansac inputs.  They must equal the two-image entry point (which extracts both images)
    for every stream of a batch, over more steps than the ring is deep, with graph replay and plain launches; with a pose hook
    the mask equals the one computed from the same pose passed in."""
    import cv2

    K = synth.intrinsics()
    B, NF = 3, 14
    streams = [synth.SyntheticStream(s) for s in range(B)]
    frames = [[st.frame(f) for f in range(NF)] for st in streams]
    grays = [[oracle.gray(fr.bgr, 0) for fr in fs] for fs in frames]
    fe = capi.Frontend(K, 640, 480, batch=B, getrt=True)         # batch <= 8: CUDA graph replay after two ring cycles
    ref = capi.Frontend(K, 640, 480, batch=B)                     # no GetRt stage: masks from the pose passed in
    hooked = capi.Frontend(K, 640, 480, batch=B, getrt=True)
    poses = {}

    def solve(stream, obj, pix):
        ok, rvec, tvec, _ = cv2.solvePnPRansac(obj, pix, K, np.zeros((4, 1), np.float32))
        if not ok:
            return None
        R, T = cv2.Rodrigues(rvec)[0].astype(np.float32), tvec.astype(np.float32).reshape(3)
        poses[stream] = (R, T)
        return R, T

    hooked.set_pose_hook(solve)
    for f in range(NF):
        bgr = [frames[b][f].bgr for b in range(B)]
        dep = [frames[b][f].depth_m for b in range(B)]
        res = fe.step(bgr, dep)                                   # identity pose: only the GetRt points matter here
        pts = fe.fetch_getrt()
        poses.clear()
        res_h = [(m.copy(), k.copy(), d.copy()) for m, k, d in hooked.step(bgr, dep, use_hook=True)]
        for b in range(B):
            if f < 5:
                assert len(pts[b][0]) == 0
                assert res_h[b][0].min() == 1                      # warm-up: all ones
                continue
            obj, pix = capi.getrt_points(grays[b][f - 5], grays[b][f], frames[b][f - 5].depth_m, K, np.zeros(4, np.float32))
            assert len(obj) >= 20
            assert np.array_equal(pts[b][0], obj) and np.array_equal(pts[b][1], pix), (f, b)
        if f >= 5:
            assert sorted(poses) == list(range(B))
            R = np.stack([poses[b][0] for b in range(B)])
            T = np.stack([poses[b][1] for b in range(B)])
        else:
            R = T = None
        res_r = ref.step(bgr, dep, R, T)
        for b in range(B):
            assert np.array_equal(res_h[b][0], res_r[b][0]), (f, b)
            assert np.array_equal(res_h[b][1], res_r[b][1]) and np.array_equal(res_h[b][2], res_r[b][2])
            assert np.array_equal(res[b][1], res_r[b][1])          # ORB unaffected by the extra stage
        if f >= 5:
            assert (res_r[0][0] == 0).any()
    for x in (fe, ref, hooked):
        x.close()


def test_geomask_handle_getrt_points(capi, oracle, synth):
    """gd_geomask_getrt_points: what the drop-in GeoMaskMaker::GetRt calls before cv::solvePnPRansac."""
    K = synth.intrinsics()
    s = synth.SyntheticStream(2)
    fr = [s.frame(f) for f in range(8)]
    gm = capi.GeoMask(K, None, 5000.0, 640, 480, 0, batch=1)
    gm.enable_getrt()
    for f in range(8):
        gm.add_new_image([fr[f].bgr], [fr[f].depth_m])
        obj, pix = gm.getrt_points()[0]
        if f < 5:
            assert len(obj) == 0
            continue
        o2, p2 = capi.getrt_points(oracle.gray(fr[f - 5].bgr, 0), oracle.gray(fr[f].bgr, 0), fr[f - 5].depth_m, K)
        assert len(obj) >= 20 and np.array_equal(obj, o2) and np.array_equal(pix, p2), f
    gm.close()
