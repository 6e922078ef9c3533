"""Oracle vs the committed golden fixtures (CPU only): GeoMaskMaker arithmetic."""
import numpy as np

from conftest import flow_tol_violations  # noqa: F401


def test_gray_matches_cv2_golden(oracle, golden):
    g = golden("geomask_small.npz")
    assert np.array_equal(oracle.gray(g["bgr_ref"], 0), g["gray_ref"])
    assert np.array_equal(oracle.gray(g["bgr_cur"], 0), g["gray_cur"])


def test_depth_edge_bit_exact_vs_literal_transcription(oracle, golden):
    g = golden("geomask_small.npz")
    assert np.array_equal(oracle.depth_edge(g["depth_ref"], g["K"]), g["edge_ref"])
    assert np.array_equal(oracle.depth_edge(g["depth_cur"], g["K"]), g["edge_cur"])
    assert (g["edge_ref"] == 255).sum() > 100


def test_mahalanobis_vs_literal_cv2(oracle, golden):
    g = golden("geomask_small.npz")
    dist, written, src = oracle.mahalanobis(g["flow"], g["depth_ref"], g["depth_cur"], g["edge_ref"], g["edge_cur"],
                                            g["K"], g["R"], g["T"])
    nviol, dmax = flow_tol_violations(dist, g["dist"])
    assert nviol == 0, (nviol, dmax)
    # the restatement follows OpenCV's accumulation widths, so it is in fact bit-exact here
    assert np.array_equal(dist, g["dist"])
    assert np.array_equal(written == 1, g["dist"] > 0) or (written.sum() >= (g["dist"] > 0).sum())


def _inputs_640(oracle, synth, g):
    """Inputs of geomask_640.npz, regenerated from the seeded generator (checked by CRC)."""
    import zlib

    s = synth.SyntheticStream(0)
    f0, f5 = s.frame(0), s.frame(5)
    assert [synth.frame_crc(f0), synth.frame_crc(f5)] == [int(v) for v in g["crc"]]
    K = synth.intrinsics()
    flow = oracle.farneback(oracle.gray(f0.bgr), oracle.gray(f5.bgr))
    assert zlib.crc32(flow.tobytes()) == int(g["flow_crc"]), "oracle Farneback is not reproducible on this machine"
    e0, e5 = oracle.depth_edge(f0.depth_m, K), oracle.depth_edge(f5.depth_m, K)
    return flow, f0.depth_m, f5.depth_m, e0, e5


def test_mahalanobis_640x480_vs_literal_cv2(oracle, synth, golden):
    """The BASELINE size: the per-pixel loop of GeoMaskMaker.cc:208-272 over all 307 200 source pixels against the literal
    cv2 transcription (every 4th row of dist kept in the fixture, the whole mask, min/max)."""
    g = golden("geomask_640.npz")
    flow, d0, d5, e0, e5 = _inputs_640(oracle, synth, g)
    dist, _, _ = oracle.mahalanobis(flow, d0, d5, e0, e5, g["K"], g["R"], g["T"])
    assert np.array_equal(dist[::4], g["dist_rows4"])
    mask, _, mm = oracle.normalize_threshold(dist)
    assert np.array_equal(mask, np.unpackbits(g["mask"])[: 480 * 640].reshape(480, 640))
    assert np.array_equal(mm, g["minmax"])


def test_nan_value_is_skipped_by_minmax_like_cv(oracle):
    """A slightly negative quadratic form gives sqrt -> NaN.  cv::normalize's min/max scan ignores NaN (except at element 0,
    where the scan starts): only that pixel is affected (8-bit 0 -> mask 1), not the whole image."""
    import cv2

    dist, wr = nan_case(oracle)
    assert np.isnan(dist[wr > 0]).sum() >= 1 and not np.isnan(dist.flat[0])
    mask, d8, mm = oracle.normalize_threshold(dist)
    n = cv2.normalize(dist, None, 0.0, 255.0, cv2.NORM_MINMAX)
    with np.errstate(invalid="ignore"):
        ref8 = np.where(np.isnan(n), 0, np.clip(np.rint(n), 0, 255)).astype(np.uint8)
    assert np.array_equal(d8, ref8) and 0 < (mask == 0).sum()
    assert mm[0] == np.nanmin(dist) and mm[1] == np.nanmax(dist)
    d2 = dist.copy()
    d2.flat[0] = np.nan  # the scan starts here: everything becomes NaN -> 0 -> all ones
    assert oracle.normalize_threshold(d2)[0].min() == 1
    assert np.isnan(cv2.normalize(d2, None, 0.0, 255.0, cv2.NORM_MINMAX)).all()


def nan_case(oracle, want_inputs=False):
    """Seeded random pair (64x48, strong rotation) on which one Mahalanobis value is NaN (found by search)."""
    import importlib

    import cv2

    synth = importlib.import_module("gd-slam_b200.synth")
    K = synth.intrinsics(640, 480)
    rng = np.random.default_rng(1)
    h, w = 48, 64
    for trial in range(114):
        scale = 10 ** rng.uniform(-4, 0.5)
        dref = (rng.random((h, w)) * scale).astype(np.float32)
        dcur = (rng.random((h, w)) * scale).astype(np.float32)
        flow = (rng.normal(0, 1.0, (h, w, 2))).astype(np.float32)
        ang = rng.normal(0, 0.3, 3)
        R = cv2.Rodrigues(ang)[0].astype(np.float32)
        T = rng.normal(0, 0.05, 3).astype(np.float32)
    e = np.zeros((h, w), np.uint8)
    dist, wr, _ = oracle.mahalanobis(flow, dref, dcur, e, e, K, R, T)
    if want_inputs:
        return flow, dref, dcur, e, K, R, T, dist
    return dist, wr


def test_normalize_threshold_vs_cv2(oracle, golden):
    g = golden("geomask_small.npz")
    mask, d8, mm = oracle.normalize_threshold(g["dist"])
    assert np.array_equal(d8, g["d8"])
    assert np.array_equal(mask, g["mask"])
    assert set(np.unique(mask)) <= {0, 1}
    assert mm[0] == g["dist"].min() and mm[1] == g["dist"].max()


def test_normalize_degenerate_all_equal(oracle):
    z = np.zeros((8, 16), np.float32)
    mask, d8, _ = oracle.normalize_threshold(z)
    assert mask.min() == 1 and d8.max() == 0


def test_scatter_last_raster_writer_wins(oracle):
    """Two sources hitting one target: the later one in raster order stays (GeoMaskMaker.cc:269)."""
    w, h = 32, 16
    K = np.array([[50, 0, 16], [0, 50, 8], [0, 0, 1]], np.float32)
    flow = np.zeros((h, w, 2), np.float32)
    d0 = np.full((h, w), 1.0, np.float32)
    d1 = d0.copy()
    d0[5, 10] = 1.3  # the earlier source is different so the two candidate values differ
    flow[5, 10] = (1.0, 0.0)   # (10,5) -> (11,5)
    flow[5, 11] = (0.4, 0.0)   # (11,5) -> (11,5) : later in raster order
    e = np.zeros((h, w), np.uint8)
    R = np.eye(3, dtype=np.float32)
    T = np.array([0.01, 0, 0], np.float32)
    dist, written, src = oracle.mahalanobis(flow, d0, d1, e, e, K, R, T)
    assert src[5, 11] == 5 * w + 11
    assert written[5, 10] == 0  # its only source moved away
    f2 = flow.copy()
    f2[5, 11] = (-50.0, 0.0)  # later source leaves the image -> the earlier one is the last writer
    dist2, _, src2 = oracle.mahalanobis(f2, d0, d1, e, e, K, R, T)
    assert src2[5, 11] == 5 * w + 10
    assert dist2[5, 11] != dist[5, 11]


def test_erode_filter_vs_cv2_golden(oracle, golden):
    """Row (f)-2: erode(mask, 31x31 ellipse) + keypoint filter of Frame.cc:258-282 against cv2.erode."""
    g = golden("erode.npz")
    mask = np.unpackbits(g["mask"])[: 480 * 640].reshape(480, 640)
    er = np.unpackbits(g["eroded"])[: 480 * 640].reshape(480, 640)
    assert np.array_equal(oracle.erode31(mask), er)
    keep = oracle.erode_filter(mask, g["kp"])
    assert np.array_equal(keep, g["keep"])
    assert 0 < keep.sum() < len(keep)


def test_stereo_grid_vs_literal_transcription(oracle, synth, golden):
    """Row (f)-3 pinned: UndistortKeyPoints, ComputeImageBounds, ComputeStereoFromRGBD, AssignFeaturesToGrid / PosInGrid
    (src/Frame.cc:576-636, 815-837, 402-417, 553-565) against the literal numpy + cv2.undistortPoints transcription, for an
    undistorted (TUM3) and a distorted (TUM1) camera."""
    g = golden("stereo_grid.npz")
    fr = synth.SyntheticStream(0).frame(5)
    assert synth.frame_crc(fr) == int(g["crc"])
    for cam in ("tum3", "tum1"):
        d, ur, cs, ci, un, bounds = oracle.stereo_grid(fr.depth_m, g["kp"], 40.0, g[cam + "_K"], g[cam + "_D"], want_undistorted=True)
        assert np.array_equal(un, g[cam + "_un"]), cam
        assert np.array_equal(bounds, g[cam + "_bounds"]), cam
        assert np.array_equal(d, g[cam + "_depth"]) and np.array_equal(ur, g[cam + "_uright"]), cam
        assert np.array_equal(cs, g[cam + "_cell_start"]) and np.array_equal(ci, g[cam + "_cell_items"]), cam
    assert not np.array_equal(g["tum1_un"], g["tum3_un"]) and not np.array_equal(g["tum1_cell_start"], g["tum3_cell_start"])
