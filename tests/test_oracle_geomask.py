"""Oracle vs the committed golden fixtures (CPU only): GeoMaskMaker arithmetic."""
import numpy as np

from conftest import flow_tol_violations


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
