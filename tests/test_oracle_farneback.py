"""Oracle Farneback restatement vs cv2.calcOpticalFlowFarneback goldens (CPU only)."""
import numpy as np

from conftest import flow_tol_violations


def test_synth_inputs_reproduce(synth, golden):
    g = golden("farneback.npz")
    s = synth.SyntheticStream(0)
    assert [synth.frame_crc(s.frame(0)), synth.frame_crc(s.frame(5))] == [int(v) for v in g["crc_640"]]


def test_farneback_640_vs_cv2(synth, oracle, golden):
    g = golden("farneback.npz")
    s = synth.SyntheticStream(0)
    g0 = oracle.gray(s.frame(0).bgr, 0)
    g5 = oracle.gray(s.frame(5).bgr, 0)
    flow = oracle.farneback(g0, g5)
    nviol, dmax = flow_tol_violations(flow[::8], g["flow_640_rows8"])
    assert nviol == 0, (nviol, dmax)
    assert abs(np.abs(flow).max() - float(g["flow_640_absmax"])) < 1e-3


def test_farneback_320_roll_vs_cv2(oracle, golden):
    g = golden("farneback.npz")
    flow = oracle.farneback(g["gray_320_a"], g["gray_320_b"])
    nviol, dmax = flow_tol_violations(flow, g["flow_320"])
    assert nviol == 0, (nviol, dmax)


def test_farneback_ragged_short_pyramid_vs_cv2(oracle, golden):
    """150x100: a level would fall under 32 px so OpenCV cuts the pyramid; odd level sizes."""
    g = golden("farneback.npz")
    flow = oracle.farneback(g["gray_150_a"], g["gray_150_b"])
    nviol, dmax = flow_tol_violations(flow, g["flow_150"])
    assert nviol == 0, (nviol, dmax)
