"""GPU parity (-m gpu) of the configuration bench.py times: one handle x 64 streams stepped from device-resident slots
(`value` leg) and 4 handles x 16 streams stepped with host buffers from 4 threads (`e2e` leg), plain launches + programmatic
dependent launch + forked streams, long enough for the 6-deep ring and the 12 staged slots to wrap.

Every stream of both configurations must equal a batch-1 handle (CUDA-graph replay path) bit for bit — mask, keypoints,
descriptors, and the resolved Mahalanobis image — and four of the 64 streams are checked against the CPU oracle."""
import threading

import numpy as np
import pytest

from conftest import load_pkg

pytestmark = pytest.mark.gpu

B, D, S, STEPS, NH = 64, 4, 12, 15, 4


@pytest.fixture(scope="module")
def capi():
    c = load_pkg("capi")
    c.lib()
    assert c.device_count() >= 1
    return c


def _same(a, b, what):
    assert np.array_equal(a[0], b[0]), (what, "mask")
    assert len(a[1]) == len(b[1]), (what, "n_kp")
    for f in a[1].dtype.names:
        assert np.array_equal(a[1][f], b[1][f]), (what, f)
    assert np.array_equal(a[2], b[2]), (what, "desc")


def test_bench_configuration_equals_single_stream_handles_and_oracle(capi, oracle, synth):
    import bench  # the benchmark's own data layout: D distinct seeded streams replicated over the batch with a frame offset

    bgr, dep, poses = bench.make_data(D, S, seed0=0)
    K = synth.intrinsics()
    hb = np.empty((S, B, 480, 640, 3), np.uint8)
    hd = np.empty((S, B, 480, 640), np.float32)
    Rs = np.zeros((S, B, 3, 3), np.float32)
    Ts = np.zeros((S, B, 3), np.float32)
    for s in range(S):
        for b in range(B):
            src, off = b % D, (b // D) % S
            f = (s + off) % S
            hb[s, b], hd[s, b] = bgr[src, f], dep[src, f]
            Rs[s, b], Ts[s, b] = poses[(src, f)]

    big = capi.Frontend(K, 640, 480, batch=B, staged_slots=S)           # `value` leg
    for s in range(S):
        big.stage(s, hb[s], hd[s])
    Bh = B // NH
    quads = [capi.Frontend(K, 640, 480, batch=Bh) for _ in range(NH)]   # `e2e` leg: 4 handles, one host thread each
    ones = [capi.Frontend(K, 640, 480, batch=1) for _ in range(B)]      # the reference: one stream per handle (graph replay)

    def copy_results(res):
        return [(m.copy(), k.copy(), d.copy()) for m, k, d in res]

    for k in range(STEPS):
        s = k % S
        big.step_staged(s, Rs[s], Ts[s])
        out_q = [None] * NH

        def run(i):
            sl = slice(i * Bh, (i + 1) * Bh)
            out_q[i] = copy_results(quads[i].step(hb[s, sl], hd[s, sl], Rs[s, sl], Ts[s, sl]))

        th = [threading.Thread(target=run, args=(i,)) for i in range(NH)]
        for t in th:
            t.start()
        r_big = copy_results(big.fetch())
        for t in th:
            t.join()
        for b in range(B):
            r1 = ones[b].step(hb[s, b:b + 1], hd[s, b:b + 1], Rs[s, b:b + 1], Ts[s, b:b + 1])[0]
            _same(r_big[b], r1, (k, b, "batch 64 staged"))
            _same(out_q[b // Bh][b % Bh], r1, (k, b, "4 x 16 host"))
            if k >= STEPS - 2:
                d1 = ones[b].debug(capi.DBG_DIST)
                assert np.array_equal(big.debug(capi.DBG_DIST, b), d1, equal_nan=True), (k, b)
                assert np.array_equal(quads[b // Bh].debug(capi.DBG_DIST, b % Bh), d1, equal_nan=True), (k, b)
    assert big.launch_count() > 0

    # four of the 64 streams of the last step against the oracle (reference pair = the frame five steps earlier)
    s, s_ref = (STEPS - 1) % S, (STEPS - 1 - 5) % S
    for b in (0, 21, 42, 63):
        mask, kp, desc = r_big[b]
        rkp, rdesc, _ = oracle.orb_extract(oracle.gray(hb[s, b], 1))
        _same((mask, kp, desc), (mask, rkp, rdesc), (b, "ORB vs oracle"))
        mo = oracle.geomask_pair(hb[s_ref, b], hb[s, b], hd[s_ref, b], hd[s, b], K, Rs[s, b], Ts[s, b])
        assert (mask == mo).mean() >= 0.999, (b, float((mask == mo).mean()))
        assert (mask == 0).any()  # a real mask, not the all-ones warm-up one
    for f in [big] + quads + ones:
        f.close()
