"""GPU (-m gpu): the C++ drop-in classes (gd-slam_b200/host: GeoMaskMaker, ORB_SLAM2::ORBextractor) driven like
Tracking::GrabImageRGBD_GD, compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def test_cpp_shim_sequence(tmp_path, synth, oracle):
    exe = os.path.join(ROOT, "gd-slam_b200", "lib", "shim_demo")
    assert os.path.exists(exe), "run __graft_entry__.build() first"
    w, h, nf = 640, 480, 7
    s = synth.SyntheticStream(4, roll_deg_per_frame=0.04)
    frames = [s.frame(f) for f in range(nf)]
    K = synth.intrinsics()
    inp, outp = tmp_path / "seq.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<3i", w, h, nf))
        for i, fr in enumerate(frames):
            R, T = s.pair_pose(i - 5, i) if i >= 5 else (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
            f.write(fr.bgr.tobytes())
            f.write(oracle.gray(fr.bgr, 1).tobytes())
            f.write(fr.depth_m.tobytes())
            f.write(np.ascontiguousarray(R, np.float32).tobytes())
            f.write(np.ascontiguousarray(T, np.float32).tobytes())
    r = subprocess.run([exe, str(inp), str(outp)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
    assert "0checking" in r.stdout  # the reference prints the frame counter (GeoMaskMaker.cc:169)
    buf = open(outp, "rb").read()
    off = 0
    for i, fr in enumerate(frames):
        (n,) = struct.unpack_from("<i", buf, off)
        off += 4
        rec = np.frombuffer(buf, dtype=np.dtype([("f", "<f4", 5), ("oct", "<i4"), ("d", "u1", 32)]), count=n, offset=off)
        off += n * 56
        mask = np.frombuffer(buf, np.uint8, w * h, off).reshape(h, w)
        off += w * h
        rkp, rdesc, _ = oracle.orb_extract(oracle.gray(fr.bgr, 1))
        assert n == len(rkp)
        assert np.array_equal(rec["f"][:, 0], rkp["x"]) and np.array_equal(rec["f"][:, 1], rkp["y"])
        assert np.array_equal(rec["f"][:, 3], rkp["angle"]) and np.array_equal(rec["oct"], rkp["octave"])
        assert np.array_equal(rec["d"], rdesc)
        if i < 5:
            assert mask.min() == 1 and mask.max() == 1
        else:
            R, T = s.pair_pose(i - 5, i)
            mo = oracle.geomask_pair(frames[i - 5].bgr, fr.bgr, frames[i - 5].depth_m, fr.depth_m, K, R, T)
            assert (mask == mo).mean() >= 0.999
    assert off == len(buf)


def test_cpp_shim_getrt_path(tmp_path, synth, oracle):
    """No pose provider: GeoMaskMaker::GetRt() itself (GPU points + solver).  The demo's stand-in solver logs the points it
    receives: they must equal gd_getrt_points for the buffered pair; the masks follow the pose it answers with."""
    from conftest import load_pkg

    capi = load_pkg("capi")
    exe = os.path.join(ROOT, "gd-slam_b200", "lib", "shim_demo")
    w, h, nf = 640, 480, 8
    s = synth.SyntheticStream(5)
    frames = [s.frame(f) for f in range(nf)]
    K = synth.intrinsics()
    inp, outp = tmp_path / "seq.bin", tmp_path / "out.bin"
    with open(inp, "wb") as f:
        f.write(struct.pack("<3i", w, h, nf))
        for i, fr in enumerate(frames):
            R, T = s.pair_pose(i - 5, i) if i >= 5 else (np.eye(3, dtype=np.float32), np.zeros(3, np.float32))
            f.write(fr.bgr.tobytes())
            f.write(oracle.gray(fr.bgr, 1).tobytes())
            f.write(fr.depth_m.tobytes())
            f.write(np.ascontiguousarray(R, np.float32).tobytes())
            f.write(np.ascontiguousarray(T, np.float32).tobytes())
    r = subprocess.run([exe, str(inp), str(outp), "getrt"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.returncode, r.stderr[-500:])
    buf = open(outp, "rb").read()
    off = 0
    for i, fr in enumerate(frames):
        (n,) = struct.unpack_from("<i", buf, off)
        off += 4 + n * 56
        mask = np.frombuffer(buf, np.uint8, w * h, off).reshape(h, w)
        off += w * h
        npts = int(np.frombuffer(buf, np.float32, 1, off)[0])
        off += 4
        obj = np.frombuffer(buf, np.float32, npts * 3, off).reshape(-1, 3)
        off += npts * 12
        pix = np.frombuffer(buf, np.float32, npts * 2, off).reshape(-1, 2)
        off += npts * 8
        if i < 5:
            assert npts == 0 and mask.min() == 1
            continue
        o2, p2 = capi.getrt_points(oracle.gray(frames[i - 5].bgr, 0), oracle.gray(fr.bgr, 0), frames[i - 5].depth_m, K)
        assert npts >= 20 and np.array_equal(obj, o2) and np.array_equal(pix, p2), i
        R, T = s.pair_pose(i - 5, i)
        mo = oracle.geomask_pair(frames[i - 5].bgr, fr.bgr, frames[i - 5].depth_m, fr.depth_m, K, R, T)
        assert (mask == mo).mean() >= 0.995  # the pose went through a rotation-vector round trip in the stand-in solver
    assert off == len(buf)
