"""CPU: gd::stdalgo (the libstdc++ algorithm restatements one device thread runs for cv::ORB's keypoint order and GetRt's
match order) against the real std::nth_element / std::partition / std::sort on randomised and adversarial inputs."""
import ctypes
import os
import subprocess
import tempfile

from conftest import ROOT


def test_stdalgo_equals_libstdcxx():
    src = os.path.join(ROOT, "tests", "native", "stdalgo_check.cpp")
    with tempfile.TemporaryDirectory() as d:
        so = os.path.join(d, "libstdalgo_check.so")
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "gd-slam_b200", "csrc"), src, "-o", so],
                       check=True)
        L = ctypes.CDLL(so)
        assert L.gd_stdalgo_selfcheck(1, 400) == 0
        assert L.gd_stdalgo_selfcheck(7, 400) == 0
        # the adversarial input costs several times n lg n comparisons: it does reach introsort's heap fallback
        assert L.gd_stdalgo_killer_depth(4096) > 25  # ~16 for random input
