"""The C-ABI library loads and exports every symbol include/gdslam_cuda.h declares (no compute calls: CPU only)."""
import os
import re

from conftest import ROOT, load_pkg


def _declared():
    txt = open(os.path.join(ROOT, "include", "gdslam_cuda.h")).read()
    return sorted(set(re.findall(r"GD_API\s+[\w\s\*]+?\b(gd_\w+)\s*\(", txt)))


def test_header_symbols_all_exported_and_bound():
    capi = load_pkg("capi")
    names = _declared()
    assert len(names) >= 35
    L = capi.lib()
    missing = [n for n in names if not hasattr(L, n)]
    assert not missing, missing
    unbound = [n for n in names if n not in capi.SYMBOLS]
    assert not unbound, unbound
    assert capi.missing_symbols() == []
    assert L.gd_abi_version() == 2


def test_no_cpu_fallback_without_device():
    """Without a CUDA device every compute entry point must fail loudly (GD_ENODEVICE), never fall back."""
    import numpy as np
    import pytest

    capi = load_pkg("capi")
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GdError) as e:
        capi.stage_gray(np.zeros((8, 8, 3), np.uint8))
    assert e.value.code == capi.GD_ENODEVICE
    with pytest.raises(capi.GdError):
        capi.GeoMask(np.eye(3, dtype=np.float32), None, 5000.0, 64, 64, 0, 1)
    with pytest.raises(capi.GdError):
        capi.Orb(500, 1.2, 4, 20, 7, 320, 240, 0, 1)


def test_product_never_imports_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may touch oracle/."""
    pkg = os.path.join(ROOT, "gd-slam_b200")
    offenders = []
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", ".cc")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                if re.search(r"(from|import)\s+oracle|#include\s+[\"<].*oracle|pyoracle|liboracle", txt):
                    offenders.append(os.path.join(dp, f))
    assert not offenders, offenders
