"""CPU: the C-ABI library builds, loads, and exports every symbol include/roomslam_b200.h declares.
No compute call is made (there is no GPU here); calls must fail loudly, never fall back."""
import ctypes
import subprocess

import pytest

from roomslam_b200 import _lib


def test_library_exports_every_declared_symbol(built_lib):
    protos = _lib.parse_header()
    assert {"rs_last_error", "rs_abi_version", "rs_heatmap_bin", "rs_heatmap_bin_host"} <= set(protos)
    lib = ctypes.CDLL(built_lib)
    for name in protos:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    exported = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True).stdout
    have = {ln.split()[-1] for ln in exported.splitlines() if " T rs_" in ln}
    assert have == set(protos), f"header and library disagree: {have ^ set(protos)}"


def test_abi_version_and_binding(built_lib):
    lib = _lib.load()
    assert lib.rs_abi_version() >= 1
    for name, (_, types) in _lib.parse_header().items():
        assert len(getattr(lib, name).argtypes or []) == len(types)


def test_sass_is_blackwell_native(built_lib):
    sass = subprocess.run(["cuobjdump", "-sass", built_lib], capture_output=True, text=True).stdout
    assert "sm_100a" in sass
    assert "UTMALDG" in sass, "the binning kernel must be fed by TMA"
    assert "ATOMS" in sass, "the binning kernel must use shared-memory privatised histograms"


def test_no_cpu_fallback_without_gpu(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import roomslam_b200
    b = roomslam_b200.OccupancyHeatmapBaseline()
    with pytest.raises(_lib.RoomSlamError):
        b.heatmap(torch.zeros(2, 4, 2))
    assert _lib.load().rs_device_ok() == 0


def test_every_symbol_is_documented_in_integration_md():
    """INTEGRATION.md maps every exported entry point to the upstream interface it replaces."""
    import os
    from roomslam_b200 import _lib
    text = open(os.path.join(os.path.dirname(os.path.dirname(__file__)), "INTEGRATION.md")).read()
    missing = [name for name in _lib.parse_header() if name not in text]
    assert not missing, missing
