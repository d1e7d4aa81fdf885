"""ctypes loader for oracle/heatmap_ref.c (TEST INFRASTRUCTURE ONLY; see that file's header)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libheatmap_ref.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "heatmap_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/libheatmap_ref.so"], stdout=subprocess.DEVNULL)
    return _SO


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        _lib.rs_ref_heatmap_bin.restype = ctypes.c_int64
        _lib.rs_ref_heatmap_bin.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_float,
                                            ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_float, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return _lib


def bin_points(points: np.ndarray, x_min, y_min, res, gx: int, gy: int, thr2, n_threads: int = 0):
    """Same contract as oracle.baseline_ref.bin_points, computed by the C restatement."""
    p = np.ascontiguousarray(points, dtype=np.float32)
    assert p.ndim == 3 and p.shape[-1] == 2
    occ = np.empty((gy, gx), dtype=np.int32)
    stat = np.empty((gy, gx), dtype=np.int32)
    nd = _load().rs_ref_heatmap_bin(p.ctypes.data, p.shape[0], p.shape[1], np.float32(x_min), np.float32(y_min),
                                    np.float32(res), gx, gy, np.float32(thr2), occ.ctypes.data, stat.ctypes.data,
                                    int(n_threads))
    return occ, stat, int(nd)
