"""ctypes binding of libroomslam_b200.so (the C ABI declared in include/roomslam_b200.h).

There is no CPU fallback: a missing library or a failing call raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List, Tuple

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libroomslam_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "roomslam_b200.h")

_C2CT = {
    "int": ctypes.c_int, "float": ctypes.c_float, "int64_t": ctypes.c_int64, "double": ctypes.c_double,
    "unsigned long long": ctypes.c_ulonglong, "uint64_t": ctypes.c_uint64, "int32_t": ctypes.c_int32,
}
_lib = None


class RoomSlamError(RuntimeError):
    pass


def parse_header(path: str = HEADER_PATH) -> Dict[str, Tuple[str, List[str]]]:
    """{symbol: (return type, [argument types])} for every prototype in the public header."""
    text = open(path).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"(const char\*|int64_t|int|void)\s+(rs_\w+)\s*\(([^)]*)\)\s*;", text):
        ret, name, args = m.group(1), m.group(2), m.group(3).strip()
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = " ".join(a.split())
                if "*" in a:
                    types.append("ptr")
                else:
                    types.append(" ".join(a.split(" ")[:-1]).replace("const ", "").strip())
        protos[name] = (ret, types)
    return protos


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RoomSlamError(
            f"{LIB_PATH} is missing: build it with `python -m roomslam_b200.build` "
            "(roomslam_b200 runs only on its sm_100a CUDA kernels; there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (ret, types) in parse_header().items():
        fn = getattr(lib, name)  # AttributeError if the header declares something the library lacks
        fn.restype = {"const char*": ctypes.c_char_p, "void": None, "int64_t": ctypes.c_int64}.get(ret, ctypes.c_int)
        fn.argtypes = [ctypes.c_void_p if t == "ptr" else _C2CT[t] for t in types]
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().rs_last_error()
        raise RoomSlamError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")


def call(name: str, *args) -> None:
    check(getattr(load(), name)(*args), name)
