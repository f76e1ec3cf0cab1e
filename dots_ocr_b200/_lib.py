"""ctypes binding of libdots_ocr_b200.so (the C ABI declared in include/dots_ocr_b200.h).

There is deliberately no fallback: if the CUDA library is missing the product path raises.
Build it with ``python -m dots_ocr_b200.build`` (or ``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
# DOTS_B200_LIB selects an alternative build of the same ABI (kernel-tuning A/B runs); the default is the in-tree library.
LIB_PATH = os.environ.get("DOTS_B200_LIB") or os.path.join(HERE, "libdots_ocr_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(HERE), "include", "dots_ocr_b200.h")

_lib = None


class DotsLibraryError(RuntimeError):
    pass


def declared_symbols() -> List[str]:
    """Every function name the public header declares."""
    with open(HEADER_PATH) as f:
        text = f.read()
    return re.findall(r"DOTS_API\s+(?:const\s+char\*|int)\s+(dots_\w+)\s*\(", text)


def header_constants() -> Dict[str, int]:
    with open(HEADER_PATH) as f:
        text = f.read()
    return {k: int(v) for k, v in re.findall(r"#define\s+(DOTS_\w+)\s+(\d+)\b", text)}


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise DotsLibraryError(
            f"{LIB_PATH} not found: the sm_100a kernels are not built. Run `python -m dots_ocr_b200.build` "
            "(there is no CPU or PyTorch fallback for the hot path).")
    lib = ctypes.CDLL(LIB_PATH)
    lib.dots_last_error.restype = ctypes.c_char_p
    for name in declared_symbols():
        if not hasattr(lib, name):
            raise DotsLibraryError(f"{LIB_PATH} does not export {name} (stale build?)")
        if name != "dots_last_error":
            getattr(lib, name).restype = ctypes.c_int
    want = header_constants()["DOTS_ABI_VERSION"]
    got = lib.dots_abi_version()
    if got != want:
        raise DotsLibraryError(f"ABI mismatch: header {want}, library {got}")
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().dots_last_error().decode(errors="replace")
        raise RuntimeError(f"dots_ocr_b200 {what} failed (code {rc}): {msg}")
