"""ctypes binding of libwdbx_b200.so (C ABI declared in include/wdbx_b200.h).

There is deliberately no fallback: if the shared library is missing or a symbol is absent the
import of the engine fails loudly -- the GPU path is the only path.
"""
from __future__ import annotations

import ctypes as C
import threading
from pathlib import Path

import os

# WDBX_B200_LIB: load another build of the same ABI (the -DWDBX_DEBUG_BOUNDS library, `python __graft_entry__.py debug`)
LIB_PATH = Path(os.environ.get("WDBX_B200_LIB") or Path(__file__).resolve().parent / "libwdbx_b200.so")

MAX_SEGMENTS = 64
MAX_K = 1024
ALL_SEGMENTS = -1
EACH_SEGMENT = -2
OK, ERR_ARG, ERR_CUDA, ERR_OOM, ERR_LIMIT = 0, -1, -2, -3, -4
F32, BF16 = 0, 1
COSINE, IP, L2 = 0, 1, 2
METRICS = {"cosine": COSINE, "ip": IP, "inner_product": IP, "dot": IP, "l2": L2, "euclidean": L2}
DTYPES = {"fp32": F32, "float32": F32, "f32": F32, "bf16": BF16, "bfloat16": BF16}


class Stats(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32), ("dim", C.c_int32), ("dim_padded", C.c_int32),
        ("dtype", C.c_int32), ("num_segments", C.c_int32), ("sm_count", C.c_int32), ("last_kernel", C.c_int32),
        ("rows_total", C.c_int64), ("rows_live", C.c_int64), ("capacity_rows", C.c_int64),
        ("bytes_resident", C.c_int64), ("kernel_launches", C.c_int64), ("searches", C.c_int64),
        ("last_search_ms", C.c_double),
        ("seg_rows", C.c_int64 * MAX_SEGMENTS), ("seg_live", C.c_int64 * MAX_SEGMENTS),
        ("last_kernel_ms", C.c_double), ("last_candidates", C.c_int64),
    ]


_P = C.c_void_p
# name -> (restype, argtypes); every symbol include/wdbx_b200.h declares
SIGNATURES = {
    "wdbx_b200_version": (C.c_int, []),
    "wdbx_b200_last_error": (C.c_char_p, []),
    "wdbx_b200_device_count": (C.c_int, []),
    "wdbx_b200_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_P)]),
    "wdbx_b200_destroy": (None, [_P]),
    "wdbx_b200_reserve": (C.c_int, [_P, C.c_int, C.c_int64]),
    "wdbx_b200_append": (C.c_int, [_P, C.c_int, _P, C.c_int64, C.c_int, _P, C.POINTER(C.c_int64)]),
    "wdbx_b200_overwrite": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "wdbx_b200_tombstone": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int]),
    "wdbx_b200_clear": (C.c_int, [_P, C.c_int]),
    "wdbx_b200_read_row": (C.c_int, [_P, C.c_int, C.c_int64, _P]),
    "wdbx_b200_read_rows": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P]),
    "wdbx_b200_search": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "wdbx_b200_search_host": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "wdbx_b200_search_filtered_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "wdbx_b200_merge": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "wdbx_b200_exchange_init": (C.c_int, [_P, C.c_int, C.c_int, _P]),
    "wdbx_b200_exchange_attach": (C.c_int, [_P, C.c_int, _P]),
    "wdbx_b200_search_exchange": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "wdbx_b200_set_tuning": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "wdbx_b200_set_kernel_timing": (C.c_int, [_P, C.c_int]),
    "wdbx_b200_set_option": (C.c_int, [_P, C.c_char_p, C.c_longlong]),
    "wdbx_b200_search_exchange_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "wdbx_b200_get_stats": (C.c_int, [_P, C.POINTER(Stats)]),
    "wdbx_b200_search_exchange_filtered_host": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P]),
    "wdbx_b200_group_create": (C.c_int, [_P, C.c_int, C.POINTER(_P)]),
    "wdbx_b200_group_destroy": (None, [_P]),
    "wdbx_b200_group_search_host": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_float, _P, _P, _P, _P, _P]),
    "wdbx_b200_group_search": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
}

_lib = None
_lock = threading.Lock()


class B200Error(RuntimeError):
    """A C-ABI call returned a negative code."""

    def __init__(self, code: int, message: str):
        super().__init__(f"libwdbx_b200 error {code}: {message}")
        self.code = code


def load_library() -> C.CDLL:
    """dlopen the in-tree library and bind every declared symbol (raises if anything is missing)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
                "wdbx_b200 has no CPU fallback.")
        lib = C.CDLL(str(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.wdbx_b200_version() != 1:
            raise ImportError(f"ABI mismatch: library reports version {lib.wdbx_b200_version()}, binding expects 1")
        _lib = lib
        return lib


def check(code: int) -> int:
    if code < 0:
        msg = load_library().wdbx_b200_last_error()
        raise B200Error(code, msg.decode("utf-8", "replace") if msg else "")
    return code
