"""CPU: the C-ABI library loads and exports every symbol include/wdbx_b200.h declares."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "wdbx_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(wdbx_b200_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_exported(built_lib):
    from wdbx_b200 import _lib

    names = _declared()
    assert len(names) >= 15
    raw = C.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/wdbx_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == names, "ctypes binding and header disagree"
    assert built_lib.wdbx_b200_version() == 1


def test_stats_struct_layout():
    from wdbx_b200 import _lib

    assert C.sizeof(_lib.Stats) == 8 * 4 + 6 * 8 + 8 + 2 * 64 * 8 + 8 + 8   # ... + last_kernel_ms + last_candidates


def test_argument_errors_need_no_device(built_lib):
    from wdbx_b200 import _lib

    h = C.c_void_p()
    assert built_lib.wdbx_b200_create(0, 0, 0, 1, C.byref(h)) == _lib.ERR_ARG
    assert b"dim" in built_lib.wdbx_b200_last_error()
    assert built_lib.wdbx_b200_create(0, 8, 7, 1, C.byref(h)) == _lib.ERR_ARG
    assert built_lib.wdbx_b200_create(0, 8, 0, 65, C.byref(h)) == _lib.ERR_LIMIT
    assert built_lib.wdbx_b200_search(None, -1, None, 1, 1, 0, None, None, None, None, None) == _lib.ERR_ARG
    assert built_lib.wdbx_b200_get_stats(None, None) == _lib.ERR_ARG


def test_fails_loudly_without_gpu(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import wdbx_b200

    with pytest.raises(wdbx_b200.B200Error, match="no CPU fallback"):
        wdbx_b200.Engine(0, 8)
    with pytest.raises(wdbx_b200.B200Error):
        wdbx_b200.WDBX(vector_dimension=4, data_dir=str(ROOT / "build" / "_t"))
