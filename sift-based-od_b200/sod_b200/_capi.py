"""ctypes binding of libsod_b200.so (declared in include/sod.h).

There is no fallback: if the library is missing or cannot be loaded, importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

# SOD_B200_LIB selects another build of the same library (kernel experiments); never a fallback.
_LIB_PATH = Path(os.environ.get("SOD_B200_LIB") or Path(__file__).resolve().parent / "libsod_b200.so")

SOD_OK = 0
DESC_DIM = 128
TILE_ROWS = 128


class SodError(RuntimeError):
    pass


def _load() -> C.CDLL:
    if not _LIB_PATH.exists():
        raise ImportError(
            f"{_LIB_PATH} not built: run `python __graft_entry__.py` (or sod_b200.build.build()); "
            "the CUDA library is the only implementation of the hot path")
    return C.CDLL(str(_LIB_PATH))


lib = _load()

_p = C.c_void_p
_i64 = C.c_int64
_i32 = C.c_int32

_PROTOS = {
    "sod_version": (C.c_int, []),
    "sod_last_error": (C.c_char_p, []),
    "sod_device_sm_count": (C.c_int, []),
    "sod_cq_ints": (_i64, [_i64]),
    "sod_pack_u8_from_f32": (C.c_int, [_p, _i64, _p, _p, _p]),
    "sod_db_prepare": (C.c_int, [_p, _i64, _p, _p]),
    "sod_query_prepare": (C.c_int, [_p, _i64, _p, _p]),
    "sod_match_workspace_bytes": (C.c_size_t, [_i64, _i64]),
    "sod_match_top2": (C.c_int, [_p, _p, _i64, _p, _p, _i64, _i32, _p, _p, _p, C.c_size_t, _p]),
    "sod_top2_merge": (C.c_int, [_p, _p, _i32, _i64, _p, _p, _p, _p, C.c_double, _p]),
}


def _bind() -> None:
    for name, (res, args) in _PROTOS.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args


_bind()

EXPORTED = tuple(_PROTOS)


def check(rc: int, what: str) -> None:
    if rc != SOD_OK:
        msg = lib.sod_last_error().decode("utf-8", "replace")
        raise SodError(f"{what} failed ({rc}): {msg}")
