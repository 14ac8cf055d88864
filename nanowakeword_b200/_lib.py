"""ctypes binding of libnwwb200.so (the C ABI in include/nww_b200.h).

The CUDA library is the product: if it is missing or fails to load this module raises —
there is no Python/NumPy/torch fallback for the compute path.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH

NWW_OK, NWW_EINVAL, NWW_ECUDA, NWW_EUNSUPPORTED = 0, -1, -2, -3


class NwwSpec(C.Structure):
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("arch", C.c_int32),
        ("activation", C.c_int32),
        ("geometry", C.c_int32),
        ("n_fft", C.c_int32), ("win_length", C.c_int32), ("hop_length", C.c_int32),
        ("n_mels", C.c_int32), ("center", C.c_int32), ("clip_samples", C.c_int32),
        ("frontend_precision", C.c_int32),
        ("chunk_windows", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


class NwwInfo(C.Structure):
    _fields_ = [
        ("device", C.c_int32), ("sm_count", C.c_int32),
        ("n_mels", C.c_int32), ("n_frames", C.c_int32), ("clip_samples", C.c_int32),
        ("feature_dim", C.c_int32), ("embedding_dim", C.c_int32),
        ("chunk_windows", C.c_int32),
        ("kernel_launches", C.c_int64), ("windows_scored", C.c_int64),
    ]


class NwwProfile(C.Structure):
    _fields_ = [
        ("stage_a_ms", C.c_double), ("stage_b_ms", C.c_double),
        ("stage_a_spans", C.c_int64), ("stage_b_spans", C.c_int64),
        ("stage_a_windows", C.c_int64), ("stage_b_windows", C.c_int64),
    ]


# name -> (restype, argtypes); kept in one table so tests can check every symbol the header declares
_P = C.c_void_p
SIGNATURES = {
    "nww_create": (C.c_int, [C.POINTER(NwwSpec), _P, C.c_size_t, C.c_int, C.POINTER(_P)]),
    "nww_destroy": (None, [_P]),
    "nww_last_error": (C.c_char_p, []),
    "nww_get_info": (C.c_int, [_P, C.POINTER(NwwInfo)]),
    "nww_run_windows": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "nww_run_windows_f32": (C.c_int, [_P, _P, C.c_int64, _P, _P, _P, _P, _P]),
    "nww_run_windows_host": (C.c_int, [_P, _P, C.c_int64, _P]),
    "nww_logmel": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, _P]),
    "nww_stream_open": (C.c_int, [_P, C.c_int64]),
    "nww_stream_push": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "nww_stream_push_host": (C.c_int, [_P, _P, C.c_int32, _P]),
    "nww_stream_push_select": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int64, _P, _P]),
    "nww_stream_push_select_host": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int64, _P]),
    "nww_stream_reset": (C.c_int, [_P, _P, C.c_int64]),
    "nww_stream_close": (C.c_int, [_P]),
    "nww_set_profiling": (C.c_int, [_P, C.c_int]),
    "nww_get_profile": (C.c_int, [_P, C.POINTER(NwwProfile)]),
    "nww_synchronize": (C.c_int, [_P]),
    "nww_microbench": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_double)]),
}

_lib = None


def load_library(path: str | None = None) -> C.CDLL:
    """Load libnwwb200.so, binding every entry point.  Raises if it is absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if path is None and not os.path.exists(p):
        # a fresh checkout: the library is a build product (git-ignored).  Build it if a compiler is here;
        # there is still no fallback — without nvcc this raises.
        try:
            from .build import build
            build(force=True, verbose=False)
        except Exception as ex:
            raise RuntimeError(
                f"CUDA engine library not found at {p} and building it failed ({ex}). "
                "nanowakeword_b200 has no CPU fallback.") from ex
    if not os.path.exists(p):
        raise RuntimeError(
            f"CUDA engine library not found at {p}. Build it with `python -m nanowakeword_b200.build` "
            "(nvcc, sm_100a). nanowakeword_b200 has no CPU fallback.")
    lib = C.CDLL(p)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


def last_error(lib: C.CDLL) -> str:
    msg = lib.nww_last_error()
    return msg.decode(errors="replace") if msg else ""


def check(lib: C.CDLL, rc: int, what: str) -> None:
    """Map C return codes onto the exception classes the reference raises
    (ValueError for bad input, RuntimeError for runtime failures)."""
    if rc == NWW_OK:
        return
    msg = f"{what}: {last_error(lib)}"
    if rc == NWW_EINVAL:
        raise ValueError(msg)
    if rc == NWW_EUNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
