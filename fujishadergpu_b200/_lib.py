"""ctypes binding of libfsg_b200.so (the C ABI declared in include/fsg_b200.h).

The product path has NO fallback: if the shared library is missing or a call fails, the caller
gets an exception (ValueError for argument errors, as the reference raises; RuntimeError otherwise).
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from .build import LIB_PATH

FSG_OUT_F32, FSG_OUT_I16, FSG_OUT_U8 = 0, 1, 2
_KIND = {"float32": FSG_OUT_F32, "int16": FSG_OUT_I16, "uint8": FSG_OUT_U8}
NONE = float("nan")  # "None" for optional doubles


class Encode(C.Structure):
    _fields_ = [("kind", C.c_int32), ("dn_min", C.c_int32), ("dn_max", C.c_int32), ("_pad", C.c_int32),
                ("a_coef", C.c_double), ("b_coef", C.c_double)]


class Window(C.Structure):
    _fields_ = [(n, C.c_int64) for n in
                ("H_global", "W", "buf_row0", "buf_rows", "out_row0", "out_rows", "ld_in", "ld_out")]


_P = C.c_void_p
_D = C.c_double
_I = C.c_int
_L = C.c_int64
_SIGNATURES = {
    "fsg_last_error": (C.c_char_p, []),
    "fsg_version": (_I, []),
    "fsg_launch_count": (_L, []),
    "fsg_reset_launch_count": (None, []),
    "fsg_profile_enable": (None, [_I]),
    "fsg_profile_read": (_I, [C.POINTER(C.c_int), C.POINTER(C.c_float), _I]),
    "fsg_hillshade": (_I, [_P, _P, C.POINTER(Window), _D, _D, _D, _D, _D, _D, C.POINTER(Encode), _P]),
    "fsg_slope": (_I, [_P, _P, C.POINTER(Window), _I, _D, _D, _D, C.POINTER(Encode), _P]),
    "fsg_curvature": (_I, [_P, _P, C.POINTER(Window), _I, _D, _D, _D, C.POINTER(Encode), _P]),
    "fsg_topousm_fast_workspace_bytes": (C.c_size_t, [_L, _L, C.POINTER(C.c_int32), _I, _D]),
    "fsg_topousm_fast": (_I, [_P, _P, _L, _L, _L, _L, C.POINTER(C.c_int32), C.POINTER(C.c_float), _I,
                              _D, _D, C.POINTER(Encode), _P, C.c_size_t, _P]),
    "fsg_topousm_fast_roi": (_I, [_P, _P, _L, _L, _L, _L, C.POINTER(C.c_int32), C.POINTER(C.c_float), _I,
                                  _D, _D, C.POINTER(Encode), _P, C.c_size_t, _L, _L, _L, _L, _P]),
    "fsg_topousm_plan": (_I, [C.POINTER(C.c_int32), _I, _D, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                              C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "fsg_pyramid_band": (_I, [_P, _L, _L, _L, C.POINTER(C.c_int32), _I, C.POINTER(_P), _P, _P]),
    "fsg_grid_mean_band_workspace_bytes": (C.c_size_t, [_L, _L]),
    "fsg_grid_mean_band": (_I, [_P, _L, _L, _L, _L, _L, _I, _L, _L, _P, _P, C.c_size_t, _P]),
    "fsg_topousm_fused_band": (_I, [_P, _L, _L, _L, _L, _L, _P, _L, _L, _L, C.POINTER(C.c_int32),
                                    C.POINTER(C.c_float), _I, _D, C.POINTER(_P), C.POINTER(_L), C.POINTER(_L),
                                    _D, C.POINTER(Encode), _P]),
    "fsg_topousm_fused_band_workspace_bytes": (C.c_size_t, [_L, _L]),
    "fsg_topousm_fused_band_ws": (_I, [_P, _L, _L, _L, _L, _L, _P, _L, _L, _L, C.POINTER(C.c_int32),
                                       C.POINTER(C.c_float), _I, _D, C.POINTER(_P), C.POINTER(_L), C.POINTER(_L),
                                       _D, _P, C.POINTER(Encode), _P, C.c_size_t, _P]),
    "fsg_debug_reload_switches": (None, []),
    "fsg_select_finish_scale": (_I, [_P, _I, C.c_float, C.c_float, _P, _P]),
    "fsg_valid_bbox": (_I, [_P, _L, _L, _L, _L, _L, _L, _P, _P]),
    "fsg_topousm_large_part": (_I, [_P, _P, _L, _L, _L, _L, _P, _L, _L, _L, _L, _L, _L, _L, _D, _P]),
    "fsg_openness": (_I, [_P, _P, C.POINTER(Window), _I, _I, _I, _D, _D, _D, _D, _D, C.POINTER(Encode), _P]),
    "fsg_ambient_occlusion_workspace_bytes": (C.c_size_t, [_L, _L]),
    "fsg_ambient_occlusion": (_I, [_P, _P, _L, _L, _L, _L, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                   C.POINTER(C.c_float), C.POINTER(C.c_float), _D, _D, _D, C.POINTER(Encode), _P,
                                   C.c_size_t, _P]),
    "fsg_overview_average": (_I, [_P, _P, _L, _L, _L, _L, _I, _D, _I, _P]),
    "fsg_decimate_workspace_bytes": (C.c_size_t, [_L, _L, _I]),
    "fsg_decimate": (_I, [_P, _P, _L, _L, _L, _I, _P, C.c_size_t, _P]),
    "fsg_upsample": (_I, [_P, _P, _L, _L, _L, _L, _P, C.c_size_t, _P]),
    "fsg_openness_samples": (_I, [_P, _P, C.POINTER(Window), _I, _I, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                  C.POINTER(C.c_int32), C.POINTER(C.c_float), _D, _D, C.POINTER(Encode), _P]),
    "fsg_encode_f32": (_I, [_P, _P, _L, C.POINTER(Encode), _P]),
    "fsg_scale_f32": (_I, [_P, _P, _L, _D, _P]),
    "fsg_stretch_f32": (_I, [_P, _P, _L, _D, _D, _P]),
    "fsg_gaussian_nan_workspace_bytes": (C.c_size_t, [_L, _L, _D]),
    "fsg_gaussian_nan": (_I, [_P, _P, _L, _L, _L, _D, _P, C.c_size_t, _P]),
    "fsg_combine_f32": (_I, [_P, _P, _L, _D, _I, _P]),
    "fsg_order_stats_workspace_bytes": (C.c_size_t, []),
    "fsg_order_stats": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _L, _I, _I,
                             _P, _P, C.c_size_t, _P]),
    "fsg_count_samples": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _I, _P, _P]),
    "fsg_grid_void_fill": (_I, [_P, _L, _L, _P, C.c_size_t, _P]),
    "fsg_key_histogram": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _I, C.c_uint32,
                               C.c_uint32, _I, _I, _P, _P, _P]),
    "fsg_key_rank_info": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, C.c_uint32, _I, _I,
                               _P, _P]),
    "fsg_key_to_float": (C.c_float, [C.c_uint32, _I]),
    "fsg_select_exchange_words": (C.c_size_t, []),
    "fsg_select_workspace_bytes": (C.c_size_t, []),
    "fsg_select_begin": (_I, [_P, C.c_size_t, _P]),
    "fsg_select_hist": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _I, _I, _I, _P, _P]),
    "fsg_select_pick": (_I, [_I, C.c_float, _P, _P]),
    "fsg_select_next": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _I, _I, _P, _P]),
    "fsg_select_finish": (_I, [_P, _I, _P, _P]),
    "fsg_debug_v8_band_rows": (C.c_int64, [_L, _L]),
    "fsg_select_compact": (_I, [C.POINTER(_P), C.POINTER(_L), C.POINTER(_L), C.POINTER(_L), _I, _I, _I, _P, _P, _L, _P]),
    "fsg_select_hist_keys": (_I, [_P, _I, _P, _P]),
    "fsg_select_next_keys": (_I, [_P, _P, _P]),
    "fsg_select_peer_slot_words": (C.c_size_t, []),
    "fsg_select_peer_publish": (_I, [_P, _P, _I, _P]),
    "fsg_select_peer_reduce": (_I, [_P, _P, _I, _I, _P]),
    "fsg_synth_dem": (_I, [_P, _L, _L, _L, _L, _L, C.c_uint64, _I, _P]),
    "fsg_copy_rect_f32": (_I, [_P, _L, _P, _L, _L, _L, _P]),
}

_lock = threading.Lock()
_lib = None


class FsgError(RuntimeError):
    pass


def exported_symbols():
    """Names include/fsg_b200.h declares (used by the CPU-side symbol test)."""
    return sorted(_SIGNATURES)


def load():
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise FsgError(
                    f"{LIB_PATH} is missing: build it with `python -m fujishadergpu_b200.build` "
                    "(there is no CPU or CuPy fallback for the hot path)")
            lib = C.CDLL(LIB_PATH)
            for name, (res, args) in _SIGNATURES.items():
                fn = getattr(lib, name)          # AttributeError if the symbol is not exported
                fn.restype = res
                fn.argtypes = args
            _lib = lib
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    msg = load().fsg_last_error().decode("utf-8", "replace")
    if rc == -1:
        raise ValueError(msg)
    raise FsgError(f"{what or 'libfsg_b200'} failed (code {rc}): {msg}")


def make_encode(output_dtype: str = "float32", qp: dict | None = None) -> Encode:
    kind = _KIND[str(output_dtype)]
    if kind != FSG_OUT_F32 and qp is None:
        # the kernels would store 4-byte floats into a 1- / 2-byte-per-pixel buffer
        raise ValueError(f"output_dtype={output_dtype!r} needs quantisation parameters (qp=quantize_params(...))")
    if kind == FSG_OUT_F32:
        return Encode(FSG_OUT_F32, 0, 0, 0, 1.0, 0.0)
    return Encode(kind, int(qp["dn_min"]), int(qp["dn_max"]), 0, float(qp["a_coef"]), float(qp["b_coef"]))


def opt(v) -> float:
    """Optional double -> NaN sentinel."""
    return NONE if v is None else float(v)
