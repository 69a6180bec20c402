"""ctypes binding of csrc/libtomatis_b200.so (the C ABI in include/tomatis_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing
callers get a RuntimeError telling them to build it.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes as C
import os

from .build import LIB_PATH, LIB_PATHS

OK = 0
ERR_UNSUPPORTED = -3
FRAMING_STREAMING, FRAMING_WHOLEFILE, FRAMING_EQ_PAD, FRAMING_EQ_NOPAD = 0, 1, 2, 3
GATE_UPDELAY, GATE_MINHOLD = 0, 1
PCM_S16, PCM_S24 = 0, 1
LEVELS_F64, LEVELS_MONO, LEVELS_HOPSUM_ONLY, LEVELS_MEANSQ_ONLY, LEVELS_LEFT, LEVELS_RIGHT, LEVELS_POWER_EPS = 1, 2, 8, 16, 32, 64, 128
(ARR_MEANSQ_F32, ARR_MEANSQ_F64, ARR_GATE_F64, ARR_STATE, ARR_ROW, ARR_C2_COUNT, ARR_CHUNK_PEAK,
 ARR_INPUT_PEAK, ARR_HOPSUM_F32, ARR_HOPSUM_F64, ARR_BISECT_T, ARR_BISECT_ITERS, ARR_BISECT_TRACE_T,
 ARR_BISECT_TRACE_C2) = range(14)
BISECT_MAX_ITER = 32


class TrackDesc(C.Structure):
    """struct tmt_track_desc"""
    _fields_ = [("pcm_in", C.c_void_p), ("pcm_out", C.c_void_p), ("total", C.c_int64),
                ("in_origin", C.c_int64), ("in_len", C.c_int64),
                ("out_origin", C.c_int64), ("out_len", C.c_int64),
                ("block_lo", C.c_int64), ("block_hi", C.c_int64)]


# every symbol include/tomatis_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "tmt_version": (C.c_int, []),
    "tmt_error_string": (C.c_char_p, [C.c_int]),
    "tmt_last_error": (C.c_char_p, []),
    "tmt_engine_create": (C.c_int, [C.POINTER(_P), C.c_int, C.c_int, C.c_int]),
    "tmt_engine_destroy": (C.c_int, [_P]),
    "tmt_engine_set_window": (C.c_int, [_P, _P, C.c_int]),
    "tmt_engine_set_gain_rows": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "tmt_plan_create": (C.c_int, [_P, C.POINTER(_P), C.c_int, C.c_int, C.POINTER(TrackDesc), C.c_int]),
    "tmt_plan_destroy": (C.c_int, [_P]),
    "tmt_plan_set_buffers": (C.c_int, [_P, C.c_int, _P, _P]),
    "tmt_plan_set_level_ranges": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "tmt_plan_total_frames": (C.c_int, [_P]),
    "tmt_plan_total_chunks": (C.c_int, [_P]),
    "tmt_plan_total_units": (C.c_int, [_P]),
    "tmt_plan_unfusable_chunks": (C.c_int, [_P]),
    "tmt_plan_debug_counters": (C.c_int, [_P, _P, C.c_int]),
    "tmt_plan_track_frames": (C.c_int, [_P, C.c_int]),
    "tmt_plan_track_frame_base": (C.c_int, [_P, C.c_int]),
    "tmt_plan_track_chunks": (C.c_int, [_P, C.c_int]),
    "tmt_plan_track_chunk_base": (C.c_int, [_P, C.c_int]),
    "tmt_plan_chunk_range": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "tmt_plan_read": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P, C.c_int, _P]),
    "tmt_plan_read_many": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "tmt_plan_write": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int64, _P, C.c_int, _P]),
    "tmt_plan_input_peaks": (C.c_int, [_P, _P]),
    "tmt_plan_levels": (C.c_int, [_P, C.c_int, _P, _P]),
    "tmt_plan_levels_multichannel": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, _P]),
    "tmt_channels_split": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P]),
    "tmt_channels_merge": (C.c_int, [_P, C.c_int64, C.c_int, _P, _P]),
    "tmt_plan_gate": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "tmt_peer_bytes": (C.c_size_t, [C.c_int]),
    "tmt_peer_alloc": (C.c_int, [C.c_int, C.c_size_t, _P, _P]),
    "tmt_peer_open": (C.c_int, [C.c_int, _P, _P]),
    "tmt_peer_close": (C.c_int, [C.c_int, _P]),
    "tmt_peer_free": (C.c_int, [C.c_int, _P]),
    "tmt_peer_status": (C.c_int, [C.c_int, _P, _P]),
    "tmt_plan_peer_publish": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P]),
    "tmt_plan_peer_wait": (C.c_int, [_P, C.c_int, _P, _P, C.c_int64, C.c_int, C.c_int, C.c_double, _P]),
    "tmt_plan_bisect": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_int, C.c_int, _P]),
    "tmt_plan_bisect_gate": (C.c_int, [_P, _P, _P, _P, _P, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "tmt_plan_stft": (C.c_int, [_P, C.c_float, C.c_int, _P]),
    "tmt_plan_stft_limited": (C.c_int, [_P, C.c_float, C.c_float, _P]),
    "tmt_plan_clear_peaks": (C.c_int, [_P, _P]),
    "tmt_plan_edge_frames": (C.c_int, [_P, C.c_float, _P, _P, C.c_int, _P]),
    "tmt_plan_stft_with_edges": (C.c_int, [_P, C.c_float, _P, _P, C.c_int, _P]),
    "tmt_plan_limiter": (C.c_int, [_P, C.c_float, _P]),
    "tmt_cond_spectrum": (C.c_int, [_P, _P, _P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "tmt_calib_envelope_decimate": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int, C.c_int, C.c_int, C.c_int64, C.c_int64, _P, _P]),
    "tmt_calib_xcorr_valid": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, _P, _P]),
    "tmt_calib_band_energies": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P]),
    "tmt_calib_gate_grid": (C.c_int, [_P, _P, _P, _P, C.c_int, _P, _P, _P, C.c_int, _P, _P, _P, _P]),
    "tmt_plan_run_streaming": (C.c_int, [_P, C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "tmt_plan_pcm_levels": (C.c_int, [_P, _P, C.c_int64, C.c_int, _P]),
    "tmt_plan_run_streaming_pcm": (C.c_int, [_P, _P, C.c_int64, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_float, C.c_float, _P]),
    "tmt_plan_launch_count": (C.c_int64, [_P]),
    "tmt_generic_meansq": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, _P, _P]),
    "tmt_generic_frames": (C.c_int, [_P, _P, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.c_float, C.c_int, _P, _P]),
    "tmt_generic_overlap_add": (C.c_int, [_P, _P, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, _P, C.c_float, _P, _P]),
    "tmt_generic_limit": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, C.c_double, _P, _P]),
    "tmt_generic_to_float": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "tmt_plan_geometry": (C.c_int, [_P, _P, _P, _P, _P]),
    "tmt_plan_chunk_ranges": (C.c_int, [_P, _P]),
    "tmt_pcm_to_float": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P]),
    "tmt_float_to_pcm": (C.c_int, [_P, C.c_int, C.c_int64, _P, _P]),
    "tmt_requantise_scale": (C.c_int, [_P, C.c_int64, C.c_float, _P]),
}

_libs = {}
_last = None          # the library whose function was fetched last: its thread-local error text belongs to the failing call


class TomatisError(RuntimeError):
    pass


class _Lib:
    """One loaded build of the library.  Attribute access hands out the ctypes function and remembers which build was used,
    so that check() asks the right one for its error text (each .so has its own thread-local message buffer)."""

    def __init__(self, cdll, n_fft):
        object.__setattr__(self, "_cdll", cdll)
        object.__setattr__(self, "n_fft", n_fft)

    def __getattr__(self, name):
        global _last
        _last = self._cdll
        return getattr(self._cdll, name)


def load(n_fft: int = 4096):
    """Load the library built for `n_fft` (once).  Raises RuntimeError if it has not been built."""
    hit = _libs.get(n_fft)
    if hit is not None:
        return hit
    if n_fft not in LIB_PATHS:
        raise RuntimeError(f"no fused library for n_fft={n_fft} (built sizes: {sorted(LIB_PATHS)})")
    path = os.environ.get("TMT_LIB", LIB_PATH) if n_fft == 4096 else os.environ.get("TMT_LIB_2048", LIB_PATHS[n_fft])   # dev: A/B builds
    if not os.path.exists(path):
        raise RuntimeError(
            f"CUDA library not built: {path} is missing. Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(needs nvcc); there is no CPU fallback.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)       # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _libs[n_fft] = _Lib(lib, n_fft)
    return _libs[n_fft]


def check(rc: int, what: str = ""):
    if rc != OK:
        lib = _last if _last is not None else load()
        msg = lib.tmt_last_error().decode("utf-8", "replace")
        kind = lib.tmt_error_string(rc).decode()
        raise TomatisError(f"{what or 'libtomatis_b200'}: {kind} ({rc}): {msg}")
