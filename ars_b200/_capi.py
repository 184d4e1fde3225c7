"""ctypes binding of libars_b200.so (C ABI in include/ars_b200.h).

There is no CPU implementation behind this module: if the shared object is missing or no
sm_100 device can be initialised, every compute call raises `ArsError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libars_b200.so")

LAYOUT_IDS = {"Stereo": 0, "5.1 (Standard)": 1, "7.1 (Surround)": 2, "5.1.2 (Atmos Light)": 3}
LUFS_OK, LUFS_NONE, LUFS_SKIPPED = 0, 1, 2


class ArsError(RuntimeError):
    """A call into libars_b200 failed (or the library / the GPU is not available)."""


class ArsMetrics(C.Structure):
    _fields_ = [("lufs", C.c_double), ("lufs_status", C.c_int32), ("reserved", C.c_int32),
                ("peak_linear", C.c_double), ("rms_linear", C.c_double),
                ("true_peak_dbfs", C.c_double), ("rms_dbfs", C.c_double),
                ("true_peak_4x_dbfs", C.c_double), ("true_peak_4x_status", C.c_int32), ("reserved2", C.c_int32)]


class ArsIrDraws(C.Structure):
    _fields_ = [("tap_delay", C.c_void_p), ("tap_base", C.c_void_p), ("ntaps", C.c_int32),
                ("reserved", C.c_int32), ("noise", C.c_void_p), ("noise_len", C.c_int64)]


class ArsRenderParams(C.Structure):
    _fields_ = [("rate", C.c_double), ("external_ir", C.c_int32), ("layout", C.c_int32),
                ("ir_duration", C.c_double), ("ir_max_delay", C.c_double), ("absorption", C.c_double),
                ("directionality", C.c_double), ("ir_split_time", C.c_double), ("diffusion", C.c_double),
                ("early_level", C.c_double), ("late_level", C.c_double), ("dry_wet", C.c_double),
                ("kill_start", C.c_double), ("bass_gain", C.c_double), ("treble_gain", C.c_double),
                ("air_absorption", C.c_double), ("x", C.c_double), ("y", C.c_double), ("z", C.c_double),
                ("want_lufs", C.c_int32), ("reserved", C.c_int32)]


class ArsLongPlan(C.Structure):
    _fields_ = [("block_frames", C.c_int64), ("n_blocks", C.c_int64), ("halo_frames", C.c_int64),
                ("frames_out", C.c_int64), ("hop_count", C.c_int32), ("route", C.c_int32)]


class ArsClip(C.Structure):
    _fields_ = [("params", C.POINTER(ArsRenderParams)), ("in_", C.c_void_p), ("n", C.c_int64), ("cin", C.c_int32),
                ("reserved", C.c_int32), ("ext_ir", C.c_void_p), ("ext_ir_len", C.c_int64),
                ("draws", C.POINTER(ArsIrDraws)), ("out_f32", C.c_void_p), ("out_pcm", C.c_void_p),
                ("metrics", C.POINTER(ArsMetrics))]


_d, _i32, _i64, _p = C.c_double, C.c_int32, C.c_int64, C.c_void_p

# name -> (restype, argtypes); must list every symbol include/ars_b200.h declares
PROTOTYPES = {
    "ars_init": (C.c_int, [C.c_int]),
    "ars_shutdown": (None, []),
    "ars_sync": (C.c_int, []),
    "ars_last_error": (C.c_char_p, []),
    "ars_version": (C.c_char_p, []),
    "ars_launch_count": (C.c_uint64, []),
    "ars_air_fold_count": (C.c_uint64, []),
    "ars_olsb_count": (C.c_uint64, []),
    "ars_head_start_count": (C.c_uint64, []),
    "ars_meter_stream_count": (C.c_uint64, []),
    "ars_tail_overlap_count": (C.c_uint64, []),
    "ars_stream": (C.c_void_p, []),
    "ars_ir_synth": (C.c_int, [_d, _d, _d, _d, _d, _d, _d, C.POINTER(ArsIrDraws), _p, _p, _i64]),
    "ars_ir_geometry": (C.c_int, [_d, _d, _d, _d, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                  C.POINTER(_i64)]),
    "ars_air_filter": (C.c_int, [_p, _i64, _d, _d, _p]),
    "ars_resample": (C.c_int, [_p, _i64, _i64, _p]),
    "ars_dry_wet_mix": (C.c_int, [_p, _i64, _p, _i64, _i32, _d, _d, _p]),
    "ars_convolve_out_len": (_i64, [_i64, _i64, _i64]),
    "ars_convolve_split": (C.c_int, [_p, _i64, _i32, _p, _i64, _p, _i64, _d, _d, _d, _d, _d, _d, _d, _d, _p]),
    "ars_convolve_external": (C.c_int, [_p, _i64, _i32, _p, _i64, _d, _d, _d, _d, _d, _p]),
    "ars_pan": (C.c_int, [_p, _i64, _d, _d, _d, _p]),
    "ars_delay": (C.c_int, [_p, _i64, _i32, _i64, _p]),
    "ars_layout_channels": (C.c_int, [_i32]),
    "ars_map_channels": (C.c_int, [_p, _i64, _i32, _d, _d, _p]),
    "ars_metrics": (C.c_int, [_p, _i64, _i32, _d, _i32, C.POINTER(ArsMetrics)]),
    "ars_true_peak_4x": (C.c_int, [_p, _i64, _i32, C.POINTER(_d)]),
    "ars_channel_rms": (C.c_int, [_p, _i64, _i32, _p, C.POINTER(C.c_float)]),
    "ars_spectrogram_segments": (_i64, [_i64, _i32]),
    "ars_spectrogram": (C.c_int, [_p, _i64, _i32, _d, _i32, _p]),
    "ars_pcm16": (C.c_int, [_p, _i64, _p]),
    "ars_render_out_len": (_i64, [C.POINTER(ArsRenderParams), _i64, _i64]),
    "ars_render": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i64, _i32, _p, _i64, C.POINTER(ArsIrDraws),
                             _p, _p, _p, C.POINTER(ArsMetrics)]),
    "ars_render_dev": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i64, _i32, _p, _i64, C.POINTER(ArsIrDraws),
                                 _p, _p, _p, C.POINTER(ArsMetrics)]),
    "ars_render_dev_async": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i64, _i32, _p, _i64, C.POINTER(ArsIrDraws),
                                 _p, _p, _p, C.POINTER(ArsMetrics)]),
    "ars_render_batch": (C.c_int, [C.POINTER(ArsClip), _i32]),
    "ars_set_option": (C.c_int, [C.c_char_p, _i32]),
    "ars_state_bytes": (_i64, []),
    "ars_ols_block_frames": (_i64, []),
    "ars_long_plan": (C.c_int, [C.POINTER(ArsRenderParams), _i64, _i64, C.POINTER(ArsLongPlan)]),
    "ars_long_loudness_hops_dev": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i64, _i64, _i64, _i64, _p, _p, _i32]),
    "ars_long_loudness_gate_dev": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i32, _i64, _p, C.POINTER(_i32)]),
    "ars_long_convolve_dev": (C.c_int, [C.POINTER(ArsRenderParams), _p, _i64, _i64, _i64, _i32, _p, _i64, _p, _i64, _i64,
                                        _i64, _p, _i64, _p]),
    "ars_long_tail_dev": (C.c_int, [C.POINTER(ArsRenderParams), _i32, _p, _i64, _i64, _i64, _i64, _p, _p, _p, _p]),
    "ars_loudness_dev": (C.c_int, [_p, _i64, _d, _p, C.POINTER(_i32)]),
    "ars_state_metrics": (C.c_int, [_p, _i64, _i32, C.POINTER(ArsMetrics)]),
    "ars_timer_begin": (C.c_int, []),
    "ars_timer_end": (C.c_int, [C.POINTER(C.c_float)]),
    "ars_profile_begin": (C.c_int, []),
    "ars_profile_end": (C.c_int, [C.POINTER(_i64), C.POINTER(_d), C.POINTER(_d)]),
    "ars_profile_report": (C.c_char_p, []),
    "ars_peer_alloc": (C.c_int, [_i64, C.POINTER(C.c_void_p), C.c_char_p]),
    "ars_peer_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "ars_peer_push": (C.c_int, [_p, _p, _i64, _p]),
    "ars_peer_close": (C.c_int, [_p]),
    "ars_peer_free": (C.c_int, [_p]),
    "ars_host_alloc": (C.c_void_p, [_i64]),
    "ars_host_free": (None, [C.c_void_p]),
}

_lib = None
_lock = threading.Lock()
_inited = False


def load_library():
    """dlopen libars_b200.so and attach prototypes.  Raises ArsError if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ArsError(f"{LIB_PATH} is missing: build it with `python -m ars_b200.build` "
                           "(the render path has no CPU fallback)")
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:
            raise ArsError(f"cannot load {LIB_PATH}: {e}") from e
        for name, (res, args) in PROTOTYPES.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as e:
                raise ArsError(f"{LIB_PATH} does not export {name}; rebuild it") from e
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error() -> str:
    msg = load_library().ars_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(code: int, what: str):
    if code != 0:
        raise ArsError(f"{what} failed (code {code}): {last_error()}")


def init(device: int | None = None):
    """Initialise the CUDA context of the library (idempotent).  Raises without a B200-class GPU."""
    global _inited
    lib = load_library()
    if _inited and device is None:
        return lib
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", os.environ.get("ARS_DEVICE", "0")))
    check(lib.ars_init(int(device)), "ars_init")
    _inited = True
    return lib


def set_option(key: str, value: int):
    check(init().ars_set_option(key.encode(), int(value)), "ars_set_option")


def shutdown():
    global _inited
    if _lib is not None:
        _lib.ars_shutdown()
    _inited = False


class _PinnedBlock:
    """A page-locked block of the library wrapped for numpy (array interface); goes back to the library's pool when the
    last array that views it dies."""

    def __init__(self, lib, address, nbytes):
        self._lib, self._address = lib, address
        self.__array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (address, False), "version": 3}

    def __del__(self):
        try:
            self._lib.ars_host_free(self._address)
        except Exception:
            pass


def result_empty(shape, dtype) -> np.ndarray:
    """np.empty for a render's results: large arrays come from the library's pinned pool, so the device -> host copy runs
    at bus speed and without page faults; the caller gets an ordinary, writable ndarray."""
    dt = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dt.itemsize
    if nbytes < (4 << 20) or os.environ.get("ARS_PINNED_RESULTS", "1") == "0":
        return np.empty(shape, dt)
    lib = init()
    address = lib.ars_host_alloc(nbytes)
    if not address:
        return np.empty(shape, dt)
    return np.asarray(_PinnedBlock(lib, address, nbytes)).view(dt).reshape(shape)


def ptr(a) -> int | None:
    """Address of a C-contiguous numpy array (None stays NULL)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data


def f32c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float32)


def make_draws(tap_delay, tap_base, noise, keep: list) -> ArsIrDraws:
    """Pack the replayed random draws; `keep` receives the arrays that must outlive the call."""
    d = ArsIrDraws()
    td = np.ascontiguousarray(tap_delay, dtype=np.int64)
    tb = np.ascontiguousarray(tap_base, dtype=np.float64)
    keep += [td, tb]
    d.tap_delay = ptr(td) if td.size else None
    d.tap_base = ptr(tb) if tb.size else None
    d.ntaps = int(td.size)
    if isinstance(noise, int):          # a device pointer (ars_render_dev)
        d.noise = noise
    else:
        nz = np.ascontiguousarray(noise, dtype=np.float64)
        keep.append(nz)
        d.noise = ptr(nz) if nz.size else None
        d.noise_len = int(nz.size)
    return d
