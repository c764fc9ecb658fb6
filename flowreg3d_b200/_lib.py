"""ctypes binding of libfr3d.so (the C ABI declared in include/fr3d.h).

There is no CPU fallback: if the CUDA library is missing or no B200 is visible the package raises.
The library that is loaded is ALWAYS flowreg3d_b200/libfr3d.so: no environment variable redirects the loader.
(tests/emu builds a CPU emulation of the kernel logic for the not-gpu tests; only test code can install it, by
calling _select_for_tests() in its own process -- it is never shipped and never reachable from the product.)
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
MAX_CHANNELS = 4
ABI_VERSION = 5

F32, F64, U8, U16, I16, I32 = range(6)
OPT_SOR_CTAS_PER_SM = 1
OPT_CC_BLOCK_SCANS = 2
OPT_WARP_FACTORED = 3
OPT_SOR_KERNEL = 4
OPT_SOR_STAGES = 5
OPT_SOR_SCHED = 9
OPT_SOR_FRAMES_PER_ITEM = 10
OPT_WARP_TILE = 11
OPT_SOR_TILE = 6
OPT_SPLINE_TMA = 7
OPT_RESIZE_X_ROWS = 8
_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float64): F64, np.dtype(np.uint8): U8,
           np.dtype(np.uint16): U16, np.dtype(np.int16): I16, np.dtype(np.int32): I32}


def dtype_code(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _DTYPES:
        raise TypeError(f"libfr3d does not take dtype {dt}; supported: {[str(k) for k in _DTYPES]}")
    return _DTYPES[dt]


class AxisTable(C.Structure):
    _fields_ = [("in_len", C.c_int32), ("out_len", C.c_int32), ("P", C.c_int32),
                ("idx", C.c_void_p), ("wt", C.c_void_p)]


class Level(C.Structure):
    _fields_ = [("size", C.c_int32 * 3), ("h", C.c_double * 3), ("alpha", C.c_double * 3),
                ("median", C.c_int32), ("from_full", AxisTable * 3), ("from_prev", AxisTable * 3)]


class Plan(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("Z", C.c_int32), ("Y", C.c_int32), ("X", C.c_int32),
                ("C", C.c_int32), ("max_batch", C.c_int32), ("n_levels", C.c_int32),
                ("levels", C.POINTER(Level)), ("to_full", AxisTable * 3),
                ("iterations", C.c_int32), ("update_lag", C.c_int32),
                ("a_data", C.c_double * MAX_CHANNELS), ("a_smooth", C.c_double),
                ("sweep", C.c_int32), ("interp", C.c_int32), ("state_dtype", C.c_int32),
                ("gauss_radius", (C.c_int32 * 3) * MAX_CHANNELS),
                ("gauss_w", (C.c_void_p * 3) * MAX_CHANNELS),
                ("gauss_radius_t", C.c_int32 * MAX_CHANNELS), ("gauss_w_t", C.c_void_p * MAX_CHANNELS)]


class Fr3dError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libfr3d error {code}: {msg}")
        self.code = code


_lib = None
_test_library = None     # (path, is_emulator) installed by _select_for_tests(); None in the product


def library_path() -> Path:
    return Path(_test_library[0]) if _test_library is not None else HERE / "libfr3d.so"


def load():
    """Load libfr3d.so; raise loudly when it is absent (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = library_path()
    if not path.exists():
        raise ImportError(
            f"{path} not found: build it with `python -m flowreg3d_b200.build` (needs nvcc). "
            "flowreg3d_b200 has no CPU fallback.")
    L = C.CDLL(str(path))
    vp, ci, cd, i64 = C.c_void_p, C.c_int, C.c_double, C.c_int64
    sig = {
        "fr3d_abi_version": (ci, []),
        "fr3d_create": (ci, [C.POINTER(vp), ci, C.POINTER(Plan), vp]),
        "fr3d_destroy": (None, [vp]),
        "fr3d_last_error": (C.c_char_p, [vp]),
        "fr3d_synchronize": (ci, [vp]),
        "fr3d_launch_count": (i64, [vp]),
        "fr3d_device_bytes": (i64, [vp]),
        "fr3d_set_option": (ci, [vp, ci, i64]),
        "fr3d_preprocess": (ci, [vp, vp, ci, ci, vp, vp, ci, vp, vp]),
        "fr3d_set_reference": (ci, [vp, vp, vp, vp]),
        "fr3d_get_displacement": (ci, [vp, vp, vp, ci, vp, ci]),
        "fr3d_level_count": (ci, [vp]),
        "fr3d_level_info": (ci, [vp, ci, vp, vp, vp, vp]),
        "fr3d_level_begin": (ci, [vp, ci, vp, vp, ci]),
        "fr3d_level_sweeps": (ci, [vp, ci, ci, ci, ci, ci]),
        "fr3d_level_state": (ci, [vp, ci, ci, vp, i64, i64]),
        "fr3d_level_sweeps_slab": (ci, [vp, ci, ci, ci, ci, ci]),
        "fr3d_level_planes": (ci, [vp, ci, ci, vp, ci, ci]),
        "fr3d_ipc_export": (ci, [vp, ci, vp]),
        "fr3d_ipc_open": (ci, [vp, vp, C.POINTER(vp)]),
        "fr3d_ipc_close": (ci, [vp, vp]),
        "fr3d_level_sweeps_slab_p2p": (ci, [vp, ci, ci, ci, vp, vp, vp, vp, i64]),
        "fr3d_level_wave_cells": (ci, [vp, ci, ci, vp, ci, ci]),
        "fr3d_level_end": (ci, [vp, ci]),
        "fr3d_level_end_range": (ci, [vp, ci, ci, ci]),
        "fr3d_flow_slab": (ci, [vp, ci, ci, vp, ci, ci]),
        "fr3d_flow_finish": (ci, [vp, vp, ci]),
        "fr3d_compensate": (ci, [vp, vp, ci, vp, vp, ci, ci, vp]),
        "fr3d_resize3d": (ci, [vp, vp, ci, ci, ci, ci, C.POINTER(AxisTable), vp]),
        "fr3d_warp": (ci, [vp, vp, ci, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp]),
        "fr3d_motion_tensor": (ci, [vp, vp, vp, ci, ci, ci, cd, cd, cd, ci, vp]),
        "fr3d_motion_tensor_alt": (ci, [vp, ci, vp, vp, ci, ci, ci, cd, cd, cd, vp]),
        "fr3d_sor_level": (ci, [vp, vp, vp, vp, ci, ci, ci, ci, vp, cd, cd, cd, ci, ci, vp, cd, ci, ci, vp]),
        "fr3d_median5": (ci, [vp, vp, ci, ci, ci, ci, vp]),
        "fr3d_mean_frames": (ci, [vp, vp, ci, i64, vp]),
        "fr3d_mean_frames_f64": (ci, [vp, vp, ci, i64, vp]),
        "fr3d_flow_stats": (ci, [vp, vp, ci, ci, ci, ci, vp]),
        "fr3d_warp_flow": (ci, [vp, vp, ci, vp, vp, ci, ci, ci, ci, ci, ci, ci, vp]),
        "fr3d_cc_project": (ci, [vp, vp, ci, ci, ci, ci, ci, vp, vp]),
        "fr3d_cc_window": (ci, [vp, vp, ci, ci, ci, vp, vp, vp]),
        "fr3d_cc_cgemm": (ci, [vp, vp, i64, vp, i64, vp, ci, ci, ci, ci]),
        "fr3d_cc_cross_power": (ci, [vp, vp, vp, vp, i64, ci, ci]),
        "fr3d_cc_abs_argmax": (ci, [vp, vp, i64, ci, vp]),
        "fr3d_cc_wrap_shift": (ci, [vp, vp, ci, ci, ci, vp, vp, vp]),
        "fr3d_cc_tile_sums": (ci, [vp, vp, vp, ci, ci, ci, vp, vp]),
        "fr3d_rigid_flow": (ci, [vp, vp, vp, ci, i64, vp]),
        "fr3d_add_flow": (ci, [vp, vp, vp, i64, vp]),
        "fr3d_profile_enable": (ci, [vp, ci]),
        "fr3d_profile_report": (i64, [vp, C.c_char_p, i64]),
        "fr3d_fill_resize_table": (ci, [ci, ci, vp, ci, vp, vp]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        f.restype = res
        f.argtypes = args
    if L.fr3d_abi_version() != ABI_VERSION:
        raise ImportError(f"{path}: ABI {L.fr3d_abi_version()} != binding ABI {ABI_VERSION}")
    _lib = L
    return L


EXPORTED_SYMBOLS = [
    "fr3d_abi_version", "fr3d_create", "fr3d_destroy", "fr3d_last_error", "fr3d_synchronize",
    "fr3d_launch_count", "fr3d_device_bytes", "fr3d_set_option", "fr3d_preprocess", "fr3d_set_reference",
    "fr3d_get_displacement", "fr3d_level_count", "fr3d_level_info", "fr3d_level_begin", "fr3d_level_sweeps",
    "fr3d_level_state", "fr3d_level_end", "fr3d_level_end_range", "fr3d_flow_slab", "fr3d_flow_finish",
    "fr3d_compensate", "fr3d_resize3d", "fr3d_warp", "fr3d_motion_tensor",
    "fr3d_sor_level", "fr3d_median5", "fr3d_mean_frames", "fr3d_mean_frames_f64", "fr3d_flow_stats", "fr3d_profile_enable",
    "fr3d_profile_report", "fr3d_fill_resize_table",
    "fr3d_warp_flow", "fr3d_cc_project", "fr3d_cc_window", "fr3d_cc_cgemm", "fr3d_cc_cross_power",
    "fr3d_cc_abs_argmax", "fr3d_cc_wrap_shift", "fr3d_cc_tile_sums", "fr3d_rigid_flow", "fr3d_add_flow",
    "fr3d_level_sweeps_slab", "fr3d_level_planes", "fr3d_level_wave_cells",
    "fr3d_ipc_export", "fr3d_ipc_open", "fr3d_ipc_close", "fr3d_level_sweeps_slab_p2p", "fr3d_motion_tensor_alt",
]


def is_emulator() -> bool:
    return _test_library is not None and bool(_test_library[1])


def _select_for_tests(path, emulator: bool = True) -> None:
    """tests/ and tools/ only: point the binding of THIS process at the kernel-logic emulator (emulator=True: CPU
    tensors, tests/emu) or at a tuning build of the CUDA library (emulator=False, tools/variants.sh), or back at
    flowreg3d_b200/libfr3d.so with None, and drop everything cached from the previous choice."""
    global _lib, _test_library
    _test_library = None if path is None else (str(path), bool(emulator))
    _lib = None
    from . import core
    for c in core._bare.values():
        c.h = None  # the old library owns them; do not destroy through the new one
    core._bare.clear()
    for r in core._PAIR_CACHE.values():
        r.ctx.h = None
    core._PAIR_CACHE.clear()
