"""ctypes binding of libequilibrium_cuda.so (include/equilibrium_cuda.h).

No torch types cross this boundary and there is no CPU fallback: if the shared
library is missing, or no CUDA device is visible when a fluid is created, the
call raises.  Build the library with ``python -m equilibrium_b200.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_LIB = os.path.join(_HERE, "libequilibrium_cuda.so")

EQ_OK = 0
MODE_EXACT, MODE_RED_BLACK = 0, 1
ROW, COL, PASSIVE = 0, 1, 2
F_DENSITY, F_VX, F_VY, F_VX0, F_VY0, F_SCRATCH, F_CELLS = range(7)


class EqParams(C.Structure):
    _fields_ = [
        ("size", C.c_uint32),
        ("delta_t", C.c_float),
        ("frames", C.c_int64),
        ("gs_iterations", C.c_int64),
        ("diffusion", C.c_float),
        ("viscosity", C.c_float),
        ("mode", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world", C.c_int32),
        ("comm_id", C.c_uint8 * 128),
    ]


class EqSource(C.Structure):
    _fields_ = [
        ("frame", C.c_int64),
        ("x", C.c_uint32),
        ("y", C.c_uint32),
        ("d_vx", C.c_float),
        ("d_vy", C.c_float),
        ("d_density", C.c_float),
    ]


class EqProfile(C.Structure):
    _fields_ = [
        ("lin_solve_ms", C.c_double), ("lin_solve_launches", C.c_int64), ("lin_solve_cell_iters", C.c_int64),
        ("advect_ms", C.c_double), ("advect_launches", C.c_int64),
        ("project_ms", C.c_double), ("project_launches", C.c_int64),
        ("boundary_ms", C.c_double), ("boundary_launches", C.c_int64),
        ("other_ms", C.c_double), ("other_launches", C.c_int64),
        ("steps", C.c_int64),
    ]


class EqNoise(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("first_frame", C.c_uint64), ("cos_t", C.c_float), ("sin_t", C.c_float),
                ("gain", C.c_float), ("reserved", C.c_float)]


class EqColors(C.Structure):
    _fields_ = [("world", C.c_uint8 * 4), ("fluid", C.c_uint8 * 4), ("obstacle", C.c_uint8 * 4)]


SNAP_DENSITY, SNAP_RGBA = 0, 1
SNAPSHOT_SLOTS = 2


class EquilibriumError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"equilibrium_cuda error {code}: {message}")
        self.code = code


# every symbol include/equilibrium_cuda.h declares: name -> (restype, argtypes)
_H = C.c_void_p
SIGNATURES = {
    "eq_last_error": (C.c_char_p, []),
    "eq_abi_version": (C.c_int, []),
    "eq_device_count": (C.c_int, []),
    "eq_create": (C.c_int, [C.POINTER(EqParams), C.POINTER(_H)]),
    "eq_destroy": (C.c_int, [_H]),
    "eq_clone": (C.c_int, [_H, C.POINTER(_H)]),
    "eq_init_default": (C.c_int, [_H]),
    "eq_add_density": (C.c_int, [_H, C.c_uint32, C.c_uint32, C.c_float]),
    "eq_add_velocity": (C.c_int, [_H, C.c_uint32, C.c_uint32, C.c_float, C.c_float]),
    "eq_rect_valid": (C.c_int, [C.c_int64] * 5),
    "eq_fill_rect": (C.c_int, [_H, C.c_int64, C.c_int64, C.c_int64, C.c_int64]),
    "eq_reset_walls": (C.c_int, [_H]),
    "eq_set_params": (C.c_int, [_H, C.POINTER(EqParams)]),
    "eq_get_params": (C.c_int, [_H, C.POINTER(EqParams)]),
    "eq_step": (C.c_int, [_H]),
    "eq_step_n": (C.c_int, [_H, C.c_int64, C.POINTER(EqSource), C.c_int64]),
    "eq_add_noise": (C.c_int, [_H, C.POINTER(EqNoise)]),
    "eq_step_n_noise": (C.c_int, [_H, C.c_int64, C.POINTER(EqNoise)]),
    "eq_sync": (C.c_int, [_H]),
    "eq_upload": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_size_t]),
    "eq_download": (C.c_int, [_H, C.c_int, C.c_void_p, C.c_size_t]),
    "eq_upload_rows": (C.c_int, [_H, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]),
    "eq_download_rows": (C.c_int, [_H, C.c_int, C.c_uint32, C.c_uint32, C.c_void_p]),
    "eq_owned_rows": (C.c_int, [_H, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
    "eq_op_set_boundaries": (C.c_int, [_H, C.c_int, C.c_int]),
    "eq_op_lin_solve": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int64]),
    "eq_op_diffuse": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int64]),
    "eq_op_project": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int64]),
    "eq_op_advect": (C.c_int, [_H, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]),
    "eq_op_add_source": (C.c_int, [_H, C.c_int, C.c_int, C.c_float]),
    "eq_snapshot_begin": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(EqColors), C.c_void_p, C.c_size_t]),
    "eq_snapshot_wait": (C.c_int, [_H, C.c_int]),
    "eq_render_rgba": (C.c_int, [_H, C.POINTER(EqColors), C.c_void_p, C.c_size_t]),
    "eq_divergence_l2": (C.c_int, [_H, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "eq_set_stream": (C.c_int, [_H, C.c_void_p]),
    "eq_timer_start": (C.c_int, [_H]),
    "eq_timer_stop": (C.c_int, [_H, C.POINTER(C.c_float)]),
    "eq_profile_enable": (C.c_int, [_H, C.c_int]),
    "eq_profile_reset": (C.c_int, [_H]),
    "eq_profile_get": (C.c_int, [_H, C.POINTER(EqProfile)]),
    "eq_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_size_t]),
    "eq_host_free": (C.c_int, [C.c_void_p]),
    "eq_l2_flush": (C.c_int, [_H]),
    "eq_ipc_blob_bytes": (C.c_int, []),
    "eq_ipc_export": (C.c_int, [_H, C.c_void_p, C.c_size_t]),
    "eq_ipc_attach": (C.c_int, [_H, C.c_void_p, C.c_size_t, C.c_int]),
}

_cache: dict[str, C.CDLL] = {}


def load(path: str | None = None) -> C.CDLL:
    """Load the C-ABI library and declare every prototype.  Raises if absent."""
    path = os.path.abspath(path or os.environ.get("EQUILIBRIUM_CUDA_LIB") or DEFAULT_LIB)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -m equilibrium_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.eq_abi_version() != 1:
        raise ImportError(f"{path}: ABI version {lib.eq_abi_version()} != 1")
    _cache[path] = lib
    return lib


def check(lib: C.CDLL, code: int) -> None:
    if code != EQ_OK:
        raise EquilibriumError(code, (lib.eq_last_error() or b"").decode("utf-8", "replace"))
