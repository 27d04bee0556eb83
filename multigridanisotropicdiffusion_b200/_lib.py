"""ctypes binding of libmadgpu.so (C-ABI declared in include/madgpu.h).

There is no fallback: if the shared library is missing or no B200 is visible the import / the
first call raises.  Nothing in this package imports the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmadgpu.so")

MADGPU_MAX_STEPS = 64

SMOOTHER_GS, SMOOTHER_WJ = 0, 1
CYCLE_V, CYCLE_FMG, CYCLE_SMOOTHER = 0, 1, 2
PIX_U8, PIX_I16, PIX_F32, PIX_F64 = 0, 1, 2, 3
OK, EINVAL, ECUDA, ENOMEM, ESTATE, ESINGULAR, ENUMERIC = 0, -1, -2, -3, -4, -5, -6

K_NAMES = ["smooth0", "smoothc", "resid0", "restrict", "prolong", "coarse", "misc", "halo", "graph"]


class Params(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("dim", C.c_int32),
        ("size", C.c_int32 * 3),
        ("spacing", C.c_double * 3),
        ("time_step", C.c_double),
        ("number_of_steps", C.c_int32),
        ("cycle", C.c_int32),
        ("iterations_per_grid", C.c_int32),
        ("tolerance", C.c_double),
        ("max_cycles", C.c_int32),
        ("verbose", C.c_int32),
        ("smoother", C.c_int32),
        ("omega", C.c_double),
        ("gs_colors", C.c_int32),
        ("device", C.c_int32),
        ("rank", C.c_int32),
        ("world_size", C.c_int32),
        ("reserved", C.c_int32 * 8),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("struct_size", C.c_int32),
        ("steps", C.c_int32),
        ("cycles_per_step", C.c_int32 * MADGPU_MAX_STEPS),
        ("final_relres", C.c_double * MADGPU_MAX_STEPS),
        ("total_cycles", C.c_int32),
        ("levels", C.c_int32),
        ("setup_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("solve_ms", C.c_double),
        ("fmg_ms", C.c_double),
        ("kernel_launches", C.c_int64),
        ("prof_ms", C.c_double * 16),
        ("prof_launches", C.c_int64 * 16),
        ("graph_launches", C.c_int64),
    ]


class VedParams(C.Structure):  # madved_params, include/madved.h
    _fields_ = [
        ("struct_size", C.c_int32),
        ("size", C.c_int32 * 3),
        ("spacing", C.c_double * 3),
        ("alpha", C.c_double),
        ("beta", C.c_double),
        ("gamma", C.c_double),
        ("epsilon", C.c_double),
        ("omega", C.c_double),
        ("sensitivity", C.c_double),
        ("device", C.c_int32),
        ("reserved", C.c_int32 * 7),
    ]


class VedStats(C.Structure):  # madved_stats
    _fields_ = [
        ("struct_size", C.c_int32),
        ("scales", C.c_int32),
        ("hessian_ms", C.c_double),
        ("vesselness_ms", C.c_double),
        ("h2d_ms", C.c_double),
        ("d2h_ms", C.c_double),
        ("diffusion_ms", C.c_double),
        ("kernel_launches", C.c_int64),
    ]


EXPORTS = [
    "madgpu_params_default", "madgpu_create", "madgpu_create_slab", "madgpu_nccl_unique_id", "madgpu_slab", "madgpu_ipc_export", "madgpu_ipc_import", "madgpu_ipc_disable", "madgpu_destroy", "madgpu_last_error", "madgpu_set_solver",
    "madgpu_set_tensor_f32", "madgpu_set_tensor_f64", "madgpu_set_tensor_device_f32", "madgpu_solve_cast",
    "madgpu_solve_u8", "madgpu_solve_i16", "madgpu_solve_f32", "madgpu_solve_f64", "madgpu_solve_device_f32",
    "madgpu_cycles_begin_device_f32", "madgpu_cycles_begin_f32", "madgpu_cycles_run",
    "madgpu_cycles_end_device_f32", "madgpu_cycles_end_f64", "madgpu_get_relres_history", "madgpu_set_profiling", "madgpu_num_levels", "madgpu_level_info", "madgpu_gs_tile", "madgpu_gs_leg_plan",
    "madgpu_op_get_tensor", "madgpu_op_assemble", "madgpu_op_smooth", "madgpu_op_residual",
    "madgpu_op_residual_f64", "madgpu_op_restrict", "madgpu_op_prolong", "madgpu_op_coarse_solve",
    "madgpu_op_vcycle", "madgpu_fetch_output",
    # include/madved.h
    "madved_params_default", "madved_create", "madved_destroy", "madved_last_error", "madved_set_params", "madved_set_image",
    "madved_set_image_device_f32", "madved_image_device", "madved_begin", "madved_hessian", "madved_update_vesselness",
    "madved_update_vesselness_host_f64", "madved_add_scale", "madved_tensor_planes", "madved_get_tensor_f64",
    "madved_get_response_f64", "madved_get_hessian_f64", "madved_get_stats", "madved_run",
]

_lib = None


def load() -> C.CDLL:
    """Load libmadgpu.so; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C multigridanisotropicdiffusion_b200/csrc` (there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, f64 = C.c_void_p, C.c_int32, C.c_double
    L.madgpu_params_default.argtypes = [C.POINTER(Params)]
    L.madgpu_params_default.restype = None
    L.madgpu_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.madgpu_destroy.argtypes = [vp]
    L.madgpu_destroy.restype = None
    L.madgpu_last_error.argtypes = [vp]
    L.madgpu_last_error.restype = C.c_char_p
    L.madgpu_set_solver.argtypes = [vp, i32, f64, i32, i32, f64, i32, i32, i32]
    L.madgpu_set_tensor_f32.argtypes = [vp, vp]
    L.madgpu_set_tensor_f64.argtypes = [vp, vp]
    L.madgpu_set_tensor_device_f32.argtypes = [vp, C.POINTER(vp)]
    L.madgpu_solve_cast.argtypes = [vp, i32, vp, i32, vp, C.POINTER(Stats)]
    for n in ("u8", "i16", "f32", "f64", "device_f32"):
        getattr(L, "madgpu_solve_" + n).argtypes = [vp, vp, vp, C.POINTER(Stats)]
    L.madgpu_cycles_begin_device_f32.argtypes = [vp, vp]
    L.madgpu_cycles_begin_f32.argtypes = [vp, vp]
    L.madgpu_cycles_run.argtypes = [vp, i32, C.POINTER(f64), C.POINTER(C.c_float), C.POINTER(Stats)]
    L.madgpu_cycles_end_device_f32.argtypes = [vp, vp]
    L.madgpu_cycles_end_f64.argtypes = [vp, vp]
    L.madgpu_get_relres_history.argtypes = [vp, C.POINTER(f64), i32]
    L.madgpu_set_profiling.argtypes = [vp, i32]
    L.madgpu_gs_tile.argtypes = [vp, i32, C.POINTER(i32)]
    L.madgpu_gs_leg_plan.argtypes = [vp, i32, i32, C.POINTER(i32), i32]
    L.madgpu_create_slab.argtypes = [C.POINTER(Params), C.c_char_p, C.POINTER(vp)]
    L.madgpu_nccl_unique_id.argtypes = [C.c_char_p]
    L.madgpu_slab.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    L.madgpu_ipc_export.argtypes = [vp, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
    L.madgpu_ipc_import.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.madgpu_ipc_disable.argtypes = [vp]
    L.madgpu_num_levels.argtypes = [vp]
    L.madgpu_level_info.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(f64), C.POINTER(i32)]
    L.madgpu_op_get_tensor.argtypes = [vp, i32, vp]
    L.madgpu_op_assemble.argtypes = [vp, i32, vp]
    L.madgpu_op_smooth.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.madgpu_op_residual.argtypes = [vp, i32, vp, vp, vp, C.POINTER(f64)]
    L.madgpu_op_residual_f64.argtypes = [vp, vp, vp, vp, C.POINTER(f64)]
    L.madgpu_op_restrict.argtypes = [vp, i32, vp, vp]
    L.madgpu_op_prolong.argtypes = [vp, i32, vp, vp]
    L.madgpu_op_coarse_solve.argtypes = [vp, vp, vp]
    L.madgpu_op_vcycle.argtypes = [vp, i32, vp, vp, vp]
    L.madgpu_fetch_output.argtypes = [vp, i32, vp]
    L.madved_params_default.argtypes = [C.POINTER(VedParams)]
    L.madved_params_default.restype = None
    L.madved_create.argtypes = [C.POINTER(VedParams), C.POINTER(vp)]
    L.madved_destroy.argtypes = [vp]
    L.madved_destroy.restype = None
    L.madved_last_error.argtypes = [vp]
    L.madved_last_error.restype = C.c_char_p
    L.madved_set_params.argtypes = [vp, f64, f64, f64, f64, f64, f64]
    L.madved_set_image.argtypes = [vp, i32, vp]
    L.madved_set_image_device_f32.argtypes = [vp, vp]
    L.madved_image_device.argtypes = [vp, C.POINTER(vp)]
    L.madved_begin.argtypes = [vp]
    L.madved_hessian.argtypes = [vp, f64]
    L.madved_update_vesselness.argtypes = [vp]
    L.madved_update_vesselness_host_f64.argtypes = [vp, vp]
    L.madved_add_scale.argtypes = [vp, f64]
    L.madved_tensor_planes.argtypes = [vp, C.POINTER(vp)]
    L.madved_get_tensor_f64.argtypes = [vp, vp]
    L.madved_get_response_f64.argtypes = [vp, vp]
    L.madved_get_hessian_f64.argtypes = [vp, vp]
    L.madved_get_stats.argtypes = [vp, C.POINTER(VedStats)]
    L.madved_run.argtypes = [vp, vp, i32, vp, i32, vp, C.POINTER(f64), i32, i32, C.POINTER(Stats)]
    _lib = L
    return L
