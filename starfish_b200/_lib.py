"""ctypes binding of ``libstarfish_gpu.so`` (C ABI in ``include/sfgpu.h``).

Fails loudly: a missing library is an ImportError-like RuntimeError naming the build command, never a
silent fallback.  ``EXPORTS`` lists every symbol the header declares; tests check the built library
exports them all.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFGPU_LIB_PATH") or os.path.join(_HERE, "libstarfish_gpu.so")  # env: kernel-variant experiments

NFIELDS = 8
FIELD_NAMES = ("den", "u", "v", "w", "uu", "vv", "ww", "mpc")

INJECT_REWIND, INJECT_DEPOSIT_NOW, INJECT_TRANSFER = 1, 2, 4
STEP_GENERIC, STEP_DEFER_FINISH, STEP_INPLACE, STEP_STREAM = 1, 2, 4, 8
SOURCE_COLD_BEAM = 1

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Particles(C.Structure):
    """``struct sfgpu_particles``"""
    _fields_ = [("n", C.c_int64)] + [(k, c_double_p) for k in ("x", "y", "z", "u", "v", "w", "mpw", "li", "lj", "dt")] + [
        ("id", c_int32_p), ("born_it", c_int32_p)]


class SlowExtra(C.Structure):
    """``struct sfgpu_slow_extra``"""
    _fields_ = [(k, c_double_p) for k in ("old_x", "old_y", "old_li", "old_lj")] + [("bounces", c_int32_p), ("mesh", c_int32_p)]


class Spline(C.Structure):
    """``struct sfgpu_spline``"""
    _fields_ = [("n_seg", C.c_int32)] + [(k, c_double_p) for k in ("x1", "y1", "x2", "y2", "nx", "ny", "area", "cum_area")] + [("spline_area", C.c_double)]


# name -> (restype, argtypes); must list every function declared in include/sfgpu.h
EXPORTS = {
    "sfgpu_abi_version": (C.c_int, []),
    "sfgpu_create": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "sfgpu_destroy": (None, [C.c_void_p]),
    "sfgpu_last_error": (C.c_char_p, [C.c_void_p]),
    "sfgpu_mesh_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, C.POINTER(C.c_void_p),
                                 C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, c_int32_p]),
    "sfgpu_mesh_set_segments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "sfgpu_take_surface_hits": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, c_int32_p, c_int32_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                          C.POINTER(C.c_int8), c_int64_p, c_int64_p]),
    "sfgpu_multi_create": (C.c_int, [C.c_int32, c_int32_p, C.c_int, C.POINTER(C.c_void_p)]),
    "sfgpu_multi_destroy": (None, [C.c_void_p]),
    "sfgpu_multi_size": (C.c_int32, [C.c_void_p]),
    "sfgpu_multi_ctx": (C.c_void_p, [C.c_void_p, C.c_int32]),
    "sfgpu_multi_last_error": (C.c_char_p, [C.c_void_p]),
    "sfgpu_multi_mesh_add": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, c_int32_p]),
    "sfgpu_multi_mesh_set_segments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p]),
    "sfgpu_multi_set_fields": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfgpu_multi_species_add": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, c_int32_p]),
    "sfgpu_multi_inject": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Particles), C.c_double, C.c_uint32, c_int64_p]),
    "sfgpu_multi_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_uint32]),
    "sfgpu_multi_finish_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_multi_get_moments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfgpu_multi_get_deposit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "sfgpu_multi_get_samples": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), c_int64_p]),
    "sfgpu_multi_clear_samples": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_multi_get_sums": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_int64_p, c_int64_p, c_int64_p]),
    "sfgpu_cell_lists": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int64_p, c_int32_p, c_int64_p]),
    "sfgpu_set_fields": (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfgpu_species_add": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int64, c_int32_p]),
    "sfgpu_inject": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(Particles), C.c_double, C.c_uint32, c_int64_p]),
    "sfgpu_source_uniform": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(Spline), C.c_uint32, C.c_double, C.c_double, C.c_int32, C.c_int64, C.c_double,
                                      C.POINTER(C.c_uint64), c_int64_p]),
    "sfgpu_step": (C.c_int, [C.c_void_p, C.c_int32, C.c_double, C.c_uint32]),
    "sfgpu_finish_step": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_get_deposit": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    "sfgpu_get_samples": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), c_int64_p]),
    "sfgpu_clear_samples": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "sfgpu_host_free": (None, [C.c_void_p]),
    "sfgpu_get_moments": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sfgpu_get_sums": (C.c_int, [C.c_void_p, C.c_int32, c_double_p, c_int64_p, c_int64_p, c_int64_p]),
    "sfgpu_np": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, c_int64_p]),
    "sfgpu_take_slowpath": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.POINTER(Particles), C.POINTER(SlowExtra), c_int64_p]),
    "sfgpu_download": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(Particles)]),
    "sfgpu_upload": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.POINTER(Particles)]),
    "sfgpu_restart_save": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, c_int64_p]),
    "sfgpu_restart_load": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int64, C.c_double, c_int64_p, c_int64_p]),
    "sfgpu_sort": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_set_sort_interval": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_set_tile_halo": (C.c_int, [C.c_void_p, C.c_int32]),
    "sfgpu_get_tile_halo": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]),
    "sfgpu_comm_unique_id": (C.c_int, [C.c_void_p]),
    "sfgpu_comm_init": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
    "sfgpu_deposit_device_ptr": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), c_int64_p]),
    "sfgpu_last_step_timing": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), c_int32_p]),
    "sfgpu_last_step_counters": (C.c_int, [C.c_void_p, c_int64_p]),
    "sfgpu_last_step_kernel": (C.c_int, [C.c_void_p, c_int32_p]),
    "sfgpu_sync": (C.c_int, [C.c_void_p]),
    "sfgpu_timer_start": (C.c_int, [C.c_void_p]),
    "sfgpu_timer_stop": (C.c_int, [C.c_void_p, C.POINTER(C.c_float)]),
    "sfgpu_launch_count": (C.c_int, [C.c_void_p, c_int64_p]),
}

_lib = None


def load():
    """Load libstarfish_gpu.so once and declare every prototype."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C starfish_b200/csrc`). starfish_b200 has no CPU fallback.")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
