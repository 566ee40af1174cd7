"""Loads libsparkl_b200.so (the CUDA product library).  There is no CPU fallback: if the library is missing or no
CUDA device is usable, every entry point raises."""
from __future__ import annotations

import ctypes as C
import os

from . import abi
from .engine import declare

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPK_LIB_PATH") or os.path.join(_HERE, "libsparkl_b200.so")  # SPK_LIB_PATH: A/B builds of the same ABI (scripts/variants.sh)
LIB2D_PATH = os.environ.get("SPK_LIB2D_PATH") or os.path.join(_HERE, "libsparkl_b200_2d.so")  # the -DSPK_DIM=2 build (same entry points)
_LIB = None
_LIBS = {}

# every symbol include/sparkl_abi.h declares
EXPORTS = [
    "spk_abi_version", "spk_dim", "spk_error_string", "spk_ctx_create", "spk_ctx_destroy", "spk_ctx_cell_width", "spk_ctx_synchronize", "spk_ctx_stream",
    "spk_particles_upload", "spk_particles_append", "spk_particles_remove_range", "spk_particles_len", "spk_particles_read",
    "spk_particles_read_async", "spk_particles_read_wait", "spk_particles_upload_async", "spk_particles_stage_async", "spk_particles_commit_staged",
    "spk_models_upload", "spk_colliders_upload", "spk_step", "spk_frame_begin", "spk_stage_sort", "spk_stage_reset_grid",
    "spk_stage_estimate_dt", "spk_dt_acc_get", "spk_dt_acc_set", "spk_stage_g2p2g", "spk_stage_grid_update", "spk_stage_recompute_fluids_volume",
    "spk_halo_begin", "spk_halo_active_keys", "spk_halo_tag", "spk_halo_count", "spk_halo_pack", "spk_halo_merge", "spk_grid_num_active_blocks",
    "spk_grid_read_headers", "spk_grid_read_nodes", "spk_grid_read_dispatch", "spk_grid_morton_box", "spk_device_ptr", "spk_g2p2g_time",
    "spk_dist_export", "spk_dist_attach", "spk_dist_detach", "spk_dist_info",
    "spk_particles_truncate", "spk_particles_capacity", "spk_particles_set_gids", "spk_particles_read_gids", "spk_particles_migrate_out", "spk_particles_migrate_in",
    "spk_set_particle_kernels_launcher", "spk_default_launch_estimate_particle_timestep_length", "spk_default_launch_g2p2g",
]


def lib(dim: int = 3):
    """the product library of the given dimension (3: libsparkl_b200.so, 2: libsparkl_b200_2d.so — one source tree, two builds,
    like the reference's sparkl3d / sparkl2d crates)"""
    global _LIB
    if dim not in (2, 3):
        raise abi.SparklError(abi.EINVAL, "dim must be 2 or 3")
    if dim not in _LIBS:
        path = LIB_PATH if dim == 3 else LIB2D_PATH
        if not os.path.exists(path):
            raise abi.SparklError(abi.ESTATE, f"{path} is not built (run `python -m sparkl_b200.build`); there is no CPU fallback")
        l = C.CDLL(path)
        declare(l, "spk_")
        vp = C.c_void_p
        l.spk_abi_version.restype = C.c_int
        l.spk_error_string.argtypes = [C.c_int]
        l.spk_error_string.restype = C.c_char_p
        l.spk_ctx_create.argtypes = [C.c_int, C.c_float, C.c_uint32, C.POINTER(vp)]
        l.spk_ctx_destroy.argtypes = [vp]
        l.spk_ctx_synchronize.argtypes = [vp]
        l.spk_ctx_stream.argtypes = [vp, C.POINTER(vp)]
        l.spk_particles_append.argtypes = [vp] * 7 + [C.c_size_t]
        l.spk_particles_remove_range.argtypes = [vp, C.c_size_t, C.c_size_t]
        l.spk_particles_read_async.argtypes = [vp, C.c_int, vp, C.c_size_t]
        l.spk_particles_read_wait.argtypes = [vp]
        l.spk_particles_upload_async.argtypes = [vp] * 7 + [C.c_size_t]
        l.spk_particles_stage_async.argtypes = [vp] * 7 + [C.c_size_t]
        l.spk_particles_commit_staged.argtypes = [vp]
        l.spk_step.argtypes = [C.POINTER(vp), C.c_size_t, C.POINTER(abi.SolverParameters), C.POINTER(C.c_float * 3),
                               C.POINTER(abi.StepTimings), C.POINTER(C.c_float)]
        l.spk_halo_begin.argtypes = [vp]
        l.spk_halo_active_keys.argtypes = [vp, vp, C.c_uint32]
        l.spk_halo_tag.argtypes = [vp, vp, C.c_uint32]
        l.spk_halo_count.argtypes = [vp, C.POINTER(C.c_uint32)]
        l.spk_halo_pack.argtypes = [vp, vp, vp, C.c_uint32]
        l.spk_halo_merge.argtypes = [vp, vp, vp, C.c_uint32]
        l.spk_device_ptr.argtypes = [vp, C.c_int, C.POINTER(vp)]
        l.spk_g2p2g_time.argtypes = [vp, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_uint32)]
        l.spk_dist_export.argtypes = [vp, vp]
        l.spk_dist_attach.argtypes = [vp, C.c_uint32, C.c_uint32, vp]
        l.spk_dist_detach.argtypes = [vp]
        l.spk_dist_info.argtypes = [vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        l.spk_set_particle_kernels_launcher.argtypes = [vp, vp]
        l.spk_default_launch_estimate_particle_timestep_length.argtypes = [vp, C.c_float, C.c_float]
        l.spk_default_launch_g2p2g.argtypes = [vp, C.c_float]
        l.spk_particles_truncate.argtypes = [vp, C.c_size_t]
        l.spk_particles_capacity.argtypes = [vp, C.POINTER(C.c_size_t)]
        l.spk_particles_set_gids.argtypes = [vp, vp, C.c_size_t]
        l.spk_particles_read_gids.argtypes = [vp, vp, C.c_size_t]
        l.spk_particles_migrate_out.argtypes = [vp, vp, C.c_uint32, C.c_uint32, vp, C.c_size_t, vp]
        l.spk_particles_migrate_in.argtypes = [vp, vp, C.c_size_t]
        for name in EXPORTS:
            if name not in ("spk_error_string",):
                getattr(l, name).restype = C.c_int
        _LIBS[dim] = l
        if dim == 3:
            _LIB = l
    return _LIBS[dim]


def error_string(code: int, dim: int = 3) -> str:
    return lib(dim).spk_error_string(code).decode()
