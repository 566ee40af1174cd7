"""ctypes plumbing shared by the product library (prefix `spk_`) and the test oracle (prefix `orc_`).

Both expose the same stage-level C entry points (include/sparkl_abi.h); `Engine` turns them into numpy-level
calls.  Nothing here computes: it only marshals arrays across the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

_vp = C.c_void_p


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(_vp)


def declare(lib, prefix: str):
    """Declare argtypes/restype for every entry point the Engine uses."""
    sig = {
        "particles_upload": [_vp] * 7 + [C.c_size_t],
        "particles_len": [_vp, C.POINTER(C.c_size_t)],
        "particles_read": [_vp, C.c_int, _vp, C.c_size_t],
        "models_upload": [_vp, _vp, C.c_size_t],
        "colliders_upload": [_vp, _vp, C.c_size_t],
        "frame_begin": [_vp, C.POINTER(abi.SolverParameters)],
        "stage_sort": [_vp, C.POINTER(abi.SortInfo)],
        "stage_reset_grid": [_vp],
        "stage_estimate_dt": [_vp, C.c_float, C.c_float, C.POINTER(C.c_float)],
        "dt_acc_get": [_vp, C.POINTER(C.c_uint64)],
        "dt_acc_set": [_vp, C.c_uint64],
        "stage_g2p2g": [_vp, C.c_float, C.c_uint32],
        "stage_grid_update": [_vp, C.c_float, C.POINTER(C.c_float * 3)],
        "grid_num_active_blocks": [_vp, C.c_int, C.POINTER(C.c_uint32)],
        "grid_read_headers": [_vp, C.c_int, _vp, C.c_uint32],
        "grid_read_nodes": [_vp, C.c_int, _vp, C.c_uint32],
        "grid_read_dispatch": [_vp, _vp, C.c_uint32],
    }
    # not part of the reference's compiled module (absent from the `ref_` harness) / a property of this library's sort only (absent from the oracle too)
    optional = {"stage_recompute_fluids_volume": [_vp], "grid_morton_box": [_vp, _vp]}
    for name, args in list(sig.items()) + [(k, v) for k, v in optional.items() if hasattr(lib, prefix + k)]:
        fn = getattr(lib, prefix + name)
        fn.argtypes = args
        fn.restype = C.c_int


class Engine:
    """One simulation context behind a C ABI (`spk_ctx*` or `orc_ctx*`).  `types` is the module holding the byte images of the
    build's dimension-dependent structs: abi (3D, default) or abi2d (the -DSPK_DIM=2 libraries)."""

    def __init__(self, lib, prefix: str, handle, error_string=None, types=abi):
        self._lib, self._prefix, self._h = lib, prefix, handle
        self._err = error_string
        self.T = types

    def _call(self, name, *args):
        rc = getattr(self._lib, self._prefix + name)(self._h, *args)
        if rc != 0:
            what = self._err(rc) if self._err else name
            raise abi.SparklError(rc, f"{self._prefix}{name}: {what}")

    # ---- particles ------------------------------------------------------------------------------------------
    def upload(self, pc: abi.ParticleComponents):
        T = self.T
        arrs = [np.ascontiguousarray(a) for a in (pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data)]
        for a, dt in zip(arrs, (T.STATUS, T.POSITION, T.VELOCITY, T.VOLUME, T.PHASE, T.DATA)):
            if a.dtype != dt:
                raise abi.SparklError(abi.EINVAL, "attribute dtype mismatch")
        n = len(arrs[0])
        if any(len(a) != n for a in arrs):  # the reference asserts equal lengths (cuda_particle_set.rs:40-59)
            raise abi.SparklError(abi.EINVAL, "All attribute buffer must have the same length.")
        self._call("particles_upload", *[_ptr(a) for a in arrs], n)

    def __len__(self):
        n = C.c_size_t()
        self._call("particles_len", C.byref(n))
        return n.value

    def read(self, which: int, n: int | None = None, out: np.ndarray | None = None) -> np.ndarray:
        """read one attribute back in particle-id order; `out` may be a caller-owned (e.g. pinned) buffer of the attribute's
        dtype, which makes the device->host copy run at full PCIe rate (the testbed's frame read-back, testbed_plugin.rs:533-588)"""
        n = len(self) if n is None else n
        if out is None:
            out = np.empty(n, self.T.ATTR_DTYPES[which])
        elif out.dtype != self.T.ATTR_DTYPES[which] or len(out) < n or not out.flags.c_contiguous:
            raise abi.SparklError(abi.EINVAL, "output buffer has the wrong dtype / length")
        if n:
            self._call("particles_read", which, _ptr(out), n)
        return out

    def read_all(self) -> abi.ParticleComponents:
        pc = self.T.ParticleComponents(0)
        pc.position, pc.velocity = self.read(abi.ATTR_POSITION), self.read(abi.ATTR_VELOCITY)
        pc.status, pc.volume = self.read(abi.ATTR_STATUS), self.read(abi.ATTR_VOLUME)
        pc.phase, pc.data = self.read(abi.ATTR_PHASE), self.read(abi.ATTR_DATA)
        return pc

    def models_upload(self, models):
        packed = abi.pack_models(models)
        self._call("models_upload", _ptr(packed), len(models))

    def colliders_upload(self, colliders):
        packed = self.T.pack_colliders(colliders)
        self._call("colliders_upload", _ptr(packed), len(colliders))

    # ---- stages ---------------------------------------------------------------------------------------------
    def frame_begin(self, params: abi.SolverParameters):
        self._call("frame_begin", C.byref(params))

    def stage_sort(self) -> abi.SortInfo:
        info = abi.SortInfo()
        self._call("stage_sort", C.byref(info))
        return info

    def stage_reset_grid(self):
        self._call("stage_reset_grid")

    def stage_estimate_dt(self, min_dt: float, max_dt: float) -> float:
        out = C.c_float()
        self._call("stage_estimate_dt", min_dt, max_dt, C.byref(out))
        return out.value

    def dt_acc(self) -> int:
        out = C.c_uint64()
        self._call("dt_acc_get", C.byref(out))
        return out.value

    def set_dt_acc(self, v: int):
        self._call("dt_acc_set", v)

    def stage_g2p2g(self, dt: float, damage_model: int = abi.DamageModel.NONE):
        self._call("stage_g2p2g", dt, damage_model)

    def stage_grid_update(self, dt: float, gravity):
        g = (C.c_float * 3)(*gravity)
        self._call("stage_grid_update", dt, C.byref(g))

    def stage_recompute_fluids_volume(self):
        """fluids-volume recomputation (src/dynamics/solver/fluids_volume.rs:9-116, src_kernels/cuda/fluids_volume.rs): between
        stage_reset_grid and stage_g2p2g"""
        self._call("stage_recompute_fluids_volume")

    def substep(self, params: abi.SolverParameters, gravity, remaining: float) -> float:
        """One iteration of the `while remaining_time > 0` loop of cuda_mpm_pipeline.rs:322-638."""
        min_dt = np.float32(params.dt) / np.float32(params.max_num_substeps)
        info = self.stage_sort()
        self.stage_reset_grid()
        dt = self.stage_estimate_dt(float(min_dt), float(min(np.float32(remaining), np.float32(params.max_substep_dt))))
        if info.num_dispatch_blocks > 0:
            self.stage_g2p2g(dt, params.damage_model)
        self.stage_grid_update(dt, gravity)
        return dt

    # ---- grid read-back -------------------------------------------------------------------------------------
    def num_active_blocks(self, which_grid: int) -> int:
        out = C.c_uint32()
        self._call("grid_num_active_blocks", which_grid, C.byref(out))
        return out.value

    def grid_headers(self, which_grid: int) -> np.ndarray:
        n = self.num_active_blocks(which_grid)
        out = np.zeros(n, abi.ACTIVE_BLOCK_HEADER)
        if n:
            self._call("grid_read_headers", which_grid, _ptr(out), n)
        return out

    def grid_nodes(self, which_grid: int) -> np.ndarray:
        n = self.num_active_blocks(which_grid)
        out = np.zeros((n, getattr(self.T, "CELLS_PER_BLOCK", 64)), self.T.GRID_NODE)
        if n:
            self._call("grid_read_nodes", which_grid, _ptr(out), n)
        return out

    def morton_box(self) -> np.ndarray:
        """spk_grid_morton_box: origin[3], bits[3], had-a-box, a-block-fell-outside of the last sort"""
        out = np.zeros(8, np.uint32)
        self._call("grid_morton_box", _ptr(out))
        return out

    def dispatch(self, n: int) -> np.ndarray:
        out = np.zeros(n, abi.DISPATCH)
        if n:
            self._call("grid_read_dispatch", _ptr(out), n)
        return out
