"""Host-side mirror of the reference's CUDA pipeline API over the C ABI (include/sparkl_abi.h).

Names, argument meaning and error behaviour follow src/cuda/*.rs of the reference:
  CudaParticleSet          cuda_particle_set.rs:12-177
  CudaParticleModelSet     cuda_particle_model_set.rs:8-84
  CudaColliderSet          cuda_collider_set.rs:13-131
  CudaSparseGrid           cuda_sparse_grid.rs:26-144
  SingleGpuMpmContext      cuda_mpm_pipeline.rs:177-194
  CudaMpmPipeline.step     cuda_mpm_pipeline.rs:262-271
The reference keeps particles, models, colliders and grid as separate Rust objects borrowed by a
`SingleGpuMpmContext`; here one `spk_ctx` owns all device memory of a context and the classes below are views on it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi
from ._lib import error_string, lib
from .engine import Engine, _ptr


class SingleGpuMpmContext(Engine):
    """One CUDA context: device, 2 streams, particle set, model set, collider set, sparse grid, dt accumulator."""

    def __init__(self, cell_width: float, device: int = 0, block_capacity: int = 1024, dim: int = 3):
        """`dim` selects the library: 3 = libsparkl_b200.so (crate sparkl3d), 2 = libsparkl_b200_2d.so (crate sparkl2d)"""
        l = lib(dim)
        h = C.c_void_p()
        rc = l.spk_ctx_create(device, cell_width, block_capacity, C.byref(h))
        if rc:
            raise abi.SparklError(rc, "spk_ctx_create: " + error_string(rc, dim))
        if dim == 2:
            from . import abi2d as types
        else:
            types = abi
        super().__init__(l, "spk_", h, lambda code: error_string(code, dim), types)
        self.cell_width, self.device, self.dim = cell_width, device, dim
        self.particles = CudaParticleSet(self)
        self.models = CudaParticleModelSet(self)
        self.colliders = CudaColliderSet(self)
        self.grid = CudaSparseGrid(self)

    def close(self):
        if self._h:
            self._lib.spk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        self._call("ctx_synchronize")

    def cuda_stream(self) -> int:
        """raw cudaStream_t of the context's main stream"""
        out = C.c_void_p()
        self._call("ctx_stream", C.byref(out))
        return out.value or 0

    def append(self, pc: abi.ParticleComponents):
        arrs = [np.ascontiguousarray(a) for a in (pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data)]
        n = len(arrs[0])
        if any(len(a) != n for a in arrs):
            raise abi.SparklError(abi.EINVAL, "All attribute buffer must have the same length.")
        self._call("particles_append", *[_ptr(a) for a in arrs], n)

    def remove_range(self, begin: int, end: int):
        self._call("particles_remove_range", begin, end)

    def truncate(self, new_len: int):
        """CudaParticleSet::truncate, cuda_particle_set.rs:81-87"""
        self._call("particles_truncate", new_len)

    def capacity(self) -> int:
        """CudaParticleSet::capacity, cuda_particle_set.rs:77-79"""
        out = C.c_size_t()
        self._call("particles_capacity", C.byref(out))
        return out.value

    def step(self, params: abi.SolverParameters, gravity, timings: bool = True):
        return CudaMpmPipeline().step(params, gravity, [self], timings)

    def upload_async(self, pc: abi.ParticleComponents, stage_only: bool = False):
        """spk_particles_upload_async: `pc`'s arrays must be pinned, contiguous, of the ABI dtypes, and stay alive until the next
        blocking call; the host->device copy overlaps the frame the GPU is still computing"""
        arrs = [pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data]
        T = self.T
        for a, dt in zip(arrs, (T.STATUS, T.POSITION, T.VELOCITY, T.VOLUME, T.PHASE, T.DATA)):
            if a.dtype != dt or not a.flags.c_contiguous or len(a) != len(arrs[0]):
                raise abi.SparklError(abi.EINVAL, "attribute dtype / layout / length mismatch")
        self._call("particles_stage_async" if stage_only else "particles_upload_async", *[_ptr(a) for a in arrs], len(arrs[0]))

    def commit_staged(self):
        """make the state staged by `upload_async(pc, stage_only=True)` the current particle state"""
        self._call("particles_commit_staged")

    def read_async(self, which: int, out: np.ndarray):
        """enqueue the read-back of one attribute into `out` (pinned memory for real overlap) and return at once; the next
        step can be issued immediately.  `read_wait()` blocks until every outstanding read has landed."""
        n = len(self)
        if out.dtype != self.T.ATTR_DTYPES[which] or len(out) < n or not out.flags.c_contiguous:
            raise abi.SparklError(abi.EINVAL, "output buffer has the wrong dtype / length")
        self._call("particles_read_async", which, _ptr(out), n)

    def read_wait(self):
        self._call("particles_read_wait")

    # ---- device-side particle migration (sparkl_b200/distributed.py::migrate_device drives it) ----------
    def set_gids(self, gids):
        """global ids of the local particles, by particle id (default: the identity); reset by upload / append / remove_range"""
        g = np.ascontiguousarray(gids, np.uint32)
        self._call("particles_set_gids", _ptr(g), len(g))

    def read_gids(self) -> np.ndarray:
        out = np.empty(len(self), np.uint32)
        self._call("particles_read_gids", _ptr(out), len(out))
        return out

    def migrate_out(self, bounds, rank: int, dev_records: int = 0, capacity: int = 0) -> np.ndarray:
        """spk_particles_migrate_out: `bounds` = world + 1 increasing block-x coordinates; `dev_records` = DEVICE address of a
        buffer of `capacity` 128-byte rows, or 0 to count only.  Returns the number of rows per destination rank."""
        b = np.ascontiguousarray(bounds, np.int64)
        counts = np.zeros(len(b) - 1, np.uint32)
        self._call("particles_migrate_out", _ptr(b), len(b) - 1, rank, C.c_void_p(dev_records or None), capacity, _ptr(counts))
        return counts

    def migrate_in(self, dev_records: int, n: int):
        """spk_particles_migrate_in: append `n` rows from DEVICE memory"""
        self._call("particles_migrate_in", C.c_void_p(dev_records or None), n)

    def g2p2g_time(self, enable: bool = True):
        """(summed g2p2g kernel time in ms, launches) since the previous call, measured with CUDA events recorded on the
        context's stream without synchronising; `enable` arms / disarms the recording for the launches that follow"""
        ms, n = C.c_float(), C.c_uint32()
        self._call("g2p2g_time", int(enable), C.byref(ms), C.byref(n))
        return ms.value, n.value


class CudaParticleSet:
    def __init__(self, ctx: SingleGpuMpmContext):
        self._ctx = ctx

    @classmethod
    def from_particles(cls, ctx, status, position, velocity, volume, phase, data=None):
        pc = abi.ParticleComponents(0)
        pc.status, pc.position, pc.velocity, pc.volume, pc.phase = status, position, velocity, volume, phase
        pc.data = data if data is not None else _default_data(len(status))
        ctx.upload(pc)
        return ctx.particles

    def len(self):
        return len(self._ctx)

    __len__ = len

    def is_empty(self):
        return len(self._ctx) == 0

    def read_positions(self):
        return self._ctx.read(abi.ATTR_POSITION)

    def read_velocities(self):
        return self._ctx.read(abi.ATTR_VELOCITY)

    def read_statuses(self):
        return self._ctx.read(abi.ATTR_STATUS)

    def read_volumes(self):
        return self._ctx.read(abi.ATTR_VOLUME)

    def read_phase(self):
        return self._ctx.read(abi.ATTR_PHASE)

    def read_sorted_particle_ids(self):
        return self._ctx.read(abi.ATTR_SORTED_IDS)

    def append(self, pc):
        self._ctx.append(pc)

    def remove_range(self, begin, end):
        self._ctx.remove_range(begin, end)

    def truncate(self, new_len):
        self._ctx.truncate(new_len)

    def capacity(self):
        return self._ctx.capacity()


def _default_data(n):
    d = np.zeros(n, abi.DATA)
    d["nacc_alpha"], d["plastic_hardening"], d["elastic_hardening"] = -0.01, 1.0, 1.0
    return d


class CudaParticleModelSet:
    def __init__(self, ctx):
        self._ctx, self.models = ctx, []

    def from_models(self, models):
        self.models = list(models)
        self._ctx.models_upload(self.models)
        return self

    def ensure_models_are_uploaded(self):
        """cuda_particle_model_set.rs:61-84: re-upload if the host models were edited."""
        self._ctx.models_upload(self.models)


class CudaColliderSet:
    def __init__(self, ctx):
        self._ctx, self.colliders = ctx, []

    def from_collider_set(self, colliders):
        self.colliders = list(colliders)
        self._ctx.colliders_upload(self.colliders)
        return self


class CudaSparseGrid:
    def __init__(self, ctx):
        self._ctx = ctx

    def cell_width(self):
        return self._ctx.cell_width

    def curr_headers(self):
        return self._ctx.grid_headers(0)

    def curr_nodes(self):
        return self._ctx.grid_nodes(0)


class CudaMpmPipeline:
    """`CudaMpmPipeline::step`: advances the contexts by params.dt and returns the simulated time."""

    def __init__(self):
        self.step_id = 0

    def step(self, params: abi.SolverParameters, gravity, contexts, timings: bool = True):
        """several contexts = the reference's in-process multi-GPU mode (one host thread drives all devices, shared grid blocks are
        exchanged between the contexts every substep): a compatibility path; the fast one is one process per GPU, sparkl_b200.distributed"""
        if len(contexts) == 0:
            raise abi.SparklError(abi.EINVAL, "no context")
        ctx = contexts[0]
        g = (C.c_float * 3)(*gravity)
        t = abi.StepTimings()
        sim = C.c_float()
        arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
        rc = ctx._lib.spk_step(arr, len(contexts), C.byref(params), C.byref(g), C.byref(t) if timings else None, C.byref(sim))
        if rc:
            raise abi.SparklError(rc, "spk_step: " + error_string(rc, getattr(ctx, "dim", 3)))
        self.step_id += t.num_substeps
        return sim.value, t


def context_from_scene(scene, device: int = 0) -> SingleGpuMpmContext:
    dim = scene.particles.position.dtype["point"].shape[0]  # the scene's particle layout decides which library runs it
    ctx = SingleGpuMpmContext(scene.cell_width, device, dim=dim)
    ctx.models.from_models(scene.models)
    ctx.colliders.from_collider_set(scene.colliders)
    ctx.upload(scene.particles)
    return ctx
