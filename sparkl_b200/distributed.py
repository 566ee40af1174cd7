"""Multi-GPU MPM substep: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).

Semantics are the reference's multi-GPU scheme (src/cuda/cuda_sparse_grid.rs:256-322,
src/cuda/cuda_mpm_pipeline.rs:393-533): every rank owns a fixed subset of the particles; after the local sort the ranks
exchange their active-block keys, blocks active on more than one rank are tagged as halo, after G2P2G the accumulated
node sums of halo blocks are exchanged and added on every rank that holds the block, the grid update is then redundant
on shared blocks, and dt is the minimum over ranks.  Because a shared block ends with the full sum on every rank that
activated it, the result does not depend on how particles are partitioned — the bench partitions the domain into
x-slabs so that only slab faces are shared.

Differences with the reference:
  * payload = the 6 summed floats per node (1536 B per block instead of the 4616-byte `HaloBlockData`);
  * fixed-capacity buffers padded with the EMPTY key are all-gathered, so no lengths travel and the exchange needs no
    host round trip; the reference reads counts back and issues one `cuMemcpyDtoDAsync` per pair of devices;
  * ONE small all-reduce per substep carries the global dt minimum and the global halo-block maximum (which sizes the
    payload exchange identically on every rank); the reference loops over devices on the host (:393-398);
  * everything is enqueued on the context's own stream (torch ops run under `ExternalStream(ctx stream)`), so the
    collectives are stream-ordered with the kernels: 3 host synchronisations per substep in total
    (sort counters, halo count, all-reduce result).
"""
from __future__ import annotations

import contextlib
import ctypes as C

import numpy as np

from . import abi

BLOCK_PAYLOAD_FLOATS = 6 * 64
EMPTY = -1  # u64::MAX viewed as int64


class DistributedMpmPipeline:
    """`CudaMpmPipeline::step` over several single-GPU contexts living in different processes."""

    def __init__(self, ctx, dist=None, device=None, key_capacity: int = 1 << 17):
        import torch
        import torch.distributed as tdist
        self.torch, self.ctx = torch, ctx
        self.dist = dist or tdist
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        self.world, self.rank = self.dist.get_world_size(), self.dist.get_rank()
        self.cuda = self.device.type == "cuda"
        self.stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=self.device) if self.cuda else None
        self.key_cap = int(key_capacity)  # agreed capacity of the active-key exchange (blocks per rank)
        z = lambda *shape, dt: torch.zeros(shape, dtype=dt, device=self.device)
        self._keys = z(self.key_cap, dt=torch.int64)
        self._allkeys = z(self.world, self.key_cap, dt=torch.int64)
        self._halo_cap = 0
        self._hk = self._hp = self._allhk = self._allhp = None
        self._red = z(2, dt=torch.int64)
        self._red_host = torch.zeros(2, dtype=torch.int64).pin_memory() if self.cuda else torch.zeros(2, dtype=torch.int64)
        self.halo_blocks_last = 0
        self.halo_bytes_last = 0

    # ---- helpers ------------------------------------------------------------------------------------------------
    def _scope(self):
        return self.torch.cuda.stream(self.stream) if self.cuda else contextlib.nullcontext()

    def _all_gather(self, out2d, inp):
        if self.cuda:
            self.dist.all_gather_into_tensor(out2d.view(-1, *inp.shape[1:]) if inp.dim() > 1 else out2d.view(-1), inp)
        else:
            lst = [out2d[r] for r in range(self.world)]
            self.dist.all_gather(lst, inp)

    def _ensure_halo(self, cap):
        if cap == self._halo_cap:
            return
        torch = self.torch
        self._halo_cap = cap
        self._hk = torch.zeros(cap, dtype=torch.int64, device=self.device)
        self._hp = torch.zeros((cap, BLOCK_PAYLOAD_FLOATS), dtype=torch.float32, device=self.device)
        self._allhk = torch.zeros((self.world, cap), dtype=torch.int64, device=self.device)
        self._allhp = torch.zeros((self.world, cap, BLOCK_PAYLOAD_FLOATS), dtype=torch.float32, device=self.device)

    def _others(self, t2d):
        """the two contiguous row ranges of a [world, ...] buffer that belong to the OTHER ranks"""
        out = []
        if self.rank > 0:
            out.append(t2d[: self.rank])
        if self.rank + 1 < self.world:
            out.append(t2d[self.rank + 1:])
        return out

    # ---- one substep --------------------------------------------------------------------------------------------
    def substep(self, params: abi.SolverParameters, gravity, remaining: float) -> float:
        ctx, vp = self.ctx, C.c_void_p
        with self._scope():
            info = ctx.stage_sort()                                            # host sync 1 (counters)
            if info.num_active_blocks > self.key_cap:
                raise abi.SparklError(abi.EOVERFLOW, f"{info.num_active_blocks} active blocks exceed the key exchange capacity {self.key_cap}")
            # ---- halo tagging: all-gather of active block keys (cuda_sparse_grid.rs:256-322)
            ctx._call("halo_begin")
            ctx._call("halo_active_keys", vp(self._keys.data_ptr()), self.key_cap)
            self._all_gather(self._allkeys, self._keys)
            for part in self._others(self._allkeys):
                ctx._call("halo_tag", vp(part.data_ptr()), int(part.numel()))
            ctx.stage_reset_grid()
            # ---- dt: local candidate, then ONE all-reduce for (global dt min, global max halo count)
            min_dt = float(np.float32(params.dt) / np.float32(params.max_num_substeps))
            max_dt = float(min(np.float32(remaining), np.float32(params.max_substep_dt)))
            ctx.stage_estimate_dt(min_dt, max_dt)
            nh = C.c_uint32()
            ctx._call("halo_count", C.byref(nh))                               # host sync 2
            self._red_host[0] = min(ctx.dt_acc(), (1 << 63) - 1)
            self._red_host[1] = -int(nh.value)
            self._red.copy_(self._red_host, non_blocking=True)
            self.dist.all_reduce(self._red, op=self.dist.ReduceOp.MIN)
            red = self._red.cpu() if self.cuda else self._red                  # host sync 3
            acc, max_halo = int(red[0]), -int(red[1])
            ctx.set_dt_acc(acc)
            dt = float(np.float32(acc) / np.float32(1.0e12))
            # ---- transfers, then halo accumulate (cuda_mpm_pipeline.rs:407-533)
            if info.num_dispatch_blocks > 0:
                ctx.stage_g2p2g(dt, params.damage_model)
            if max_halo > 0:
                cap = (max_halo + 255) // 256 * 256
                self._ensure_halo(cap)
                hk, hp = self._hk[:cap], self._hp[:cap]
                ctx._call("halo_pack", vp(hk.data_ptr()), vp(hp.data_ptr()), cap)
                allhk, allhp = self._allhk, self._allhp
                self._all_gather(allhk, hk)
                self._all_gather(allhp, hp)
                for pk, pp in zip(self._others(allhk), self._others(allhp)):
                    ctx._call("halo_merge", vp(pk.data_ptr()), vp(pp.data_ptr()), int(pk.numel()))
            self.halo_blocks_last = int(nh.value)
            self.halo_bytes_last = int(nh.value) * (8 + 4 * BLOCK_PAYLOAD_FLOATS)
            ctx.stage_grid_update(dt, gravity)
        return dt

    def step(self, params: abi.SolverParameters, gravity):
        """one frame: returns (simulated_time, number of substeps)"""
        self.ctx.frame_begin(params)
        remaining = np.float32(params.dt)
        n = 0
        if params.dt == 0.0:
            return float(params.dt), 0
        while remaining > 0:
            dt = self.substep(params, gravity, float(remaining))
            n += 1
            remaining = np.float32(remaining - np.float32(dt))
            if remaining <= np.float32(10.0) / np.float32(1.0e12) or params.stop_after_one_substep:
                break
        return float(np.float32(params.dt) - remaining), n


class NvlinkMpmPipeline:
    """The product multi-GPU path: halo accumulation and the dt minimum run inside `spk_step` with device-initiated NVLink
    loads / atomics on peer memory (sparkl_b200/csrc/spk_dist.cu).  torch.distributed is used ONCE, to all-gather the
    128-byte IPC descriptors; no collective runs in the substep loop.  Every rank must own particles and make the same
    sequence of step calls."""

    def __init__(self, ctx, dist=None):
        import torch.distributed as tdist
        self.ctx, self.dist = ctx, dist or tdist
        self.world, self.rank = self.dist.get_world_size(), self.dist.get_rank()
        desc = (C.c_ubyte * 128)()
        ctx._call("dist_export", C.cast(desc, C.c_void_p))
        mine = bytes(desc)
        every = [None] * self.world
        self.dist.all_gather_object(every, mine)
        blob = (C.c_ubyte * (128 * self.world)).from_buffer_copy(b"".join(every))
        ctx._call("dist_attach", self.rank, self.world, C.cast(blob, C.c_void_p))
        self.dist.barrier()  # every rank has mapped every arena before the first device-side barrier

    def step(self, params: abi.SolverParameters, gravity, timings: bool = True):
        """one frame; returns (simulated_time, StepTimings) like SingleGpuMpmContext.step"""
        return self.ctx.step(params, gravity, timings)

    def halo_blocks(self) -> int:
        n, e = C.c_uint32(), C.c_uint32()
        self.ctx._call("dist_info", C.byref(n), C.byref(e))
        return n.value

    def close(self):
        self.ctx._call("dist_detach")
