"""Multi-GPU MPM substep: one process per GPU (torch.distributed, NCCL over NVLink on the GPU box, gloo in CPU tests).

Semantics are the reference's multi-GPU scheme (src/cuda/cuda_sparse_grid.rs:256-322,
src/cuda/cuda_mpm_pipeline.rs:393-533): every rank owns a fixed subset of the particles; after the local sort the ranks
exchange their active-block keys, blocks active on more than one rank are tagged as halo, after G2P2G the accumulated
node sums of halo blocks are exchanged and added on every rank that holds the block, the grid update is then redundant
on shared blocks, and dt is the minimum over ranks.  Because a shared block ends with the full sum on every rank that
activated it, the result does not depend on how particles are partitioned — the bench partitions the domain into
x-slabs so that only slab faces are shared.

Differences with the reference: the payload is the 6 summed floats per node (1536 B per block instead of the 4616-byte
`HaloBlockData`), the dt minimum is an all-reduce on the u64 accumulator instead of a host loop, and ranks are
processes, so the exchange goes through `torch.distributed` collectives on device buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import abi

BLOCK_PAYLOAD_FLOATS = 6 * 64


class HaloExchange:
    """Backend-agnostic exchange of (keys, payload) between ranks.  Works on torch tensors; with NCCL the tensors are
    CUDA tensors handed to the library as raw device pointers, with gloo (CPU tests) the same protocol runs on host
    tensors against a `HaloPeer` implementation."""

    def __init__(self, dist, device):
        import torch
        self.torch, self.dist, self.device = torch, dist, device
        self.world, self.rank = dist.get_world_size(), dist.get_rank()

    def all_gather_varlen(self, t, count: int):
        """all-gather the first `count` rows of `t` from every rank -> (concatenated rows of the OTHER ranks, counts)"""
        torch, dist = self.torch, self.dist
        counts = torch.zeros(self.world, dtype=torch.int64, device=self.device)
        mine = torch.tensor([count], dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(counts, mine) if hasattr(dist, "all_gather_into_tensor") and self.device.type == "cuda" else self._gather_counts(counts, mine)
        counts_h = counts.cpu().tolist()
        mx = max(max(counts_h), 1)
        row = t.shape[1:] if t.dim() > 1 else ()
        send = torch.zeros((mx,) + tuple(row), dtype=t.dtype, device=self.device)
        if count:
            send[:count] = t[:count]
        recv = [torch.empty_like(send) for _ in range(self.world)]
        dist.all_gather(recv, send)
        others = [recv[r][:counts_h[r]] for r in range(self.world) if r != self.rank and counts_h[r] > 0]
        if others:
            return torch.cat(others, dim=0), counts_h
        return send[:0], counts_h

    def _gather_counts(self, counts, mine):
        lst = [self.torch.zeros_like(mine) for _ in range(self.world)]
        self.dist.all_gather(lst, mine)
        counts.copy_(self.torch.cat(lst))

    def all_reduce_min_u64(self, value: int) -> int:
        torch, dist = self.torch, self.dist
        # u64 min through two i64-safe halves (NCCL has no u64 min in torch): dt_acc < 2^63 always (dt * 1e12)
        t = torch.tensor([min(value, (1 << 63) - 1)], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return int(t.item())


class DistributedMpmPipeline:
    """`CudaMpmPipeline::step` over several single-GPU contexts living in different processes."""

    def __init__(self, ctx, dist=None, device=None):
        import torch
        import torch.distributed as tdist
        self.ctx = ctx
        self.dist = dist or tdist
        self.device = device if device is not None else torch.device("cuda", ctx.device)
        self.x = HaloExchange(self.dist, self.device)
        self.torch = torch
        self._keys = torch.zeros(1024, dtype=torch.int64, device=self.device)
        self._hkeys = torch.zeros(256, dtype=torch.int64, device=self.device)
        self._hpay = torch.zeros((256, BLOCK_PAYLOAD_FLOATS), dtype=torch.float32, device=self.device)
        self.halo_blocks_last = 0
        self.halo_bytes_last = 0

    def _ensure(self, name, n, cols=None):
        t = getattr(self, name)
        if t.shape[0] < n:
            shape = (int(n * 1.5) + 64,) + ((cols,) if cols else ())
            setattr(self, name, self.torch.zeros(shape, dtype=t.dtype, device=self.device))
        return getattr(self, name)

    def substep(self, params: abi.SolverParameters, gravity, remaining: float) -> float:
        ctx, l, h = self.ctx, self.ctx._lib, self.ctx._h
        vp = C.c_void_p
        info = ctx.stage_sort()
        # ---- halo tagging: all-gather of active block keys (cuda_sparse_grid.rs:256-322)
        nab = info.num_active_blocks
        keys = self._ensure("_keys", nab)
        cnt = C.c_uint32()
        ctx._call("halo_active_keys", vp(keys.data_ptr()), keys.shape[0], C.byref(cnt))
        remote, _ = self.x.all_gather_varlen(keys, nab)
        nh = C.c_uint32()
        remote = remote.contiguous()
        ctx._call("halo_tag", vp(remote.data_ptr()) if remote.numel() else vp(0), int(remote.shape[0]), C.byref(nh))
        ctx.stage_reset_grid()
        # ---- dt: local candidate, global min (cuda_mpm_pipeline.rs:393-398)
        min_dt = float(np.float32(params.dt) / np.float32(params.max_num_substeps))
        max_dt = float(min(np.float32(remaining), np.float32(params.max_substep_dt)))
        ctx.stage_estimate_dt(min_dt, max_dt)
        acc = self.x.all_reduce_min_u64(ctx.dt_acc())
        ctx.set_dt_acc(acc)
        dt = float(np.float32(acc) / np.float32(1.0e12))
        # ---- transfers + halo accumulate (cuda_mpm_pipeline.rs:407-533)
        if info.num_dispatch_blocks > 0:
            ctx.stage_g2p2g(dt, params.damage_model)
        nhalo = nh.value
        hk = self._ensure("_hkeys", nhalo)
        hp = self._ensure("_hpay", nhalo, BLOCK_PAYLOAD_FLOATS)
        got = C.c_uint32()
        ctx._call("halo_pack", vp(hk.data_ptr()), vp(hp.data_ptr()), hk.shape[0], C.byref(got))
        rk, _ = self.x.all_gather_varlen(hk, got.value)
        rp, _ = self.x.all_gather_varlen(hp, got.value)
        rk, rp = rk.contiguous(), rp.contiguous()
        if rk.shape[0]:
            ctx._call("halo_merge", vp(rk.data_ptr()), vp(rp.data_ptr()), int(rk.shape[0]))
        self.halo_blocks_last = int(got.value)
        self.halo_bytes_last = int(got.value) * (8 + 4 * BLOCK_PAYLOAD_FLOATS)
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
