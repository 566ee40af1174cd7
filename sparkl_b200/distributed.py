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
        # accumulated floats per block on the wire: 6 channels x 4^DIM cells (HaloBlockData minus the unsummed fields)
        self.payload_floats = 6 * (16 if getattr(ctx, "dim", 3) == 2 else 64)
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
        self._hp = torch.zeros((cap, self.payload_floats), dtype=torch.float32, device=self.device)
        self._allhk = torch.zeros((self.world, cap), dtype=torch.int64, device=self.device)
        self._allhp = torch.zeros((self.world, cap, self.payload_floats), dtype=torch.float32, device=self.device)

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
            self.halo_bytes_last = int(nh.value) * (8 + 4 * self.payload_floats)
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


# ---------------------------------------------------------------------------------------------------- particle migration
# The reference never moves particles between devices (static ownership); results do not depend on ownership, but
# the halo does: a slab decomposition only keeps its thin halo while every rank's particles stay inside its slab.  `migrate`
# re-homes the particles that left: 128-byte records (the six reference structs + a global id) are exchanged point to point
# (`torch.distributed` send / recv: NCCL over NVLink on the GPU box, gloo in the CPU tests) and the local set is re-uploaded.
# It is a collective, host-orchestrated slow path meant to run every few frames (a particle moves at most one cell per substep,
# SURVEY quirk 3, so ownership may lag by a block without harm); a device-side pack + NVLink push is the planned fast path.
PARTICLE_RECORD = np.dtype([("status", abi.STATUS), ("position", abi.POSITION), ("velocity", abi.VELOCITY), ("volume", abi.VOLUME),
                            ("phase", abi.PHASE), ("data", abi.DATA), ("gid", "<u4")])
assert PARTICLE_RECORD.itemsize == 128


class SlabPartition:
    """ownership by block-x coordinate: rank r owns the particles whose block (gpu_grid.rs:271-279) has x in [bounds[r], bounds[r+1])"""

    def __init__(self, bounds):
        self.bounds = np.asarray(bounds, np.int64)  # world + 1 increasing block-x coordinates; the outer two are open ended
        assert len(self.bounds) >= 2 and (np.diff(self.bounds) > 0).all()

    @property
    def world(self):
        return len(self.bounds) - 1

    @classmethod
    def uniform(cls, x_min: float, x_max: float, world: int, cell_width: float):
        """`world` slabs of equal width between the world coordinates x_min and x_max, snapped to block boundaries"""
        lo, hi = cls.block_x(np.array([x_min, x_max], np.float32), cell_width)
        edges = np.round(np.linspace(lo, hi + 1, world + 1)).astype(np.int64)
        for r in range(1, world + 1):
            edges[r] = max(edges[r], edges[r - 1] + 1)
        return cls(edges)

    @classmethod
    def balanced(cls, x: np.ndarray, world: int, cell_width: float):
        """`world` slabs holding (nearly) equal particle counts: the edges are the quantiles of the particles' block-x coordinates"""
        bx = np.sort(cls.block_x(x, cell_width))
        edges = np.empty(world + 1, np.int64)
        edges[0], edges[world] = bx[0], bx[-1] + 1
        for r in range(1, world):
            edges[r] = bx[min(len(bx) - 1, (len(bx) * r) // world)]
        for r in range(1, world + 1):
            edges[r] = max(edges[r], edges[r - 1] + 1)
        return cls(edges)

    @staticmethod
    def block_x(x: np.ndarray, cell_width: float) -> np.ndarray:
        q = np.asarray(x, np.float32) / np.float32(cell_width)
        cell = np.trunc(q + np.copysign(np.float32(0.5), q)).astype(np.int64)  # Rust f32::round
        return (cell - 2) // 4

    def owner(self, x: np.ndarray, cell_width: float) -> np.ndarray:
        bx = self.block_x(x, cell_width)
        return np.clip(np.searchsorted(self.bounds[1:-1], bx, side="right"), 0, self.world - 1)


def pack_records(pc: abi.ParticleComponents, gids: np.ndarray) -> np.ndarray:
    rec = np.zeros(len(pc), PARTICLE_RECORD)
    for name in ("status", "position", "velocity", "volume", "phase", "data"):
        rec[name] = getattr(pc, name)
    rec["gid"] = gids
    return rec


def unpack_records(rec: np.ndarray):
    pc = abi.ParticleComponents(len(rec))
    for name in ("status", "position", "velocity", "volume", "phase", "data"):
        setattr(pc, name, np.ascontiguousarray(rec[name]))
    return pc, np.ascontiguousarray(rec["gid"])


def migrate(engine, gids: np.ndarray, partition: SlabPartition, dist=None, device=None):
    """Collective.  Moves every particle of `engine` (a SingleGpuMpmContext, or the oracle in CPU tests) whose block left this
    rank's slab to the rank that owns it, and re-uploads the local set.  `gids` are the global ids of the local particles (in
    local particle-id order).  Returns (new_gids, number sent, number received).  (A rank may end up without particles:
    the NVLink path keeps it in its device-side barriers.)"""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    device = device if device is not None else torch.device("cpu")
    world, rank = dist.get_world_size(), dist.get_rank()
    assert partition.world == world
    pc = engine.read_all()
    rec = pack_records(pc, np.asarray(gids, np.uint32))
    owner = partition.owner(pc.position["point"][:, 0], engine.cell_width)
    counts = torch.tensor([int((owner == r).sum()) if r != rank else 0 for r in range(world)], dtype=torch.int64, device=device)
    table = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(table, counts)
    table = torch.stack(table).cpu().numpy()  # table[src, dst]
    ops, recv_bufs, send_bufs = [], {}, {}
    for peer in range(world):
        if peer == rank:
            continue
        if table[rank, peer]:
            buf = torch.from_numpy(np.ascontiguousarray(rec[owner == peer]).view(np.uint8).copy()).to(device)
            send_bufs[peer] = buf
            ops.append(dist.P2POp(dist.isend, buf, peer))
        if table[peer, rank]:
            recv_bufs[peer] = torch.empty(int(table[peer, rank]) * PARTICLE_RECORD.itemsize, dtype=torch.uint8, device=device)
            ops.append(dist.P2POp(dist.irecv, recv_bufs[peer], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    stay = rec[owner == rank]
    incoming = [recv_bufs[p].cpu().numpy().view(PARTICLE_RECORD) for p in sorted(recv_bufs)]
    merged = np.concatenate([stay] + incoming) if incoming else stay
    n_sent, n_recv = int(table[rank].sum()), int(table[:, rank].sum())
    if table.sum() > 0:
        # EVERY rank re-uploads as soon as any particle moved anywhere: the peer-memory path requires the same call sequence on
        # all ranks (a re-upload invalidates the fused dt candidate, and the stand-alone dt kernel that follows ends in a
        # device-side barrier every rank must reach)
        new_pc, new_gids = unpack_records(merged)
        engine.upload(new_pc)
    else:
        new_gids = np.asarray(gids, np.uint32)
    return new_gids, n_sent, n_recv


def migrate_device(ctx, partition: SlabPartition, dist=None, device=None, wire=None, hysteresis: int = 0):
    """Collective.  The device-side counterpart of `migrate` for a SingleGpuMpmContext (dim3): the leaving particles are packed
    into 128-byte rows by the GPU (`spk_particles_migrate_out`: stable partition by destination slab), the rows travel between the
    ranks' device buffers (`torch.distributed` send / recv: NCCL over NVLink), and the arrivals are appended on the device
    (`spk_particles_migrate_in`).  Only the per-destination counts cross PCIe; global ids live in the context (`ctx.set_gids`,
    `ctx.read_gids`).  When nothing moves anywhere no particle state is touched and the sort state stays valid.  `wire` (default:
    `device`) is where the rows sit during the exchange: the CPU when the process group is gloo (tests on one GPU), which cannot
    send CUDA tensors.  `hysteresis` (blocks): a rank keeps the particles that are less than that many block columns outside its slab —
    ownership may lag behind the position without any effect on the result (only the halo grows by as many columns), and particles
    that oscillate across a slab boundary (a lattice on cell-boundary ties does, every substep) are not sent back and forth.
    Returns (number sent, number received)."""
    import torch
    import torch.distributed as tdist
    dist = dist or tdist
    device = device if device is not None else torch.device("cuda", ctx.device)
    world, rank = dist.get_world_size(), dist.get_rank()
    assert partition.world == world
    row = PARTICLE_RECORD.itemsize
    bounds = np.array(partition.bounds, np.int64)
    if hysteresis:  # this rank's view: its own slab widened; the bounds between the OTHER ranks stay where they are
        if rank > 0:
            bounds[rank] = max(bounds[rank] - hysteresis, bounds[rank - 1] + 1)
        if rank < world - 1:
            bounds[rank + 1] = min(bounds[rank + 1] + hysteresis, bounds[rank + 2] - 1)
    import os as _os, time as _time
    _dbg = _os.environ.get("SPK_MIGRATE_DEBUG")
    _t = [_time.perf_counter()]
    def _mark():
        if _dbg:
            torch.cuda.synchronize()
            _t.append(_time.perf_counter())
    counts = ctx.migrate_out(bounds, rank)  # count only
    _mark()
    mine = torch.tensor(counts.astype(np.int64), dtype=torch.int64, device=device)
    table = [torch.zeros(world, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(table, mine)
    table = torch.stack(table).cpu().numpy()  # table[src, dst]
    n_sent, n_recv = int(table[rank].sum()), int(table[:, rank].sum())
    _mark()
    if table.sum() == 0:
        return 0, 0
    send = torch.empty(max(n_sent, 1) * row, dtype=torch.uint8, device=device)
    if n_sent:
        again = ctx.migrate_out(bounds, rank, send.data_ptr(), n_sent)
        assert np.array_equal(again, counts)
    _mark()
    offs = np.concatenate([[0], np.cumsum(table[rank])]) * row  # rows are grouped by destination rank
    recv = torch.empty(max(n_recv, 1) * row, dtype=torch.uint8, device=device)
    roffs = np.concatenate([[0], np.cumsum(table[:, rank])]) * row
    wire = wire if wire is not None else device
    staged = wire != device
    wsend = send.to(wire) if staged else send
    wrecv = torch.empty_like(recv, device=wire) if staged else recv
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if table[rank, peer]:
            ops.append(dist.P2POp(dist.isend, wsend[int(offs[peer]):int(offs[peer + 1])], peer))
        if table[peer, rank]:
            ops.append(dist.P2POp(dist.irecv, wrecv[int(roffs[peer]):int(roffs[peer + 1])], peer))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    if staged:
        recv.copy_(wrecv)
    if device.type == "cuda":
        torch.cuda.current_stream(device).synchronize()  # the rows have landed before the context's stream reads them
    _mark()
    ctx.migrate_in(recv.data_ptr() if n_recv else 0, n_recv)  # n_recv == 0: marks the set as edited (same path on every rank)
    _mark()
    if _dbg and rank == 0:
        print("migrate_device ms: count %.2f gather %.2f out %.2f exchange %.2f in %.2f (sent %d recv %d)" % tuple(
            [1e3 * (_t[i + 1] - _t[i]) for i in range(5)] + [n_sent, n_recv]), flush=True)
    return n_sent, n_recv
