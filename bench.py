#!/usr/bin/env python
"""bench.py — particle-substeps/s of the 3D sand workload (BASELINE.json metric) + G2P2G HBM roofline.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

ours:      one "step" = one MPM substep (sort -> reset -> dt -> fused G2P2G -> grid update) of the sand3 scene
           (examples3d/sand3.rs scaled to 4M particles per GPU: BASELINE.json configs[2]) through the C ABI
           (spk_step with stop_after_one_substep).  N > 1: x-slabs of 4M particles per rank with halo exchange (weak).
reference: the CPU arm — the oracle port of the reference algorithm (oracle/mpm_oracle.cpp, OpenMP, all host cores) on
           a bounded sample of the same scene; the reference's Rust CPU path cannot be built here (no cargo/rustc).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "particle_substeps_per_sec"
UNIT = "particle-substeps/s"
# algorithmic bytes per particle-substep of the fused G2P2G kernel, sand (SURVEY.md §8d / DESIGN.md §5):
# particle records R112 + W108 + ParticleData R16+W16 = 252, plus grid (72 R + 24 W) * 64 / 512 = 12
A_G2P2G_SAND = 264.0
FULL = (200, 100, 100)       # per-GPU lattice (x2 stacked) = 4,000,000 particles
CPU_SAMPLE = (100, 50, 50)   # bounded CPU sample = the reference's own sand3 size (examples3d/sand3.rs:51-71): 500,000 particles
# BASELINE.json configs[1..4].  weak: every rank owns an x-slab of `dims`; strong: ONE scene of `dims`, cut into x-slabs of blocks
# (SlabPartition) with device-side particle migration between the ranks every MIGRATE_EVERY substeps
CONFIGS = {
    "S4M": {"scene": "sand3", "dims": (200, 100, 100), "scaling": "weak", "what": "examples3d/sand3.rs scaled: 2 stacked lattices, Drucker-Prager sand over corotated block"},
    "F1M": {"scene": "fluids3", "dims": (160, 80, 80), "scaling": "weak", "what": "examples3d/fluids3.rs scaled: Monaghan weakly-compressible dam break (1 GPU)"},
    "C16M": {"scene": "cube_through_sand3", "dims": (400, 200, 200, 100), "scaling": "strong", "what": "examples3d/cube_through_sand3.rs scaled: kinematic block through Drucker-Prager sand"},
    "N32M": {"scene": "neo_hookean_block", "dims": (400, 400, 200), "scaling": "weak", "what": "synthetic uniform-density Neo-Hookean block, 32 M particles per GPU (256 M on 8)"},
}
MIGRATE_EVERY = int(os.environ.get("SPK_BENCH_MIGRATE_EVERY", "20"))
CPU_SAMPLES = {"S4M": (100, 50, 50), "F1M": (80, 40, 40), "C16M": (100, 50, 50, 25), "N32M": (100, 100, 50)}  # bounded CPU samples of the same scenes


def _peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


_SAMPLER_SRC = r"""
import sys, time
idx = int(sys.argv[1])
try:
    import pynvml as N
    N.nvmlInit(); h = N.nvmlDeviceGetHandleByIndex(idx)
    print("max", N.nvmlDeviceGetMaxClockInfo(h, N.NVML_CLOCK_SM), flush=True)
    while True:
        try: m = N.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception: m = N.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        print(time.time(), N.nvmlDeviceGetClockInfo(h, N.NVML_CLOCK_SM), int(m), flush=True)
        time.sleep(0.005)
except Exception as e:
    print("err", e, flush=True)
"""


def summarize_clock_samples(lines, t0, t1, t_load0, t_load1, reasons_bits):
    """(pure: unit-tested on the CPU) `lines` = the sampler's output.  Samples taken inside the timed region [t0, t1] are used when there
    are any; a timed region shorter than one NVML poll (a poll of a busy GPU can take tens of ms) falls back to the samples of the
    loaded phases around it [t_load0, t_load1] — settle, warm-up, the timed steps and the repeat of the same K substeps that collects
    the per-kernel timings, all the same kernels at the same load — and says so in `window`."""
    mx, timed, loaded = None, [], []
    for line in lines:
        f = line.split()
        if len(f) == 2 and f[0] == "max":
            mx = float(f[1])
        elif len(f) == 3:
            try:
                t, mhz, mask = float(f[0]), float(f[1]), int(f[2])
            except ValueError:
                continue
            if t0 <= t <= t1:
                timed.append((mhz, mask))
            if t_load0 <= t <= t_load1:
                loaded.append((mhz, mask))
    use, window = (timed, "timed region") if timed else (loaded, "loaded phases around the timed region (it was shorter than one NVML poll)")
    mask = 0
    for _, m in use:
        mask |= m
    reasons = sorted(n for n, bit in reasons_bits.items() if mask & bit)
    sm = [mhz for mhz, _ in use]
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(sm), "window": window if sm else None,
            "source": "nvml (separate process)"}


class ClockSampler:
    """samples SM clocks / throttle reasons during the timed region from a SEPARATE process (NVML, 200 Hz): an in-process
    poller competes with the launching thread for the GIL and the driver lock and was measured to slow the multi-GPU step"""

    REASONS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.proc = subprocess.Popen([sys.executable, "-c", _SAMPLER_SRC, str(index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t0 = self.t1 = self.t_load0 = None
        self.head = []

    def wait_ready(self, timeout=20.0):
        """blocks until the sampler has delivered its first sample (interpreter start + NVML initialisation take a second or two on a fresh
        box; the loaded phases of the default run are shorter than that)"""
        import select
        t_end = time.time() + timeout
        while time.time() < t_end and self.proc.poll() is None:
            r, _, _ = select.select([self.proc.stdout], [], [], 0.25)
            if r:
                line = self.proc.stdout.readline()
                if not line:
                    break
                self.head.append(line)
                if len(line.split()) == 3 or line.startswith("err"):
                    break

    def load_begins(self):
        self.t_load0 = time.time()

    def start(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()

    def finish(self):
        t_load1 = time.time()
        if self.t1 is None:
            self.t1 = t_load1
        time.sleep(0.02)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        return summarize_clock_samples([l.rstrip("\n") for l in self.head] + out.splitlines(), self.t0, self.t1, self.t_load0 if self.t_load0 is not None else self.t0, t_load1, self.REASONS)


def build_scene(rank, world, dims, config="S4M"):
    """this rank's particles of the configuration, and (strong scaling) the slab partition they were cut with"""
    from sparkl_b200 import scenes
    cfg = CONFIGS[config]
    if cfg["scene"] == "sand3":
        ni, nj, nk = dims
        ext = float(np.float32(0.2) / np.float32(4.0)) * 2.0
        # slabs side by side along x; one floor + outer walls for the whole domain, identical on every rank
        sc = scenes.sand3(ni, nj, nk, x_offset=rank * ni * ext, with_colliders=False)
        sc.colliders = scenes.box_colliders((-4.0, 2.0, -4.0), (world * ni * ext + 4.0, 2.0 + 2 * nj * ext + 8.0, nk * ext + 4.0))
        return sc, None
    if cfg["scene"] == "neo_hookean_block":
        ni, nj, nk = dims
        sc = scenes.neo_hookean_block(ni, nj, nk, x_offset=rank * ni * 0.1)
        return sc, None
    if cfg["scene"] == "fluids3":
        if world != 1:
            raise SystemExit("config F1M is a single-GPU configuration (BASELINE.json configs[1])")
        return scenes.fluids3(*dims), None
    # strong scaling: the same scene on every rank count, particles dealt to the ranks by the block column that holds them
    from sparkl_b200.distributed import SlabPartition
    sc = scenes.cube_through_sand3(*dims)
    if world == 1:
        return sc, None
    x = sc.particles.position["point"][:, 0]
    part = SlabPartition.balanced(x, world, sc.cell_width)  # equal particle counts per slab (the kinematic block starts beside the sand: equal widths would not be)
    mine = np.nonzero(part.owner(x, sc.cell_width) == rank)[0]
    sub = scenes.Scene(sc.name, sc.cell_width, sc.particles.take(mine), sc.models, sc.colliders, sc.gravity, sc.params)
    sub.global_ids = mine.astype(np.uint32)
    sub.total_particles = sc.num_particles
    return sub, part


def host_threads():
    """the host cores this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm must not inherit that)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_baseline(steps, warmup, config="S4M"):
    """the oracle port timed on the host cores (all threads) on a bounded sample of the workload"""
    from oracle import binding as ob
    from sparkl_b200 import abi
    threads = ob.set_threads(host_threads(), fast=True)
    dims = CPU_SAMPLES[config]
    sc, _ = build_scene(0, 1, dims, config)
    e = ob.OracleEngine(sc.cell_width, fast=True)  # liboracle_fast.so: the oracle's source built for speed (f32 SVD and node sums, -O3)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    p = abi.SolverParameters(stop_after_one_substep=1)
    for _ in range(warmup):
        e.step(p, sc.gravity)
    t0 = time.perf_counter()
    n = 0
    for _ in range(steps):
        _, t = e.step(p, sc.gravity)
        n += t.num_substeps
    dt = time.perf_counter() - t0
    return {"value": sc.num_particles * n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{config} scene at lattice {'x'.join(str(d) for d in dims)} = {sc.num_particles} particles, {n} substeps, oracle/mpm_oracle.cpp built for speed (ORC_FAST: f32 Jacobi SVD and node sums, 8-colour block-parallel fused G2P2G, -O3, OpenMP on {threads} threads)"}, dt / max(n, 1)


def reference_gpu(sc, settle, steps):
    """the reference's OWN GPU kernels (its shipped PTX assembled for sm_100a, oracle/_ref) on the same scene and the
    same B200, sequenced as src/cuda/cuda_mpm_pipeline.rs does — "the kernel to beat" """
    try:
        from oracle import binding as ob
        from sparkl_b200 import abi
        if not ob.ref_available():
            return None
        r = ob.RefEngine(sc.cell_width)
        r.models_upload(sc.models); r.colliders_upload(sc.colliders); r.upload(sc.particles)
        p = abi.SolverParameters(stop_after_one_substep=1)
        for _ in range(settle + 3):
            r.step(p, sc.gravity)
        tot = g = 0.0
        for _ in range(steps):
            _, t = r.step(p, sc.gravity)
            tot += t.total_ms
            g += t.g2p2g_ms
        return {"value": sc.num_particles * steps / (tot * 1e-3), "unit": UNIT, "ms_per_step": tot / steps, "g2p2g_ms_per_launch": g / steps,
                "g2p2g_GBps": A_G2P2G_SAND * sc.num_particles * steps / (g * 1e-3) / 1e9,
                "what": "reference sparkl3d-kernels.ptx assembled for sm_100a (-maxrregcount 130), launch sequence of cuda_mpm_pipeline.rs, same scene, same GPU"}
    except Exception as e:  # the reference arm is optional evidence, never a reason to lose the bench line
        return {"error": str(e)[:200]}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    base, ms = cpu_baseline(args.steps, args.warmup, args.config)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {CONFIGS[args.config]['what']}, CPU arm on a bounded sample", "name": args.config, "sample": base["sample"]},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def pinned_like(arr):
    import torch
    t = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
    out = t.numpy().view(arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out, t


def multi_gpu_check(rank, world, local, dist, args):
    """driver-side evidence that the N-GPU path computes what one GPU computes: a small jittered sand scene cut into `world` x-slabs
    runs two frames through the same halo path as the benchmark; rank 0 runs the whole scene alone; positions are compared"""
    import torch
    from sparkl_b200 import abi, scenes
    from sparkl_b200.pipeline import context_from_scene
    from sparkl_b200.distributed import DistributedMpmPipeline, NvlinkMpmPipeline
    sc = scenes.jitter(scenes.sand3(8 * world, 8, 8), 0.02, seed=9)
    x = sc.particles.position["point"][:, 0]
    edges = np.linspace(x.min(), x.max() + 1e-3, world + 1)
    mine = np.nonzero((x >= edges[rank]) & (x < edges[rank + 1]))[0]
    part = scenes.Scene(sc.name, sc.cell_width, sc.particles.take(mine), sc.models, sc.colliders, sc.gravity, sc.params)
    ctx = context_from_scene(part, device=local)
    nsub = 0
    if args.halo == "nvlink":
        pipe = NvlinkMpmPipeline(ctx, dist)
        for _ in range(2):
            nsub += pipe.step(sc.params, sc.gravity)[1].num_substeps
        halo = pipe.halo_blocks()
        pipe.close()
    else:
        pipe = DistributedMpmPipeline(ctx, dist, torch.device("cuda", local))
        for _ in range(2):
            nsub += pipe.step(sc.params, sc.gravity)[1]
        halo = None
    pos = torch.zeros((len(sc.particles), 3), dtype=torch.float32, device="cuda")
    pos[torch.from_numpy(mine).cuda()] = torch.from_numpy(ctx.read(abi.ATTR_POSITION)["point"].copy()).cuda()
    dist.all_reduce(pos)
    del ctx
    if rank != 0:
        return None
    ref = context_from_scene(sc, device=local)
    nref = sum(ref.step(sc.params, sc.gravity)[1].num_substeps for _ in range(2))
    err = float(np.abs(pos.cpu().numpy() - ref.read(abi.ATTR_POSITION)["point"]).max() / sc.cell_width)
    return {"scene": f"sand3 {8 * world}x8x8 x2 jittered, {len(sc.particles)} particles in {world} x-slabs, 2 frames", "substeps": nsub, "substeps_single_gpu": nref,
            "max_dx_over_h_vs_single_gpu": err, "halo_blocks_rank0": halo, "ok": bool(err < 1e-4 and nsub == nref)}


def h2d_ceiling(pins, dist, world, local):
    """what the box gives when every rank copies its pinned frame input host -> device at the same time (the e2e frame cannot be faster)"""
    import torch
    dev = [torch.empty(p[1].numel(), dtype=torch.uint8, device="cuda") for p in pins]
    nbytes = sum(p[1].numel() for p in pins)
    best = 0.0
    for _ in range(3):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for d, p in zip(dev, pins):
            d.copy_(p[1], non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        best = max(best, nbytes / dt / 1e9)
    return {"per_rank_GBps": best, "aggregate_GBps": best * world, "what": "concurrent cudaMemcpyAsync of every rank's pinned frame input, best of 3 (slowest rank)"}


def run_ours(args):
    import torch
    from sparkl_b200 import abi
    from sparkl_b200.pipeline import context_from_scene
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # each rank keeps to its own share of the host cores (and, with them, of the memory controllers next to its GPU's PCIe root):
    # eight ranks staging 0.5 GB per frame from pinned memory otherwise contend for the same cores and NUMA node
    nloc = int(os.environ.get("LOCAL_WORLD_SIZE", world))
    try:
        cores = sorted(os.sched_getaffinity(0))
        if nloc > 1 and len(cores) >= nloc:
            per = len(cores) // nloc
            os.sched_setaffinity(0, cores[local * per:(local + 1) * per])
    except (AttributeError, OSError):
        pass
    cfg = CONFIGS[args.config]
    dims = tuple(args.dims) if args.dims else cfg["dims"]
    check = multi_gpu_check(rank, world, local, dist, args) if world > 1 else None
    sc, partition = build_scene(rank, world, dims, args.config)
    npart = sc.num_particles                                           # this rank's particles
    total = getattr(sc, "total_particles", None) or npart * world       # the job's particles
    ctx = context_from_scene(sc, device=local)
    if partition is not None:
        ctx.set_gids(sc.global_ids)
        # NCCL opens its point-to-point connections lazily, ~100 ms the first time a PAIR of ranks talks; the migration of a block that
        # crosses the slabs reaches new pairs one after the other: open all of them once, outside the timed region
        warm = torch.zeros(world, dtype=torch.uint8, device=torch.device("cuda", local))
        dist.all_to_all_single(torch.empty_like(warm), warm)
        torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local))
    params = abi.SolverParameters(stop_after_one_substep=1)
    pipe = None
    if world > 1:
        if args.halo == "nvlink":  # device-initiated NVLink peer-memory exchange inside spk_step (spk_dist.cu)
            from sparkl_b200.distributed import NvlinkMpmPipeline
            NvlinkMpmPipeline(ctx, dist)
        else:                      # the portable torch.distributed (NCCL) protocol, kept for A/B
            from sparkl_b200.distributed import DistributedMpmPipeline
            pipe = DistributedMpmPipeline(ctx, dist, torch.device("cuda", local))
    migrated = [0, 0, 0.0]  # particles sent by this rank, migration calls, host seconds spent in them (they synchronise)
    substeps_done = [0]

    def one_substep(timings=False):
        if partition is not None and substeps_done[0] and substeps_done[0] % MIGRATE_EVERY == 0:
            # strong-scaled scene with a moving body: particles whose block column left this rank's slab move to their new owner
            # (spk_particles_migrate_out / _in + NCCL point-to-point over NVLink); a collective, every rank calls it at the same substep
            from sparkl_b200.distributed import migrate_device
            t_m = time.perf_counter()
            sent, _ = migrate_device(ctx, partition, dist, hysteresis=1)
            migrated[0] += sent
            migrated[1] += 1
            migrated[2] += time.perf_counter() - t_m
        substeps_done[0] += 1
        if pipe is None:  # without timings spk_step returns once the substep is enqueued: no host-side gap between steps
            return ctx.step(params, sc.gravity, timings=timings)[1]
        pipe.step(params, sc.gravity)
        return None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None  # started early: its process is up by the time the timed region begins
    # bring the scene into a loaded state first (material compressing under gravity, F != I): the rest lattice makes the
    # per-particle SVD converge in one sweep and would flatter the kernel.  Untimed, not part of the warm-up contract.
    if sampler:
        sampler.wait_ready()
        sampler.load_begins()
    for _ in range(args.settle):
        one_substep()
    for _ in range(max(args.warmup, 3)):
        one_substep()
    barrier()
    if sampler:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g2p2g_ms = launches = 0.0
    stage_ms = np.zeros(5)
    ctx.g2p2g_time(enable=True)  # bracket every g2p2g launch of the timed region with CUDA events (no synchronisation)
    ev0.record(stream)
    for _ in range(args.steps):
        one_substep()
    ev1.record(stream)
    barrier()
    ms = ev0.elapsed_time(ev1)
    g2p2g_live_ms, g2p2g_live_n = ctx.g2p2g_time(enable=False)
    if sampler:
        sampler.stop()
    # per-kernel CUDA-event timings (inside spk_step, on its stream) over the SAME K substeps repeated: they need one event
    # synchronisation per substep, so they are taken in a second pass and do not perturb the throughput figure above
    if world == 1:
        for _ in range(args.steps):
            t = one_substep(timings=True)
            g2p2g_ms += t.g2p2g_ms
            launches += t.kernel_launches
            stage_ms += [t.grid_resize_and_sort_ms, t.reset_grid_ms, t.estimate_timestep_ms, t.g2p2g_ms, t.grid_update_ms]
    if world > 1:
        torch.cuda.synchronize()
    clocks = sampler.finish() if sampler else None
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = total * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: upload (pinned H2D) -> one frame -> read back pos/vel/phase
    # into pinned host buffers; every rank does it for its own slab, the frame time is the max over ranks
    e2e = None
    if args.e2e_frames > 0:
        host = ctx.read_all()  # the LOADED state reached by the timed substeps (not the rest lattice): every frame re-uploads it
        pins = [pinned_like(a) for a in (host.status, host.position, host.velocity, host.volume, host.phase, host.data)]
        pc = abi.ParticleComponents(0)
        pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data = [p[0] for p in pins]
        npart = len(host)  # (strong scaling with migration: this rank's count may have changed since the upload)
        outs = {w: pinned_like(np.empty(npart, abi.ATTR_DTYPES[w])) for w in (abi.ATTR_POSITION, abi.ATTR_VELOCITY, abi.ATTR_PHASE)}
        frame = abi.SolverParameters()
        h2d = sum(p[0].nbytes for p in pins)
        d2h = sum(o[0].nbytes for o in outs.values())
        # (a) blocking calls: upload -> frame -> three reads, each waiting for its transfer
        subs = 0
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_frames):
            ctx.upload(pc)
            _, t = ctx.step(frame, sc.gravity)
            subs += t.num_substeps
            for w, o in outs.items():
                ctx.read(w, out=o[0])
        barrier()
        wall_blocking = time.perf_counter() - t0
        per_frame = subs // args.e2e_frames
        # (b) the same work through the asynchronous entry points: the upload of frame k+1 (copy stream) and the read-back of
        # frame k (second copy stream) overlap the substeps of the neighbouring frames; every frame still moves all its
        # input host->device and its result device->host inside the timed region
        outs2 = {w: pinned_like(np.empty(npart, abi.ATTR_DTYPES[w])) for w in outs}
        nfr = max(6, 2 * args.e2e_frames)
        ctx.upload_async(pc)                                       # untimed warm-up frame: the staging buffers of the asynchronous
        ctx.step(frame, sc.gravity, timings=False)                 # paths are allocated on first use
        for w in outs:
            ctx.read_async(w, outs2[w][0])
        ctx.read_wait()
        ctx.synchronize()
        barrier()
        t0 = time.perf_counter()
        ctx.upload_async(pc, stage_only=True)                      # prefetch of frame 0's input
        for f in range(nfr):
            ctx.commit_staged()                                    # frame f's input becomes the particle state (repack kernel)
            if f + 1 < nfr:
                ctx.upload_async(pc, stage_only=True)              # frame f+1's input travels while frame f is computed
            ctx.step(frame, sc.gravity, timings=False)
            for w in outs:
                ctx.read_async(w, (outs if f % 2 == 0 else outs2)[w][0])
        ctx.read_wait()
        ctx.synchronize()
        barrier()
        wall = time.perf_counter() - t0
        if world > 1:
            tt = torch.tensor([wall, wall_blocking], dtype=torch.float64, device="cuda")
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            wall, wall_blocking = float(tt[0].item()), float(tt[1].item())
        ceiling = h2d_ceiling(pins, dist, world, local)
        e2e = {"value": total * per_frame * nfr / wall, "h2d_ceiling": ceiling, "h2d_frac_of_ceiling": (h2d * (total / npart) / (wall / nfr) / 1e9) / ceiling["aggregate_GBps"] if ceiling else None, "unit": UNIT, "h2d_bytes_per_step": int(h2d * total / npart), "d2h_bytes_per_step": int(d2h * total / npart),
               "step": f"per rank and frame: spk_particles_stage_async + _commit_staged (pinned host, all 6 attributes) + spk_step(dt=1/60: {per_frame} substeps) + "
                       "spk_particles_read_async(pos, vel, phase) into pinned host buffers; transfers of neighbouring frames overlap the compute",
               "frames": nfr, "ms_per_frame": wall * 1e3 / nfr,
               "blocking_calls": {"value": total * subs / wall_blocking, "ms_per_frame": wall_blocking * 1e3 / args.e2e_frames,
                                  "what": "spk_particles_upload + spk_step + spk_particles_read x3, nothing overlapped"}}

    # ---- the reference testbed's frame loop (testbed_plugin.rs:490-588): step one frame, read pos / vel / phase back; no
    # re-upload.  Blocking reads vs spk_particles_read_async (the copy of frame k overlaps frame k + 1)
    frame_loop = None
    if world == 1 and args.e2e_frames > 0:
        frame_loop = {}
        for mode in ("blocking", "async"):
            subs = 0
            for it in range(1 + max(4, args.e2e_frames * 2)):
                if it == 1:  # the first frame warms up (async staging buffers are allocated on first use)
                    ctx.read_wait()
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    subs = 0
                _, t = ctx.step(frame, sc.gravity)
                subs += t.num_substeps
                for w, o in outs.items():
                    if mode == "async":
                        ctx.read_async(w, o[0])
                    else:
                        ctx.read(w, out=o[0])
            ctx.read_wait()
            torch.cuda.synchronize()
            frame_loop[mode] = npart * subs / (time.perf_counter() - t0)
        frame_loop["unit"] = UNIT
        frame_loop["what"] = "spk_step(dt=1/60) + read-back of pos, vel, phase into pinned buffers per frame, no re-upload"

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = _peaks()
    roof = None
    if world == 1 and g2p2g_live_n > 0:
        g2p2g_ms_pass2, g2p2g_ms = g2p2g_ms, g2p2g_live_ms * args.steps / g2p2g_live_n  # the roofline uses the timed region's own launches
        achieved = A_G2P2G_SAND * npart * args.steps / (g2p2g_ms * 1e-3) / 1e9
        roof = {"bound": "hbm", "kernel": "g2p2g_cell_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "peak_source": peak_src, "bytes_per_particle_substep": A_G2P2G_SAND,
                "kernel_ms_per_launch": g2p2g_ms / args.steps,
                "stage_ms_per_substep": dict(zip(["sort", "reset_grid", "estimate_dt", "g2p2g", "grid_update"], (stage_ms / args.steps).round(4).tolist()))}
    base = None
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(12, 1, args.config)  # ~10-20 s of host time
    ref_gpu = None
    if world == 1 and not args.no_ref_gpu:
        ref_gpu = reference_gpu(sc, args.settle, min(args.steps, 10))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.config}: {cfg['what']}; lattice {'x'.join(str(d) for d in dims)}" + (" per GPU" if cfg["scaling"] == "weak" else " in total") +
                                   f" = {total} particles on {world} GPU(s)", "name": args.config, "particles_total": int(total), "particles_rank0": int(sc.num_particles),
                       "parallelism": (f"x-slabs x{world}, halo accumulate over {'NVLink peer memory (device-initiated)' if args.halo == 'nvlink' else 'NCCL all-gather'}"
                                       + (f", device-side particle migration every {MIGRATE_EVERY} substeps ({migrated[1]} calls, {migrated[0]} particles sent by rank 0, {1e3 * migrated[2] / max(migrated[1], 1):.2f} ms per call, one block column of hysteresis)" if partition is not None else ""))
                       if world > 1 else "single GPU", "l2": "inputs larger than L2 (particle state >= 0.25 GB per GPU)"},
            "clocks": clocks, "gpu_launches": int(launches) if launches else None, "e2e": e2e, "roofline": roof, "cpu_baseline": base,
            "reference_gpu": ref_gpu, "frame_loop": frame_loop, "multi_gpu_check": check}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


G2P2G_SOURCES = ["spk_g2p2g.cu", "spk_g2p2g_kernel.cuh", "spk_updater.cuh", "spk_math.cuh", "spk_device.cuh", "spk_internal.cuh"]


def _lib_sha16():
    """identity of the benched G2P2G build: sha256 over the sources the kernel is compiled from (stable across rebuilds of the same
    tree, unlike the bytes of the .so) and the nvcc flags"""
    import hashlib
    from sparkl_b200 import build as b
    try:
        h = hashlib.sha256()
        for f in G2P2G_SOURCES:
            h.update(open(os.path.join(b.CSRC, f), "rb").read())
        h.update(" ".join(b.FLAGS).encode())
        return h.hexdigest()[:16]
    except Exception:
        return None


def _ncu_traffic():
    """dram bytes per g2p2g launch from the committed `ncu --set full` summary (profiles/g2p2g_ncu_summary.json) — only when that
    capture was taken with THIS build of the kernel (the summary records the sha256 of the kernel's sources and flags)"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "g2p2g_ncu_summary.json")))
        if d.get("lib_sha16") and d["lib_sha16"] != _lib_sha16():
            return None
        return float(d["dram_bytes_per_launch"])
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="S4M", choices=sorted(CONFIGS), help="BASELINE.json configuration: S4M (default, the headline), F1M, C16M (strong scaling, migration on), N32M (256 M on 8 GPUs)")
    ap.add_argument("--dims", type=int, nargs="+", default=None, help="lattice of the configuration (default: its BASELINE size)")
    ap.add_argument("--e2e-frames", type=int, default=3)
    ap.add_argument("--settle", type=int, default=120, help="untimed substeps run before the warm-up to load the material")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip timing the reference's own GPU kernels (oracle/_ref)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--halo", default="nvlink", choices=["nvlink", "nccl"], help="multi-GPU halo exchange: NVLink peer memory inside spk_step, or NCCL collectives")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
