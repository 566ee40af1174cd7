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
        return summarize_clock_samples(out.splitlines(), self.t0, self.t1, self.t_load0 if self.t_load0 is not None else self.t0, t_load1, self.REASONS)


def build_scene(rank, world, dims):
    from sparkl_b200 import scenes
    ni, nj, nk = dims
    ext = float(np.float32(0.2) / np.float32(4.0)) * 2.0
    # slabs side by side along x; one floor + outer walls for the whole domain, identical on every rank
    sc = scenes.sand3(ni, nj, nk, x_offset=rank * ni * ext, with_colliders=False)
    sc.colliders = scenes.box_colliders((-4.0, 2.0, -4.0), (world * ni * ext + 4.0, 2.0 + 2 * nj * ext + 8.0, nk * ext + 4.0))
    return sc


def cpu_baseline(steps, warmup, dims=CPU_SAMPLE):
    """the oracle port timed on the host cores (all threads) on a bounded sample of the workload"""
    from oracle import binding as ob
    from sparkl_b200 import abi
    threads = ob.set_threads(0)
    sc = build_scene(0, 1, dims)
    e = ob.OracleEngine(sc.cell_width)
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
            "sample": f"sand3 lattice {dims[0]}x{dims[1]}x{dims[2]} x2 = {sc.num_particles} particles, {n} substeps, oracle/mpm_oracle.cpp (OpenMP)"}, dt / max(n, 1)


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
    base, ms = cpu_baseline(args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "sand3 (examples3d/sand3.rs) Drucker-Prager sand on corotated block, CPU arm on a bounded sample", "sample": base["sample"]},
            "cpu_baseline": base, "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def pinned_like(arr):
    import torch
    t = torch.empty(arr.nbytes, dtype=torch.uint8).pin_memory()
    out = t.numpy().view(arr.dtype).reshape(arr.shape)
    out[...] = arr
    return out, t


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
    dims = tuple(args.dims) if args.dims else FULL
    sc = build_scene(rank, world, dims)
    npart = sc.num_particles
    ctx = context_from_scene(sc, device=local)
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

    def one_substep(timings=False):
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
    value = npart * world * args.steps / (ms * 1e-3)

    # ---- end to end through the public API with HOST buffers: upload (pinned H2D) -> one frame -> read back pos/vel/phase
    # into pinned host buffers; every rank does it for its own slab, the frame time is the max over ranks
    e2e = None
    if args.e2e_frames > 0:
        host = sc.particles
        pins = [pinned_like(a) for a in (host.status, host.position, host.velocity, host.volume, host.phase, host.data)]
        pc = abi.ParticleComponents(0)
        pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data = [p[0] for p in pins]
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
        e2e = {"value": npart * world * per_frame * nfr / wall, "unit": UNIT, "h2d_bytes_per_step": h2d * world, "d2h_bytes_per_step": d2h * world,
               "step": f"per rank and frame: spk_particles_stage_async + _commit_staged (pinned host, all 6 attributes) + spk_step(dt=1/60: {per_frame} substeps) + "
                       "spk_particles_read_async(pos, vel, phase) into pinned host buffers; transfers of neighbouring frames overlap the compute",
               "frames": nfr, "ms_per_frame": wall * 1e3 / nfr,
               "blocking_calls": {"value": npart * world * subs / wall_blocking, "ms_per_frame": wall_blocking * 1e3 / args.e2e_frames,
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
        roof = {"bound": "hbm", "kernel": "g2p2g_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": _ncu_traffic(), "peak_source": peak_src, "bytes_per_particle_substep": A_G2P2G_SAND,
                "kernel_ms_per_launch": g2p2g_ms / args.steps,
                "stage_ms_per_substep": dict(zip(["sort", "reset_grid", "estimate_dt", "g2p2g", "grid_update"], (stage_ms / args.steps).round(4).tolist()))}
    base = None
    if world == 1 and not args.no_cpu_baseline:
        base, _ = cpu_baseline(12, 1)  # ~10-20 s of host time
    ref_gpu = None
    if world == 1 and not args.no_ref_gpu:
        ref_gpu = reference_gpu(sc, args.settle, min(args.steps, 10))
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"sand3 (examples3d/sand3.rs scaled): 2 stacked lattices {dims[0]}x{dims[1]}x{dims[2]} per GPU = {npart} particles/GPU, h=0.2, "
                                   "Drucker-Prager sand over corotated block, cuboid floor+walls", "particles_per_gpu": npart,
                       "parallelism": f"x-slabs x{world}, halo accumulate over {'NVLink peer memory (device-initiated)' if args.halo == 'nvlink' else 'NCCL all-gather'}" if world > 1 else "single GPU", "l2": "inputs larger than L2 (particle state ~0.8 GB per GPU)"},
            "clocks": clocks, "gpu_launches": int(launches) if launches else None, "e2e": e2e, "roofline": roof, "cpu_baseline": base,
            "reference_gpu": ref_gpu, "frame_loop": frame_loop}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _ncu_traffic():
    """dram bytes per g2p2g launch from the committed ncu summary of the same command (profiles/), if present"""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "g2p2g_ncu_summary.json")))
        return float(d["dram_bytes_per_launch"])
    except Exception:
        return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--dims", type=int, nargs=3, default=None, help="per-GPU lattice (default 200 100 100)")
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
