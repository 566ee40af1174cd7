"""dev aid: per-substep wall-time distribution of the NVLink multi-GPU path (torchrun)"""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sparkl_b200 import abi
from sparkl_b200.pipeline import context_from_scene
from sparkl_b200.distributed import NvlinkMpmPipeline
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = bench.build_scene(rank, world, bench.FULL)
ctx = context_from_scene(sc, device=local)
if world > 1: NvlinkMpmPipeline(ctx, dist)
p = abi.SolverParameters(stop_after_one_substep=1)
tm = int(os.environ.get("TIMINGS", "1"))
for _ in range(130): ctx.step(p, sc.gravity, timings=bool(tm))
dist.barrier(); torch.cuda.synchronize()
ts = []
for _ in range(300):
    t0 = time.perf_counter(); ctx.step(p, sc.gravity, timings=bool(tm)); ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"rank {rank}: timings={tm} mean {ts.mean():.3f} ms  p50 {np.percentile(ts,50):.3f} p90 {np.percentile(ts,90):.3f} p99 {np.percentile(ts,99):.3f} max {ts.max():.3f}; slow (>1.5x median): {(ts > 1.5*np.median(ts)).sum()}", flush=True)
dist.destroy_process_group()
