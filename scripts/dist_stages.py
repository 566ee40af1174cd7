"""dev aid: per-stage event timings of the NVLink multi-GPU path (torchrun), rank 0"""
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sparkl_b200 import abi
from sparkl_b200.pipeline import context_from_scene
from sparkl_b200.distributed import NvlinkMpmPipeline
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc, _ = bench.build_scene(rank, world, bench.FULL)
ctx = context_from_scene(sc, device=local)
if world > 1: NvlinkMpmPipeline(ctx, dist)
p = abi.SolverParameters(stop_after_one_substep=1)
for _ in range(130): ctx.step(p, sc.gravity, timings=False)
acc = np.zeros(6)
for _ in range(50):
    _, t = ctx.step(p, sc.gravity)
    acc += [t.grid_resize_and_sort_ms, t.reset_grid_ms, t.estimate_timestep_ms, t.g2p2g_ms, t.grid_update_ms, t.total_ms]
print(f"rank {rank}/{world}: sort {acc[0]/50:.3f} reset {acc[1]/50:.3f} dt+S1+detect {acc[2]/50:.3f} g2p2g(+push) {acc[3]/50:.3f} S2+merge+grid_update {acc[4]/50:.3f} total {acc[5]/50:.3f} ms", flush=True)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
stream = torch.cuda.ExternalStream(ctx.cuda_stream(), device=torch.device("cuda", local))
with torch.cuda.stream(stream):
    e0.record(stream)
    for _ in range(50): ctx.step(p, sc.gravity, timings=False)
    e1.record(stream)
torch.cuda.synchronize()
print(f"rank {rank}/{world}: untimed loop {e0.elapsed_time(e1)/50:.3f} ms/substep", flush=True)
dist.barrier()
dist.destroy_process_group()
