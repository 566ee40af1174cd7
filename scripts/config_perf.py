"""per-config throughput of the BASELINE.json configurations (development aid; bench.py is the contract).
   python scripts/config_perf.py fluids3|sand3|cube_through_sand3|neo_hookean  [substeps]     (torchrun for N > 1: x-slabs)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene
name = sys.argv[1]
nsub = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
import torch
torch.cuda.set_device(local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
t0 = time.time()
if name == "fluids3": sc = scenes.fluids3(160, 80, 80)                                  # F1M
elif name == "sand3": sc = scenes.sand3(200, 100, 100)                                  # S4M
elif name == "cube_through_sand3": sc = scenes.cube_through_sand3(400, 200, 200, 100)   # C16M
elif name == "neo_hookean": sc = scenes.neo_hookean_block(400, 400, 200, x_offset=rank * 400 * 0.1)  # N256M: 32 M per GPU, 8 slabs
else: raise SystemExit(name)
ctx = context_from_scene(sc, device=local)
build_s = time.time() - t0
if world > 1:
    from sparkl_b200.distributed import NvlinkMpmPipeline
    NvlinkMpmPipeline(ctx, dist)
p = abi.SolverParameters(stop_after_one_substep=1, max_substep_dt=sc.params.max_substep_dt)  # the N-H scene caps the substep
for _ in range(10): ctx.step(p, sc.gravity, timings=False)
torch.cuda.synchronize()
if world > 1: dist.barrier()
acc = np.zeros(5); tot = 0.0; t0 = time.perf_counter()
for _ in range(nsub):
    _, t = ctx.step(p, sc.gravity, timings=(world == 1))
    if world == 1:
        acc += [t.grid_resize_and_sort_ms, t.reset_grid_ms, t.estimate_timestep_ms, t.g2p2g_ms, t.grid_update_ms]; tot += t.total_ms
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) / nsub * 1e3
if world > 1:
    tt = torch.tensor([wall], dtype=torch.float64, device="cuda"); dist.all_reduce(tt, op=dist.ReduceOp.MAX); wall = float(tt.item())
if rank == 0:
    n = sc.num_particles
    print(f"{name}: {n} particles/GPU x {world} GPU, scene+upload {build_s:.1f}s | wall {wall:.3f} ms/substep -> {n * world / wall / 1e6:.2f} G particle-substeps/s"
          + (f" | stages ms: sort {acc[0]/nsub:.3f} reset {acc[1]/nsub:.3f} dt {acc[2]/nsub:.3f} g2p2g {acc[3]/nsub:.3f} grid {acc[4]/nsub:.3f}; g2p2g {n/(acc[3]/nsub)/1e6:.2f} G/s" if world == 1 else ""), flush=True)
if world > 1: dist.destroy_process_group()
