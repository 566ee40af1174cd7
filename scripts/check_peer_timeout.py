"""torchrun check (two ranks, may share one GPU: SPK_CHECK_ONE_GPU=1): a rank that reaches a device-side barrier more than
SPK_DIST_TIMEOUT_MS late must surface as SPK_EPEER on the rank that waited — not as a silently diverged run."""
import os, sys, time
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene
from sparkl_b200.distributed import NvlinkMpmPipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
one_gpu = os.environ.get("SPK_CHECK_ONE_GPU") == "1"
if one_gpu:
    local = 0
torch.cuda.set_device(local)
dist.init_process_group("gloo" if one_gpu else "nccl", **({} if one_gpu else {"device_id": torch.device("cuda", local)}))
sc = scenes.jitter(scenes.sand3(16, 8, 8), 0.02, seed=9)
x = sc.particles.position["point"][:, 0]
edges = np.linspace(x.min(), x.max() + 1e-3, world + 1)
mine = np.nonzero((x >= edges[rank]) & (x < edges[rank + 1]))[0]
part = scenes.Scene(sc.name, sc.cell_width, sc.particles.take(mine), sc.models, sc.colliders, sc.gravity, sc.params)
ctx = context_from_scene(part, device=local)
pipe = NvlinkMpmPipeline(ctx, dist)
pipe.step(sc.params, sc.gravity)          # one frame in step: the path works
dist.barrier()
if rank == 1:
    time.sleep(float(os.environ.get("SPK_CHECK_DELAY_S", "2.0")))  # arrives long after the time-out of rank 0's barrier
code = 0
try:
    for _ in range(3):
        pipe.step(sc.params, sc.gravity)
    ctx.synchronize()
except abi.SparklError as e:
    code = e.code
print(f"peer timeout check: rank {rank} error code {code}", "OK" if (rank != 0 or code == abi.EPEER) else "FAIL", flush=True)
dist.barrier()
dist.destroy_process_group()
