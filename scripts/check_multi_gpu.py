"""torchrun check: N ranks (NCCL) each owning an x-slab must reproduce the single-GPU result (run under gpurun --gpus N)."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene
from sparkl_b200.distributed import DistributedMpmPipeline, NvlinkMpmPipeline

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
# SPK_CHECK_ONE_GPU=1: every rank on cuda:0 (time-sliced), gloo for the descriptor exchange — lets the single-GPU test tier
# exercise the peer-memory path (tests/test_nvlink_gpu.py)
one_gpu = os.environ.get("SPK_CHECK_ONE_GPU") == "1"
if one_gpu:
    local = 0
torch.cuda.set_device(local)
if one_gpu:
    dist.init_process_group("gloo")
else:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sc = scenes.jitter(scenes.sand3(24, 8, 8), 0.02, seed=9)
x = sc.particles.position["point"][:, 0]
edges = np.linspace(x.min(), x.max() + 1e-3, world + 1)
if os.environ.get("SPK_CHECK_EMPTY_RANK") == "1":  # the last rank owns nothing: it still has to walk through every barrier of every substep
    edges = np.concatenate([np.linspace(x.min(), x.max() + 1e-3, world), [x.max() + 1.0]])
mine = np.nonzero((x >= edges[rank]) & (x < edges[rank + 1]))[0]
part = scenes.Scene(sc.name, sc.cell_width, sc.particles.take(mine), sc.models, sc.colliders, sc.gravity, sc.params)
ctx = context_from_scene(part, device=local)
mode = sys.argv[1] if len(sys.argv) > 1 else "nvlink"
nsub = 0
if mode == "nvlink":
    pipe = NvlinkMpmPipeline(ctx, dist)
    for _ in range(2):
        sim, t = pipe.step(sc.params, sc.gravity)
        nsub += t.num_substeps
    pipe.halo_blocks_last = pipe.halo_blocks()
else:
    pipe = DistributedMpmPipeline(ctx, dist, torch.device("cuda", local))
    for _ in range(2):
        sim, n = pipe.step(sc.params, sc.gravity)
        nsub += n
dev = "cpu" if one_gpu else "cuda"
pos = torch.zeros((len(sc.particles), 3), dtype=torch.float32, device=dev)
pos[torch.from_numpy(mine).to(dev)] = torch.from_numpy(ctx.read(abi.ATTR_POSITION)["point"].copy()).to(dev)
dist.all_reduce(pos)
if rank == 0:
    ref = context_from_scene(sc, device=local)
    nref = 0
    for _ in range(2):
        sim, t = ref.step(sc.params, sc.gravity)
        nref += t.num_substeps
    want = ref.read(abi.ATTR_POSITION)["point"]
    err = np.abs(pos.cpu().numpy() - want).max() / sc.cell_width
    print(f"multi-gpu check [{mode}]: world={world} substeps {nsub} vs {nref}, halo blocks {pipe.halo_blocks_last}, max |dx|/h = {err:.3e}", "OK" if err < 1e-4 and nsub == nref else "FAIL", flush=True)
dist.destroy_process_group()
