"""torchrun check of particle migration on the product path: N ranks (NVLink peer-memory halo path inside spk_step) own x-slabs, a
kinematic block is pushed through the slab boundaries, `distributed.migrate` re-homes the particles after every frame, and the
result must equal the single-GPU run.  Run under `gpurun --gpus N`, or on one GPU with SPK_CHECK_ONE_GPU=1 (all ranks share cuda:0,
gloo for the host exchange; tests/test_nvlink_gpu.py).  SPK_MIGRATE=device uses the device-side path (`distributed.migrate_device`:
spk_particles_migrate_out / _in, global ids kept in the context) instead of the host-orchestrated one."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparkl_b200 import abi, scenes  # noqa: E402
from sparkl_b200.distributed import NvlinkMpmPipeline, SlabPartition, migrate, migrate_device  # noqa: E402
from sparkl_b200.pipeline import context_from_scene  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
one_gpu = os.environ.get("SPK_CHECK_ONE_GPU") == "1"
if one_gpu:
    local = 0
torch.cuda.set_device(local)
dist.init_process_group("gloo") if one_gpu else dist.init_process_group("nccl", device_id=torch.device("cuda", local))
xdev = torch.device("cpu") if one_gpu else torch.device("cuda", local)  # where the exchanged records live (gloo: host, NCCL: device)

sc = scenes.cube_through_sand3(16, 6, 6, 4)
sc.params = abi.SolverParameters(max_substep_dt=2.0e-3)
x = sc.particles.position["point"][:, 0]
part = SlabPartition.uniform(float(x.min()), float(x.max()), world, sc.cell_width)
owner = part.owner(x, sc.cell_width)
gids = np.nonzero(owner == rank)[0].astype(np.uint32)
assert len(gids) > 0, "every rank must own particles"
ctx = context_from_scene(scenes.Scene(sc.name, sc.cell_width, sc.particles.take(gids), sc.models, sc.colliders, sc.gravity, sc.params), device=local)
device_side = os.environ.get("SPK_MIGRATE") == "device"
if device_side:
    ctx.set_gids(gids)
pipe = NvlinkMpmPipeline(ctx, dist)
frames, moved, nsub = 10, 0, 0  # the kinematic block (10 m/s) travels 1.67 m: across the first slab boundary
for _ in range(frames):
    sim, t = pipe.step(sc.params, sc.gravity)
    nsub += t.num_substeps
    if device_side:
        sent, recv = migrate_device(ctx, part, dist, torch.device("cuda", local), wire=xdev)
        gids = ctx.read_gids()
    else:
        gids, sent, recv = migrate(ctx, gids, part, dist, xdev)
    moved += sent
    assert (part.owner(ctx.read(abi.ATTR_POSITION)["point"][:, 0], sc.cell_width) == rank).all()
pos = torch.zeros((len(sc.particles), 3), dtype=torch.float32, device=xdev)
pos[torch.from_numpy(gids.astype(np.int64)).to(xdev)] = torch.from_numpy(ctx.read(abi.ATTR_POSITION)["point"].copy()).to(xdev)
cnt = torch.tensor([moved, len(gids)], dtype=torch.int64, device=xdev)
dist.all_reduce(pos)
dist.all_reduce(cnt)
if rank == 0:
    ref = context_from_scene(sc, device=local)
    nref = 0
    for _ in range(frames):
        sim, t = ref.step(sc.params, sc.gravity)
        nref += t.num_substeps
    err = np.abs(pos.cpu().numpy() - ref.read(abi.ATTR_POSITION)["point"]).max() / sc.cell_width
    ok = err < 1e-3 and nsub == nref and int(cnt[0]) > 0 and int(cnt[1]) == len(sc.particles)
    print(f"migration check ({'device-side' if device_side else 'host-orchestrated'}): world={world} substeps {nsub} vs {nref}, particles moved {int(cnt[0])}, total {int(cnt[1])}/{len(sc.particles)}, "
          f"max |dx|/h = {err:.3e}", "OK" if ok else "FAIL", flush=True)
dist.destroy_process_group()
