"""world_size-2 test of the multi-process pipeline (sparkl_b200/distributed.py) on CPU: gloo backend, the oracle as
the per-rank engine.  Two ranks owning the two halves of the particle set must reproduce the single-rank result:
the halo scheme makes the outcome independent of particle ownership."""
import os
import socket

import numpy as np
import pytest

from sparkl_b200 import abi, scenes
from tests import util


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, nsub, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as ob
        from sparkl_b200.distributed import DistributedMpmPipeline
        ob.set_threads(2)
        sc = scenes.jitter(scenes.sand3(12, 5, 6), 0.02, seed=9)
        x = sc.particles.position["point"][:, 0]
        mine = np.nonzero((x < 0.6) if rank == 0 else (x >= 0.6))[0]
        e = ob.OracleEngine(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders)
        e.upload(sc.particles.take(mine))
        pipe = DistributedMpmPipeline(e, dist, torch.device("cpu"))
        e.frame_begin(sc.params)
        remaining = np.float32(sc.params.dt)
        halo = 0
        for _ in range(nsub):
            dt = pipe.substep(sc.params, sc.gravity, float(remaining))
            halo = max(halo, pipe.halo_blocks_last)
            remaining = np.float32(remaining - np.float32(dt))
        pc = e.read_all()
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), ids=mine, pos=pc.position, vel=pc.velocity, vol=pc.volume, data=pc.data, halo=halo,
                 dt_acc=e.dt_acc())
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_match_single_rank(tmp_path, oracle):
    import torch.multiprocessing as mp
    nsub = 4
    port = _free_port()
    mp.spawn(_worker, args=(2, port, nsub, str(tmp_path)), nprocs=2, join=True)
    sc = scenes.jitter(scenes.sand3(12, 5, 6), 0.02, seed=9)
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    e.frame_begin(sc.params)
    remaining = np.float32(sc.params.dt)
    for _ in range(nsub):
        dt = e.substep(sc.params, sc.gravity, float(remaining))
        remaining = np.float32(remaining - np.float32(dt))
    ref = e.read_all()
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(2)]
    assert r[0]["halo"] > 0 and r[1]["halo"] > 0                     # the slabs do share blocks
    assert int(r[0]["dt_acc"]) == int(r[1]["dt_acc"]) == e.dt_acc()  # global dt minimum
    got = abi.ParticleComponents(len(ref))
    for k in range(2):
        ids = r[k]["ids"]
        got.position[ids], got.velocity[ids], got.volume[ids], got.data[ids] = r[k]["pos"], r[k]["vel"], r[k]["vol"], r[k]["data"]
    got.status, got.phase = ref.status, ref.phase
    assert util.rel_err(got.position["point"], ref.position["point"]) < 1e-6
    assert util.rel_err(got.velocity["vector"], ref.velocity["vector"], floor=1e-2) < 2e-4
    assert util.rel_err(got.volume["deformation_gradient"], ref.volume["deformation_gradient"], floor=1.0) < 1e-5
    assert util.rel_err(got.data["plastic_hardening"], ref.data["plastic_hardening"], floor=1.0) < 1e-5


# ---------------------------------------------------------------------------------------------- particle migration
def test_slab_partition_owner():
    from sparkl_b200.distributed import SlabPartition
    part = SlabPartition([-100, 0, 3, 100])
    h = 0.2
    x = np.array([-5.0, -0.31, -0.29, 0.49, 0.51, 2.69, 2.71, 9.0], np.float32)  # block x = floor((round(x / h) - 2) / 4)
    assert list(SlabPartition.block_x(x, h)) == [-7, -1, -1, 0, 0, 2, 3, 10]
    assert list(part.owner(x, h)) == [0, 0, 0, 1, 1, 1, 2, 2]
    u = SlabPartition.uniform(0.0, 8.0, 4, h)
    assert u.world == 4 and (np.diff(u.bounds) > 0).all()


def _migration_worker(rank, world, port, nframes, out_dir):
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import binding as ob
        from sparkl_b200.distributed import DistributedMpmPipeline, SlabPartition, migrate
        ob.set_threads(2)
        sc = _migration_scene()
        part = SlabPartition([-(1 << 40), -1, 1 << 40])  # rank 0: block x < -1 (x < -0.5: the kinematic block at start), rank 1: the sand
        owner = part.owner(sc.particles.position["point"][:, 0], sc.cell_width)
        gids = np.nonzero(owner == rank)[0].astype(np.uint32)
        e = ob.OracleEngine(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders)
        e.upload(sc.particles.take(gids))
        pipe = DistributedMpmPipeline(e, dist, torch.device("cpu"))
        moved = 0
        for _ in range(nframes):
            pipe.step(sc.params, sc.gravity)
            gids, sent, recv = migrate(e, gids, part, dist, torch.device("cpu"))
            moved += sent
            now = part.owner(e.read(abi.ATTR_POSITION)["point"][:, 0], sc.cell_width)
            assert (now == rank).all()  # after a migration every local particle is inside the rank's slab
        pc = e.read_all()
        np.savez(os.path.join(out_dir, f"mig{rank}.npz"), ids=gids, pos=pc.position, vel=pc.velocity, vol=pc.volume, moved=moved)
    finally:
        dist.destroy_process_group()


def _migration_scene():
    """a kinematic block pushed in +x through the slab boundary (kinematic particles move at 10 m/s: 1/6 m per frame)"""
    sc = scenes.cube_through_sand3(10, 5, 6, 4)
    sc.params = abi.SolverParameters(max_substep_dt=2.0e-3)
    return sc


@pytest.mark.timeout(300)
def test_migration_keeps_results_and_rehomes_particles(tmp_path, oracle):
    import torch.multiprocessing as mp
    nframes = 4
    port = _free_port()
    mp.spawn(_migration_worker, args=(2, port, nframes, str(tmp_path)), nprocs=2, join=True)
    sc = _migration_scene()
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    for _ in range(nframes):
        e.step(sc.params, sc.gravity)
    ref = e.read_all()
    r = [np.load(tmp_path / f"mig{k}.npz") for k in range(2)]
    assert int(r[0]["moved"]) + int(r[1]["moved"]) > 0                        # particles did change rank
    ids = np.concatenate([r[0]["ids"], r[1]["ids"]])
    assert np.array_equal(np.sort(ids), np.arange(len(ref)))                   # nobody lost, nobody duplicated
    pos = np.zeros(len(ref), abi.POSITION); vel = np.zeros(len(ref), abi.VELOCITY); vol = np.zeros(len(ref), abi.VOLUME)
    for k in range(2):
        pos[r[k]["ids"]], vel[r[k]["ids"]], vol[r[k]["ids"]] = r[k]["pos"], r[k]["vel"], r[k]["vol"]
    # ownership does not change the physics: same state as the single-rank run up to summation order
    assert util.rel_err(pos["point"], ref.position["point"]) < 2e-5
    assert util.rel_err(vel["vector"], ref.velocity["vector"], floor=1e-2) < 2e-3
    assert util.rel_err(vol["deformation_gradient"], ref.volume["deformation_gradient"], floor=1.0) < 2e-4


def test_balanced_slab_partition_splits_by_particle_count():
    """`SlabPartition.balanced` (bench.py --config C16M): slab edges at the quantiles of the particles' block columns — every rank owns
    about the same number of particles even when the cloud is lopsided, edges strictly increase, ownership covers every particle"""
    from sparkl_b200.distributed import SlabPartition
    rng = np.random.Generator(np.random.PCG64(3))
    x = np.concatenate([rng.uniform(-9.0, -8.0, 2000), rng.uniform(0.0, 40.0, 98000)]).astype(np.float32)  # a small far cluster + the bulk
    for world in (2, 4, 8):
        part = SlabPartition.balanced(x, world, 0.2)
        assert part.world == world and (np.diff(part.bounds) > 0).all()
        owner = part.owner(x, 0.2)
        counts = np.bincount(owner, minlength=world)
        assert counts.sum() == len(x) and counts.min() > 0.7 * len(x) / world and counts.max() < 1.3 * len(x) / world, counts
        uni = SlabPartition.uniform(float(x.min()), float(x.max()), world, 0.2)
        assert np.bincount(uni.owner(x, 0.2), minlength=world).max() >= counts.max()
    # degenerate: fewer occupied block columns than ranks still yields strictly increasing edges
    part = SlabPartition.balanced(np.full(100, 1.0, np.float32), 4, 0.2)
    assert (np.diff(part.bounds) > 0).all() and (part.owner(np.full(100, 1.0, np.float32), 0.2) < 4).all()
