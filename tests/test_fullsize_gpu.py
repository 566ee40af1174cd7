"""BASELINE.json's configurations at FULL size on the B200, checked through size-independent properties (SURVEY.md §8c:
the oracle cannot follow at these sizes):

  * sort: the sorted ids are a permutation; every particle sits in the block its position maps to (IEEE division +
    round-half-away recomputed on the host: bit-exact membership); the active set is exactly the particles' blocks plus
    their {0,1}^3 neighbours; headers are an exclusive scan; sorting twice gives the same canonical state;
  * P2G: quadratic B-splines are a partition of unity with linear precision, so the grid's total mass equals the
    particles' mass and the grid's total momentum equals sum(m v) (the APIC / stress terms cancel exactly);
  * update: particle count, masses, rest volumes and model indices are invariant, everything stays finite, kinematic
    particles advance by exactly kinematic_vel * dt, dt obeys min_dt <= dt <= max_dt.
"""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes

pytestmark = pytest.mark.gpu

CONFIGS = {  # BASELINE.json configs[1..4]; N256M is 8 slabs of the N-H block (one slab here, scaled to 8 M to bound host RAM)
    "F1M_fluids3": lambda: scenes.fluids3(160, 80, 80),
    "S4M_sand3": lambda: scenes.sand3(200, 100, 100),
    "C16M_cube_through_sand3": lambda: scenes.cube_through_sand3(400, 200, 200, 100),
    "N8M_neo_hookean_slab": lambda: scenes.neo_hookean_block(200, 200, 200),
}


def _block_keys(points, h):
    q = points.astype(np.float32) / np.float32(h)                                    # div.rn.f32
    q64 = q.astype(np.float64)
    cell = (np.sign(q64) * np.floor(np.abs(q64) + 0.5)).astype(np.int64)             # round half away from zero
    blk = (cell - 2 + (1 << 22)) // 4                                                # gpu_grid.rs:271-279
    return blk


def _pack(b):
    m = (1 << 21) - 1
    return (b[..., 0] & m) | ((b[..., 1] & m) << 21) | ((b[..., 2] & m) << 42)


@pytest.mark.parametrize("name", list(CONFIGS))
def test_full_size_properties(name):
    from sparkl_b200.pipeline import context_from_scene
    sc = CONFIGS[name]()
    n = sc.num_particles
    h = sc.cell_width
    ctx = context_from_scene(sc)
    p = sc.params
    mass = sc.particles.volume["mass"].astype(np.float64)
    ctx.frame_begin(p)
    remaining = np.float32(p.dt)
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    for s in range(3):
        info = ctx.stage_sort()
        hdr = ctx.grid_headers(1)
        ids = ctx.read(abi.ATTR_SORTED_IDS)
        pos = ctx.read(abi.ATTR_POSITION)["point"]
        # ---- sort invariants
        assert len(hdr) == info.num_active_blocks
        seen = np.zeros(n, np.uint8)
        seen[ids] = 1
        assert len(ids) == n and seen.all(), "sorted ids are not a permutation"
        assert int(hdr["num_particles"].sum()) == n
        assert np.array_equal(hdr["first_particle"], np.concatenate([[0], np.cumsum(hdr["num_particles"])[:-1]])), "headers are not an exclusive scan"
        key_of_particle = _pack(_block_keys(pos, h)).astype(np.uint64)
        block_of_slot = np.repeat(hdr["virtual_id"], hdr["num_particles"])
        assert np.array_equal(key_of_particle[ids], block_of_slot), "a particle is not in the block its position maps to"
        okeys = np.unique(key_of_particle).astype(np.int64)
        m21 = (1 << 21) - 1
        occupied = np.stack([okeys & m21, (okeys >> 21) & m21, (okeys >> 42) & m21], axis=1)
        shifts = np.array([[i, j, k] for k in (0, 1) for j in (0, 1) for i in (0, 1)], np.int64)
        expect = np.unique(_pack((occupied[:, None, :] + shifts[None, :, :]).reshape(-1, 3)).astype(np.uint64))
        assert np.array_equal(np.sort(hdr["virtual_id"]), expect), "active block set != particle blocks + {0,1}^3"
        if s == 0:  # idempotence
            info2 = ctx.stage_sort()
            hdr2 = ctx.grid_headers(1)
            o1, o2 = np.argsort(hdr["virtual_id"]), np.argsort(hdr2["virtual_id"])
            assert info2.num_active_blocks == info.num_active_blocks and info2.num_dispatch_blocks == info.num_dispatch_blocks
            assert np.array_equal(hdr["virtual_id"][o1], hdr2["virtual_id"][o2]) and np.array_equal(hdr["num_particles"][o1], hdr2["num_particles"][o2])
        # ---- dt rules
        ctx.stage_reset_grid()
        max_dt = float(min(remaining, np.float32(p.max_substep_dt)))
        dt = ctx.stage_estimate_dt(min_dt, max_dt)
        assert 0.0 < dt <= max_dt * (1 + 1e-6) and dt >= min(min_dt, max_dt) * (1 - 1e-6)
        # ---- P2G conservation
        pos_before = pos
        ctx.stage_g2p2g(dt)
        nodes = ctx.grid_nodes(1)
        vel = ctx.read(abi.ATTR_VELOCITY)["vector"].astype(np.float64)
        st = ctx.read(abi.ATTR_STATUS)
        alive = st["failed"] == 0
        pm = float(mass[alive].sum())
        gm = float(nodes["mass"].astype(np.float64).sum())
        assert abs(gm - pm) <= 2e-5 * pm, f"grid mass {gm} vs particle mass {pm}"
        mom_p = (mass[alive, None] * vel[alive]).sum(axis=0)
        mom_g = nodes["momentum_velocity"].astype(np.float64).reshape(-1, 3).sum(axis=0)
        scale = float((mass[alive, None] * np.abs(vel[alive])).sum()) + 1e-3 * pm
        assert np.abs(mom_g - mom_p).max() <= 1e-4 * scale, f"grid momentum {mom_g} vs particles {mom_p}"
        # ---- update invariants
        pos_after = ctx.read(abi.ATTR_POSITION)["point"]
        assert np.isfinite(pos_after).all() and np.isfinite(vel).all()
        vol = ctx.read(abi.ATTR_VOLUME)
        assert np.array_equal(vol["mass"], sc.particles.volume["mass"]) and np.array_equal(vol["volume0"], sc.particles.volume["volume0"])
        assert np.array_equal(st["model_index"], sc.particles.status["model_index"])
        assert np.isfinite(vol["deformation_gradient"]).all()
        kin = sc.particles.status["kinematic_vel_enabled"] != 0
        if kin.any():
            step = (sc.particles.status["kinematic_vel"][kin] * np.float32(dt)).astype(np.float32)
            want = (pos_before[kin] + step).astype(np.float32)  # the device may contract the multiply-add: 1 ulp
            assert np.abs(pos_after[kin].astype(np.float64) - want).max() <= 2e-7 * max(1.0, float(np.abs(want).max())), "kinematic particles must advance by kinematic_vel * dt"
        assert float(np.abs(pos_after - pos_before).max()) <= h * (1 + 1e-6), "CFL clamp: nobody moves more than one cell per substep"
        ctx.stage_grid_update(dt, sc.gravity)
        remaining = np.float32(remaining - np.float32(dt))
        if remaining <= 1e-11:  # the frame is over (a fluid at rest takes dt = 1/60 in one substep): start the next one
            ctx.frame_begin(p)
            remaining = np.float32(p.dt)
    assert len(ctx) == n
