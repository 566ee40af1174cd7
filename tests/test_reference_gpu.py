"""Three-way parity on the B200 box: the reference's OWN GPU kernels (oracle/_ref: its shipped PTX assembled for
sm_100a, driven by oracle/ref_harness.cpp) vs the CPU oracle vs the CUDA product.

This is what pins the CPU oracle (the reference has no golden vectors for this path) and what backs the
product's parity claim against the reference itself.  The reference is racy (block ids and in-block order come from
atomics, grid sums are float `red.add`s), so integer structures are compared on canonical form and floating-point state
within a tolerance (tests/util.py:Tol)."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes
from tests import util

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref (reference kernels) not built on this box")
    return oracle


def _three(scene, ob):
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(scene)
    engines = {"product": g, "oracle": ob.OracleEngine(scene.cell_width), "reference": ob.RefEngine(scene.cell_width)}
    for k in ("oracle", "reference"):
        engines[k].models_upload(scene.models)
        engines[k].colliders_upload(scene.colliders)
        engines[k].upload(scene.particles)
    return engines


def _lockstep(scene, ob, nsub, base, growth=3.0):
    E = _three(scene, ob)
    p = scene.params
    for e in E.values():
        e.frame_begin(p)
    remaining = np.float32(p.dt)
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    worst = {}
    for s in range(nsub):
        tol = util.scene_tol(scene, base * (1 + growth * s))
        max_dt = float(min(remaining, np.float32(p.max_substep_dt)))
        canon = {}
        for k, e in E.items():
            e.stage_sort()
            canon[k] = util.canonical_sort_state(e)
        # membership is exact unless a particle sits on a rounding boundary of round(p / h) within the position
        # tolerance; the scenes are jittered so that this does not happen
        pos_ref = E["reference"].read(abi.ATTR_POSITION)["point"].astype(np.float64) / scene.cell_width
        on_boundary = (np.abs(np.abs(pos_ref - np.floor(pos_ref)) - 0.5) < 1e-5).any(axis=1)
        assert not on_boundary.any(), f"substep {s}: {int(on_boundary.sum())} particles sit on a cell boundary; membership is ill-defined there"
        for k in ("product", "oracle"):
            assert np.array_equal(canon[k][0], canon["reference"][0]), f"substep {s}: active blocks of {k} differ from the reference"
            assert canon[k][2] == canon["reference"][2], f"substep {s}: dispatch multiset of {k} differs from the reference"
            for vid, ids in canon["reference"][1].items():
                assert np.array_equal(canon[k][1][vid], ids), f"substep {s}: block {vid:#x} particle set of {k} differs"
        for e in E.values():
            e.stage_reset_grid()
        dts = {k: e.stage_estimate_dt(min_dt, max_dt) for k, e in E.items()}
        accs = {k: e.dt_acc() for k, e in E.items()}
        for k in ("product", "oracle"):
            assert abs(accs[k] - accs["reference"]) <= max(2, int(1e-5 * accs["reference"])), f"substep {s}: dt accumulator {k}={accs[k]} reference={accs['reference']}"
        dt = dts["reference"]
        for e in E.values():
            e.set_dt_acc(accs["reference"])
            e.stage_g2p2g(dt)
        pr = E["reference"].read_all()
        for k in ("product", "oracle"):
            errs = util.assert_particles_close(E[k].read_all(), pr, scene.cell_width, tol, f"substep {s} {k} vs reference, particles after g2p2g")
            for f, v in errs.items():
                worst[(k, f)] = max(worst.get((k, f), 0.0), v)
        kr, nr = util.nodes_by_key(E["reference"], 1)
        for k in ("product", "oracle"):
            kk, nn = util.nodes_by_key(E[k], 1)
            assert np.array_equal(kk, kr)
            util.assert_nodes_close(nn, nr, tol, f"substep {s} {k} vs reference, next grid after g2p2g", scene, kr, ob)
        for e in E.values():
            e.stage_grid_update(dt, scene.gravity)
        kr, nr = util.nodes_by_key(E["reference"], 0)
        for k in ("product", "oracle"):
            kk, nn = util.nodes_by_key(E[k], 0)
            util.assert_nodes_close(nn, nr, tol, f"substep {s} {k} vs reference, grid after grid_update", scene, kr, ob)
        remaining = np.float32(remaining - np.float32(dt))
        if remaining <= 1e-11:
            for e in E.values():
                e.frame_begin(p)
            remaining = np.float32(p.dt)
    return worst


def test_sand_against_reference_kernels(ref):
    worst = _lockstep(scenes.jitter(scenes.sand3(12, 6, 10), 0.02, seed=7), ref, 12, 2e-5)
    print("worst relative deviations vs the reference kernels:", {f"{k[0]}.{k[1]}": f"{v:.2e}" for k, v in worst.items()})


def test_fluid_against_reference_kernels(ref):
    _lockstep(scenes.jitter(scenes.fluids3(14, 10, 10), 0.01, seed=2), ref, 8, 5e-5)


def test_neo_hookean_against_reference_kernels(ref):
    _lockstep(scenes.jitter(scenes.neo_hookean_block(12, 12, 8), 0.02, seed=4), ref, 8, 2e-5)


def test_kinematic_block_against_reference_kernels(ref):
    _lockstep(scenes.jitter(scenes.cube_through_sand3(12, 6, 6, 4), 0.02, seed=6), ref, 8, 2e-5)


def test_reference_run_to_run_spread(ref):
    """the tolerance protocol of SURVEY §8c: the reference against ITSELF (its atomics make it non-deterministic)"""
    sc = scenes.jitter(scenes.sand3(12, 6, 10), 0.02, seed=7)
    runs = []
    for _ in range(2):
        e = ref.RefEngine(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
        for _ in range(3):
            e.step(sc.params, sc.gravity)
        runs.append(e.read_all())
    spread = util.rel_err(runs[0].velocity["vector"], runs[1].velocity["vector"], floor=1e-2)
    print("reference run-to-run spread of particle velocities after 3 frames:", spread)
    util.assert_particles_close(runs[0], runs[1], sc.cell_width, util.scene_tol(sc, 2e-3), "reference vs itself")


def test_whole_frames_product_vs_reference(ref):
    sc = scenes.jitter(scenes.sand3(12, 6, 10), 0.02, seed=7)
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(sc)
    r = ref.RefEngine(sc.cell_width)
    r.models_upload(sc.models); r.colliders_upload(sc.colliders); r.upload(sc.particles)
    for frame in range(3):
        sg, tg = g.step(sc.params, sc.gravity)
        sr, tr = r.step(sc.params, sc.gravity)
        assert tg.num_substeps == tr.num_substeps and abs(sg - sr) < 1e-6
    util.assert_particles_close(g.read_all(), r.read_all(), sc.cell_width, util.scene_tol(sc, 2e-3), "3 frames product vs reference")
