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
    if scene.particles.position.dtype["point"].shape[0] == 2:  # dim2: liboracle2d.so and the reference's sparkl2d kernels
        engines = {"product": g, "oracle": ob.OracleEngine2D(scene.cell_width), "reference": ob.RefEngine2D(scene.cell_width)}
    else:
        engines = {"product": g, "oracle": ob.OracleEngine(scene.cell_width), "reference": ob.RefEngine(scene.cell_width)}
    for k in ("oracle", "reference"):
        engines[k].models_upload(scene.models)
        engines[k].colliders_upload(scene.colliders)
        engines[k].upload(scene.particles)
    return engines


def _lockstep(scene, ob, nsub, base, growth=3.0, robust=False, dt_rtol=1e-5):
    E = _three(scene, ob)
    p = scene.params
    for e in E.values():
        e.frame_begin(p)
    remaining = np.float32(p.dt)
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    worst = {}
    for s in range(nsub):
        tol = util.scene_tol(scene, base * (1 + growth * s))
        if any(getattr(m.plastic_model, "tag", None) == 0 for m in scene.models):
            tol.few = 2e-3  # Drucker-Prager: see util.Tol.few
        max_dt = float(min(remaining, np.float32(p.max_substep_dt)))
        canon = {}
        for k, e in E.items():
            e.stage_sort()
            canon[k] = util.canonical_sort_state(e)
        # membership is exact unless a particle sits on a rounding boundary of round(p / h) within the position
        # tolerance
        boundary = util.boundary_particles(E["reference"].read(abi.ATTR_POSITION)["point"], scene.cell_width)
        for k in ("product", "oracle"):
            util.assert_membership_equal(canon[k], canon["reference"], boundary, f"substep {s}: {k} vs reference")
        for e in E.values():
            e.stage_reset_grid()
        dts = {k: e.stage_estimate_dt(min_dt, max_dt) for k, e in E.items()}
        accs = {k: e.dt_acc() for k, e in E.items()}
        for k in ("product", "oracle"):
            assert abs(accs[k] - accs["reference"]) <= max(2, int(dt_rtol * accs["reference"])), f"substep {s}: dt accumulator {k}={accs[k]} reference={accs['reference']}"
        dt = dts["reference"]
        for e in E.values():
            e.set_dt_acc(accs["reference"])
            e.stage_g2p2g(dt)
        pr = E["reference"].read_all()
        for k in ("product", "oracle"):
            if robust:  # Snow / Rankine: the reference's own f32 SVD noise dominates (tests/util.py: ROBUST_MODELS)
                errs = util.assert_particles_close_robust(E[k].read_all(), pr, scene.cell_width, f"substep {s} {k} vs reference, particles after g2p2g")
            else:
                errs = util.assert_particles_close(E[k].read_all(), pr, scene.cell_width, tol, f"substep {s} {k} vs reference, particles after g2p2g")
            for f, v in errs.items():
                worst[(k, f)] = max(worst.get((k, f), 0.0), v)
        if robust:  # the product must still match the (f64-SVD) oracle tightly on the same substep
            util.assert_particles_close(E["product"].read_all(), E["oracle"].read_all(), scene.cell_width, tol, f"substep {s} product vs oracle")
        for k in ("product", "oracle"):
            if not robust:
                util.assert_grids_close(E[k], E["reference"], 1, tol, f"substep {s} {k} vs reference, next grid after g2p2g", scene, ob)
        for e in E.values():
            e.stage_grid_update(dt, scene.gravity)
        for k in ("product", "oracle"):
            if not robust:
                util.assert_grids_close(E[k], E["reference"], 0, tol, f"substep {s} {k} vs reference, grid after grid_update", scene, ob, velocity_field=True)
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
    _lockstep(scenes.jitter(scenes.neo_hookean_block(12, 12, 8), 0.002, seed=4), ref, 8, 2e-5)


def test_kinematic_block_against_reference_kernels(ref):
    _lockstep(scenes.jitter(scenes.cube_through_sand3(12, 6, 6, 4), 0.02, seed=6), ref, 8, 2e-5)


@pytest.mark.parametrize("kind", ["nacc", "rankine", "snow", "failure", "failure_split"])
def test_other_models_against_reference_kernels(ref, kind):
    """NACC / Rankine / Snow return mappings and the maximum-stress failure criterion (with and without the split
    stress), on a pre-strained block: product and oracle in lock-step with the reference's own kernels"""
    _lockstep(scenes.jitter(scenes.model_block(kind), 0.01, seed=3), ref, 8, 5e-5, robust=kind in util.ROBUST_MODELS)


def _run_frames(engine, sc, frames):
    for _ in range(frames):
        sim, t = engine.step(sc.params, sc.gravity)
    return sim, t


def test_reference_run_to_run_spread_and_product_within_it(ref):
    """Tolerance protocol of SURVEY §8c.  The reference is non-deterministic (atomics) AND its Drucker-Prager projection
    is discontinuous at zero volumetric strain (plasticity_drucker_prager.rs:47-49 resets all singular values when
    strain_trace > 0), so round-off can flip whole particles between branches: two runs of the REFERENCE ITSELF may
    differ by far more than round-off on a few particles.  What must hold, for reference-vs-reference and for
    product-vs-reference alike: almost all particles agree tightly, and the bulk motion agrees."""
    from sparkl_b200.pipeline import context_from_scene
    sc = scenes.jitter(scenes.sand3(12, 6, 10), 0.02, seed=7)
    runs = []
    for _ in range(2):
        e = ref.RefEngine(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
        _run_frames(e, sc, 3)
        runs.append(e.read_all())
    g = context_from_scene(sc)
    sg, tg = _run_frames(g, sc, 3)
    mine = g.read_all()
    frac_rr, mean_rr = util.robust_particle_agreement(runs[0], runs[1], sc.cell_width)
    frac_pr, mean_pr = util.robust_particle_agreement(mine, runs[0], sc.cell_width)
    print(f"reference vs reference: {frac_rr:.4f} of particles differ by > 1e-3 h, mean position differs by {mean_rr:.2e} h")
    print(f"product   vs reference: {frac_pr:.4f} of particles differ by > 1e-3 h, mean position differs by {mean_pr:.2e} h")
    assert frac_pr <= max(3 * frac_rr, 0.02) and mean_pr <= max(3 * mean_rr, 1e-3)
    assert np.array_equal(mine.status["failed"], runs[0].status["failed"])


# ---------------------------------------------------------------------------------------------- non-cuboid colliders vs the reference
# The harness converts the host descriptions of HeightField / Polyline colliders into the reference's own device images (parry's
# CudaHeightFieldPtr and CudaArrayPointer1 layouts, recovered from the ld.global offsets of `grid_update` in the shipped PTX), so the
# two shapes run through the reference's kernels too (first run on a B200 in round 2: heightfield ground and polyline trough agree with
# the reference's kernels within the lock-step tolerances; see DESIGN §6 for the one deviation kept as an expected failure below).


def test_sand_on_heightfield_ground_against_reference_kernels(ref):
    """the ground of examples3d/sand3.rs through the reference's own kernels: pins the HeightField restatement"""
    _lockstep(scenes.jitter(scenes.sand3(8, 4, 6, ground="heightfield", y0=0.9), 0.02, seed=13), ref, 8, 2e-5)


def test_sand_in_a_polyline_trough_against_reference_kernels(ref):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=0.5), 0.02, seed=6)
    trough = np.array([[-1.03, 0.31], [-1.03, 6.1], [-1.52, 6.1], [-1.52, -0.53], [3.51, -0.53], [3.51, 6.1], [3.07, 6.1], [3.07, 0.31]], np.float32)
    sc.colliders = [abi.PolylineCollider(trough)]
    _lockstep(sc, ref, 8, 2e-5)


@pytest.mark.xfail(strict=False, reason="ROTATED heightfields: on the massless nodes 1.9-2.0 h away from a rotated field (12 of 1152 nodes of this scene) the "
                                       "reference's kernels keep one component of the finite-difference SDF gradient where product and oracle get none "
                                       "(collision_normal (0,-1,0) vs 0); tags, projections, velocities and particles agree (scripts/diag_colliders.py hf3)")
def test_sand_penetration3_rotated_heightfields_against_reference_kernels(ref):
    """examples3d/sand_penetration3.rs: four heightfields, three of them rotated — through the reference's own kernels"""
    _lockstep(scenes.jitter(scenes.sand_penetration3(8, 6, 6, origin=(0.0, 13.7, 3.52)), 0.02, seed=12), ref, 8, 2e-5)
