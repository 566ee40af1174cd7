"""GPU parity of the dim2 build (libsparkl_b200_2d.so): the CUDA path through the C ABI against the dim2 CPU
oracle (oracle/mpm_oracle2d.cpp) on the same seeded inputs, stage by stage, exactly like tests/test_parity_gpu.py does for 3D.
Integer structures (active blocks, per-block particle sets, dispatch table, status flags, projection tags) EXACT on canonical
form; floating-point state within RTOL of the field scale (2e-5 for one substep, growing with the substep index).
BASELINE.json configs[0] (`elasticity2`) is the first scene."""
import numpy as np
import pytest

from sparkl_b200 import abi, abi2d, scenes2d
from tests import util

pytestmark = pytest.mark.gpu

RTOL_1 = 2e-5


def _engines(scene):
    from oracle import binding as ob
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(scene)
    assert g.dim == 2 and g._lib.spk_dim() == 2
    o = ob.OracleEngine2D(scene.cell_width)
    o.models_upload(scene.models)
    o.colliders_upload(scene.colliders)
    o.upload(scene.particles)
    return g, o


def _is_fluid_scene(scene):
    return any(m.constitutive_model.tag == 2 for m in scene.models)


def _substep_both(g, o, scene, remaining, rtol, tag, robust=False):
    from oracle import binding as ob
    tol = util.scene_tol(scene, rtol)
    p = scene.params
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    max_dt = float(min(np.float32(remaining), np.float32(p.max_substep_dt)))
    ig, io = g.stage_sort(), o.stage_sort()
    boundary = util.boundary_particles(o.read(abi.ATTR_POSITION)["point"], scene.cell_width)
    util.assert_membership_equal(util.canonical_sort_state(g), util.canonical_sort_state(o), boundary, tag)
    if not boundary.any():
        assert ig.num_active_blocks == io.num_active_blocks, tag
    g.stage_reset_grid(); o.stage_reset_grid()
    g.stage_estimate_dt(min_dt, max_dt)
    dto = o.stage_estimate_dt(min_dt, max_dt)
    # Monaghan's single-particle bound (eos_monaghan_sph.rs:86-90) is (h / J) sqrt(rho0 (J - 1) / (k p(J) d)): at J = 1 -+ 1 ulp it
    # is a ratio of two numbers that are each a few ulps of 1 away from zero, so a 1-ulp difference in J (summation order of the
    # G2P sums) moves it by tens of percent, and at J == 1 exactly it is NaN = "no bound".  For fluid scenes the accumulators are
    # therefore not compared; both sides are driven with the oracle's dt below, so everything else still is.
    dt_rtol = 1e9 if _is_fluid_scene(scene) else 1e-5
    assert abs(g.dt_acc() - o.dt_acc()) <= max(2, int(dt_rtol * o.dt_acc())), f"{tag}: dt accumulators differ {g.dt_acc()} vs {o.dt_acc()}"
    g.set_dt_acc(o.dt_acc())  # drive both with the ORACLE's dt so that one substep is compared on identical inputs
    dt = dto
    g.stage_g2p2g(dt); o.stage_g2p2g(dt)
    if robust:
        util.assert_particles_close_robust(g.read_all(), o.read_all(), scene.cell_width, tag + " after g2p2g")
    else:
        util.assert_particles_close(g.read_all(), o.read_all(), scene.cell_width, tol, tag + " after g2p2g")
        util.assert_grids_close(g, o, 1, tol, tag + " next grid after g2p2g", scene, ob)
    g.stage_grid_update(dt, scene.gravity); o.stage_grid_update(dt, scene.gravity)
    if not robust:
        util.assert_grids_close(g, o, 0, tol, tag + " grid after grid_update", scene, ob, velocity_field=True)
    return dt


def _run(scene, nsub, growth=3.0, robust=False):
    g, o = _engines(scene)
    g.frame_begin(scene.params); o.frame_begin(scene.params)
    remaining = float(np.float32(scene.params.dt))
    for s in range(nsub):
        dt = _substep_both(g, o, scene, remaining, RTOL_1 * (1 + growth * s), f"{scene.name} substep {s}", robust)
        remaining = float(np.float32(remaining) - np.float32(dt))
        if remaining <= 1e-11:
            g.frame_begin(scene.params); o.frame_begin(scene.params)
            remaining = float(np.float32(scene.params.dt))
    return g, o


def test_upload_read_roundtrip_dim2():
    sc = scenes2d.jitter(scenes2d.model_block2("rankine", 12, 9), 0.01, seed=3)
    sc.particles.status["kinematic_vel"][::3] = (0.5, -0.25)
    sc.particles.velocity["vector"][:, 0] = np.linspace(-1, 1, len(sc.particles), dtype=np.float32)
    g, _ = _engines(sc)
    back = g.read_all()
    for name in ("position", "velocity", "volume", "status", "phase", "data"):
        assert np.array_equal(getattr(back, name), getattr(sc.particles, name)), name
    assert np.array_equal(g.read(abi.ATTR_SORTED_IDS), np.arange(len(sc.particles), dtype=np.uint32))


def test_elasticity2_lockstep():
    """BASELINE.json configs[0]: star-shaped corotated + Rankine blocks over the cuboid ground (examples2d/elasticity2.rs)"""
    sc = scenes2d.elasticity2(nblocks=3, scale=0.35)
    _run(sc, 12)


def _elasticity2_on_the_ground_scene():
    sc = scenes2d.elasticity2(nblocks=2, scale=0.3)
    pts = sc.particles.position["point"]
    pts[:, 1] -= pts[:, 1].min() - np.float32(sc.cell_width * 40.0 + 0.01)  # lowest particle 0.01 above the ground surface
    sc.particles.velocity["vector"][:, 1] = -1.0
    return sc


def test_elasticity2_on_the_ground():
    """the same blocks resting on / pressed into the ground so that the collider branch of grid_update is exercised"""
    _run(_elasticity2_on_the_ground_scene(), 10)


@pytest.mark.parametrize("kind", ["corotated", "neo_hookean", "fluid", "sand", "nacc", "failure", "failure_split"])
def test_model_block2_lockstep(kind):
    sc = scenes2d.jitter(scenes2d.model_block2(kind, 14, 12), 0.01, seed=11)
    _run(sc, 6)


@pytest.mark.parametrize("kind", ["rankine", "snow"])
def test_model_block2_lockstep_clamping_models(kind):
    """Snow / Rankine clamp singular values: where two of them coalesce U, V are ill-defined and f32 vs f64 SVDs recompose F with
    a visible difference on a few particles (see tests/util.py: ROBUST_MODELS); compared statistically, as in 3D"""
    sc = scenes2d.jitter(scenes2d.model_block2(kind, 14, 12), 0.01, seed=11)
    _run(sc, 6, robust=True)


def test_fluids2_dam_break_lockstep():
    sc = scenes2d.fluids2(40, 30)
    _run(sc, 8)


def _polyline_trough_scene():
    sc = scenes2d.model_block2("sand", 12, 10)
    ext = 12 * 0.1
    # counter-clockwise = solid; no vertex / edge passes through a grid node (multiples of h = 0.2): a node exactly on an edge has
    # side == 0 in exact arithmetic and its inside / outside flag is decided by round-off
    trough = abi2d.PolylineCollider([(-1.03, 3.51), (-1.03, 2.23), (ext / 2 + 0.013, 1.77), (ext + 1.07, 2.23), (ext + 1.07, 3.51), (ext + 2.09, 3.51),
                                     (ext + 2.09, 0.47), (-2.11, 0.47), (-2.11, 3.51)][::-1])
    sc.colliders = [trough]
    sc.particles.velocity["vector"][:, 1] = -2.0
    return sc


def test_polyline_trough_dim2():
    """sand dropped into a polyline trough: the in-tree polyline_project_point (gpu_collider.rs:105-201, dim2 normal)"""
    _run(_polyline_trough_scene(), 8)


def _basic2_scene():
    sc = scenes2d.basic2(drop=0.004)
    sc.particles.position["point"][:, 1] -= np.float32(0.045)  # start ~1 cell above the surface
    return sc


def test_basic2_on_the_heightfield_ground():
    """examples2d/basic2.rs: snow and sand columns dropped on the reference's heightfield ground (parry2d HeightField restated:
    product vs oracle; the columns reach the ground within the compared substeps)"""
    g, o = _run(_basic2_scene(), 14)
    tags = np.bincount(o.grid_nodes(0)["projection_status_tag"].ravel().astype(int), minlength=4)
    assert tags[1] + tags[2] > 0 and tags[3] > 0  # nodes near the surface got a projection, far ones TooFar (max_dist = 2 h)


def test_full_step_matches_oracle_dim2():
    """whole frames through spk_step (adaptive substeps, dt fused into g2p2g) against orc_step: same substep count, same
    simulated time, particle state within the multi-substep tolerance"""
    sc = scenes2d.elasticity2(nblocks=2, scale=0.3)
    g, o = _engines(sc)
    for _ in range(2):
        sim_g, t = g.step(sc.params, sc.gravity)
        sim_o, to = o.step(sc.params, sc.gravity)
        assert t.num_substeps == to.num_substeps
        assert abs(sim_g - sim_o) <= 1e-6
    util.assert_particles_close(g.read_all(), o.read_all(), sc.cell_width, util.scene_tol(sc, RTOL_1 * (1 + 3.0 * 2 * t.num_substeps)), "two frames")


def test_momentum_conservation_and_free_fall_dim2():
    """property checks that do not involve the oracle: without colliders P2G conserves mass and momentum, and the centre of
    mass follows the free-fall parabola"""
    sc = scenes2d.elasticity2(nblocks=3, scale=0.35, with_colliders=False)
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(sc)
    m = sc.particles.volume["mass"].astype(np.float64)
    y0 = float((sc.particles.position["point"][:, 1] * m).sum() / m.sum())
    T = 0.0
    for _ in range(6):
        sim, t = g.step(sc.params, sc.gravity)
        T += sim
        nodes = g.grid_nodes(0)
        assert abs(float(nodes["mass"].astype(np.float64).sum()) - m.sum()) <= 1e-5 * m.sum()
    v = g.read(abi.ATTR_VELOCITY)["vector"].astype(np.float64)
    p = g.read(abi.ATTR_POSITION)["point"].astype(np.float64)
    vy = float((v[:, 1] * m).sum() / m.sum())
    y = float((p[:, 1] * m).sum() / m.sum())
    # the first substep of a run advects with the velocities of an empty grid: v lags by one substep (~0.35 %)
    assert abs(vy - (-9.81 * T)) <= 1e-2 * 9.81 * T
    assert abs((y - y0) - (-0.5 * 9.81 * T * T)) <= 0.02 * 0.5 * 9.81 * T * T
    assert abs(float((v[:, 0] * m).sum())) <= 1e-5 * m.sum()


def test_unsupported_dim2_features_fail_loudly():
    import ctypes as C
    from sparkl_b200.pipeline import SingleGpuMpmContext
    g = SingleGpuMpmContext(0.1, dim=2)
    buf = np.zeros((1, 72), np.uint8)
    buf.view("<u4")[0, 0] = 7  # not a GpuColliderShape
    assert g._lib.spk_colliders_upload(g._h, buf.ctypes.data_as(C.c_void_p), 1) == abi.EUNSUPPORTED
    for tag in (1, 2):  # heightfield without samples, trimesh without vertices
        buf.view("<u4")[0, 0] = tag
        assert g._lib.spk_colliders_upload(g._h, buf.ctypes.data_as(C.c_void_p), 1) == abi.EINVAL
    desc = np.zeros(128, np.uint8)
    assert g._lib.spk_dist_export(g._h, desc.ctypes.data_as(C.c_void_p)) == abi.EUNSUPPORTED


# ---------------------------------------------------------------------------------------------- against the reference's own dim2 kernels
@pytest.fixture(scope="module")
def ref2d():
    from oracle import binding as ob
    if not ob.ref2d_available():
        pytest.skip("oracle/_ref (reference dim2 kernels) not built on this box")
    return ob


@pytest.mark.parametrize("case", ["elasticity2", "fluids2", "corotated2", "neo_hookean2", "sand2", "nacc2", "failure2", "failure_split2"])
def test_three_way_lockstep_with_reference_dim2_kernels(ref2d, case):
    """product vs dim2 oracle vs resources/sparkl2d-kernels.ptx (assembled for sm_100a, driven by the -DSPK_DIM=2 harness) on the
    collider-free fixture scenes: canonical sort state exact, dt accumulator, particles and grids within tolerance, per substep"""
    from scripts.make_golden import build_case
    from tests.test_reference_gpu import _lockstep
    sc, nsub, _ = build_case(case)
    worst = _lockstep(sc, ref2d, nsub, 5e-5 if case == "fluids2" else 2e-5, dt_rtol=1e9 if case == "fluids2" else 1e-5)
    print(case, {f"{k[0]}.{k[1]}": f"{v:.1e}" for k, v in worst.items()})


@pytest.mark.parametrize("case", ["rankine2", "snow2"])
def test_three_way_lockstep_with_reference_dim2_kernels_clamping_models(ref2d, case):
    from scripts.make_golden import build_case
    from tests.test_reference_gpu import _lockstep
    sc, nsub, _ = build_case(case)
    _lockstep(sc, ref2d, nsub, 2e-5, robust=True)



# The dim2 GpuCollider byte image of the reference (280 bytes; parry2d's device heightfield / polyline layouts) was recovered from the
# ld.global offsets of `grid_update` in resources/sparkl2d-kernels.ptx (oracle/ref_harness.cpp: ref_colliders_upload), so the collider
# scenes go through the reference's own dim2 kernels too (first run on a B200 in round 2: polyline trough and heightfield ground agree).

@pytest.mark.parametrize("which", [pytest.param("cuboid_ground", marks=pytest.mark.xfail(strict=False, reason=(
    "grid nodes EXACTLY one cell above the axis-aligned ground: the reference's kernels get a projection distance of h + 3e-8 and skip "
    "the friction branch (grid_update.rs:171-186), product and oracle get <= h and apply it (velocity 0 vs -8.4e-5 on 10 light nodes); "
    "scripts/diag_colliders.py c2"))), "polyline_trough", "heightfield_ground"])
def test_three_way_lockstep_with_colliders_dim2(ref2d, which):
    from tests.test_reference_gpu import _lockstep
    sc, nsub = {"cuboid_ground": (_elasticity2_on_the_ground_scene, 10), "polyline_trough": (_polyline_trough_scene, 8),
                "heightfield_ground": (_basic2_scene, 14)}[which]
    _lockstep(sc(), ref2d, nsub, 2e-5)
