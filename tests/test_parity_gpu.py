"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs.

Bars: integer structures (active blocks, per-block particle sets, dispatch table, status flags,
projection tags) EXACT on canonical form; floating-point state within RTOL of the field scale.  RTOL is stated per
test: one substep 2e-5, growing with the number of substeps because f32 round-off and summation order (atomics in the
product and in the reference, fixed order in the oracle) diverge chaotically."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes
from tests import util

pytestmark = pytest.mark.gpu

RTOL_1 = 2e-5


def _engines(scene, oracle):
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(scene)
    o = oracle.OracleEngine(scene.cell_width)
    o.models_upload(scene.models)
    o.colliders_upload(scene.colliders)
    o.upload(scene.particles)
    return g, o


def _substep_both(g, o, scene, remaining, check=True, rtol=RTOL_1, tag=""):
    from oracle import binding as ob
    tol = util.scene_tol(scene, rtol)
    p = scene.params
    min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
    max_dt = float(min(np.float32(remaining), np.float32(p.max_substep_dt)))
    ig, io = g.stage_sort(), o.stage_sort()
    if check:
        boundary = util.boundary_particles(o.read(abi.ATTR_POSITION)["point"], scene.cell_width)
        util.assert_membership_equal(util.canonical_sort_state(g), util.canonical_sort_state(o), boundary, tag)
        if not boundary.any():
            assert ig.num_active_blocks == io.num_active_blocks, tag
    g.stage_reset_grid(); o.stage_reset_grid()
    dtg, dto = g.stage_estimate_dt(min_dt, max_dt), o.stage_estimate_dt(min_dt, max_dt)
    if check:
        assert abs(g.dt_acc() - o.dt_acc()) <= max(2, int(1e-5 * o.dt_acc())), f"{tag}: dt accumulators differ {g.dt_acc()} vs {o.dt_acc()}"
    # drive both with the ORACLE's dt so that one substep is compared on identical inputs
    g.set_dt_acc(o.dt_acc())
    dt = dto
    g.stage_g2p2g(dt); o.stage_g2p2g(dt)
    if check:
        util.assert_particles_close(g.read_all(), o.read_all(), scene.cell_width, tol, tag + " after g2p2g")
        util.assert_grids_close(g, o, 1, tol, tag + " next grid after g2p2g", scene, ob)
    g.stage_grid_update(dt, scene.gravity); o.stage_grid_update(dt, scene.gravity)
    if check:
        util.assert_grids_close(g, o, 0, tol, tag + " grid after grid_update", scene, ob, velocity_field=True)
    return dt


def _run_frames(g, o, scene, nsub, rtol_of_step):
    """lock-step substeps with per-substep comparison at a tolerance growing with the step index"""
    g.frame_begin(scene.params); o.frame_begin(scene.params)
    remaining = float(np.float32(scene.params.dt))
    for s in range(nsub):
        dt = _substep_both(g, o, scene, remaining, rtol=rtol_of_step(s), tag=f"substep {s}")
        remaining = float(np.float32(remaining) - np.float32(dt))
        if remaining <= 1e-11:
            g.frame_begin(scene.params); o.frame_begin(scene.params)
            remaining = float(np.float32(scene.params.dt))


def test_sand_first_substeps(oracle):
    sc = scenes.sand3(16, 8, 8)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 3, lambda s: RTOL_1 * (1 + 2 * s))


def test_sand_jittered_20_substeps(oracle):
    sc = scenes.jitter(scenes.sand3(12, 6, 10), 0.02, seed=7)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 20, lambda s: RTOL_1 * (1 + 3 * s))


def test_fluid_dam_break(oracle):
    sc = scenes.fluids3(16, 10, 10)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 8, lambda s: 5e-5 * (1 + 3 * s))


def test_neo_hookean_100_substeps(oracle):
    """N = 100 substeps in lock-step with the oracle (a smooth elastic scene: no discontinuous return mapping, so the
    per-substep tolerance only has to grow linearly with the substep index)"""
    sc = scenes.jitter(scenes.neo_hookean_block(10, 10, 8), 0.002, seed=9)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 100, lambda s: RTOL_1 * (1 + 3 * s))


def test_neo_hookean_block(oracle):
    sc = scenes.jitter(scenes.neo_hookean_block(12, 12, 8), 0.002, seed=4)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 10, lambda s: RTOL_1 * (1 + 3 * s))


@pytest.mark.parametrize("kind", ["nacc", "rankine", "snow", "failure", "failure_split"])
def test_other_plasticity_and_failure_models(oracle, kind):
    sc = scenes.jitter(scenes.model_block(kind), 0.01, seed=3)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 8, lambda s: 5e-5 * (1 + 3 * s))


def test_sand_on_heightfield_ground(oracle):
    """the reference scenes' own ground (examples3d/sand3.rs:23-38): HeightField collider, §8f-1"""
    sc = scenes.jitter(scenes.sand3(10, 5, 8, ground="heightfield", y0=0.45), 0.02, seed=5)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 10, lambda s: RTOL_1 * (1 + 3 * s))
    tags = np.bincount(g.grid_nodes(0)["projection_status_tag"].ravel().astype(int), minlength=4)
    assert tags[1] > 0 and tags[2] > 0 and tags[3] > 0  # inside / outside / too far all occur


def test_sand_in_a_polyline_trough(oracle):
    """Polyline collider (gpu_collider.rs:105-201), §8f-1"""
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=0.5), 0.02, seed=6)
    trough = np.array([[-1, 0.3], [-1, 6], [-1.5, 6], [-1.5, -0.5], [3.5, -0.5], [3.5, 6], [3, 6], [3, 0.3]], np.float32)  # counter-clockwise: the U-shaped wall is the solid
    sc.colliders = [abi.PolylineCollider(trough)]
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 10, lambda s: RTOL_1 * (1 + 3 * s))


def test_sand_on_a_trimesh_ramp(oracle):
    """TriMesh collider (gpu_collider.rs:58-68): a bumpy triangulated ramp over a cuboid floor"""
    from tests.test_device_grid3d_cpu import ramp_mesh
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=2.4), 0.02, seed=15)
    verts, tris = ramp_mesh()
    sc.colliders = [abi.TriMeshCollider(verts, tris, friction=0.4), abi.CuboidCollider((9.0, 0.5, 9.0), (0.0, 1.0, 0.0))]
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 10, lambda s: RTOL_1 * (1 + 3 * s))


def test_invalid_collider_shapes_are_rejected():
    from sparkl_b200.pipeline import SingleGpuMpmContext
    g = SingleGpuMpmContext(0.2)
    img = np.zeros((1, 336), np.uint8)
    img[0, 0] = 7  # no such GpuColliderShape variant
    assert g._lib.spk_colliders_upload(g._h, img.ctypes.data, 1) == abi.EUNSUPPORTED
    img[0, 0] = 2  # TriMesh without vertices / indices
    assert g._lib.spk_colliders_upload(g._h, img.ctypes.data, 1) == abi.EINVAL
    from tests.test_device_grid3d_cpu import ramp_mesh
    verts, tris = ramp_mesh()
    tris[3, 1] = len(verts)  # vertex index out of range
    with pytest.raises(abi.SparklError):
        g.colliders_upload([abi.TriMeshCollider(verts, tris)])


def test_cube_through_sand_kinematic(oracle):
    sc = scenes.cube_through_sand3(12, 6, 6, 4)
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 10, lambda s: RTOL_1 * (1 + 3 * s))


def test_full_step_matches_oracle(oracle):
    """spk_step (whole frame, adaptive substeps) against orc_step: same substep count, state within tolerance."""
    sc = scenes.jitter(scenes.sand3(10, 6, 6), 0.01, seed=3)
    g, o = _engines(sc, oracle)
    for frame in range(2):
        sg, tg = g.step(sc.params, sc.gravity)
        so, to = o.step(sc.params, sc.gravity)
        assert tg.num_substeps == to.num_substeps
        assert abs(sg - so) <= 1e-6
        assert tg.g2p2g_launches == tg.num_substeps and tg.kernel_launches >= 8 * tg.num_substeps
        util.assert_particles_close(g.read_all(), o.read_all(), sc.cell_width, util.scene_tol(sc, 1e-4 * (frame + 1)), f"frame {frame}")


def test_empty_and_single_particle(oracle):
    sc = scenes.sand3(1, 1, 1)
    from sparkl_b200.pipeline import SingleGpuMpmContext
    g = SingleGpuMpmContext(sc.cell_width)
    g.models.from_models(sc.models)
    g.upload(abi.ParticleComponents(0))
    sim, t = g.step(sc.params, sc.gravity)
    assert sim == pytest.approx(sc.params.dt) and t.num_substeps == 0
    g2, o = _engines(sc, oracle)
    _run_frames(g2, o, sc, 4, lambda s: RTOL_1 * (1 + 2 * s))


def test_ragged_blocks_and_chunk_boundaries(oracle):
    """blocks holding 1 .. >512 particles (several g2p2g dispatch entries per block, partial last chunk)"""
    rng = np.random.Generator(np.random.PCG64(11))
    h = 0.2
    pts = [rng.uniform(2.5, 2.5 + 4 * h, (1500, 3)), rng.uniform(6.0, 6.0 + 12 * h, (700, 3)), np.array([[9.05, 9.05, 9.05]])]
    pts = np.concatenate(pts).astype(np.float32)
    sc = scenes.sand3(1, 1, 1)
    sc.particles = abi.particles_from_positions(pts, 0, 0.05, 2700.0)
    sc.particles.status["model_index"][::2] = 1
    g, o = _engines(sc, oracle)
    _run_frames(g, o, sc, 3, lambda s: 5e-5 * (1 + 3 * s))


def test_async_readback_matches_blocking_reads():
    """spk_particles_read_async / _wait (the testbed's per-frame read-back, §8f-2): same bytes as the blocking reads, several
    attributes in flight, frames enqueued back to back"""
    from sparkl_b200.pipeline import context_from_scene
    sc = scenes.jitter(scenes.sand3(12, 6, 8), 0.02, seed=8)
    g = context_from_scene(sc)
    n = len(sc.particles)
    for frame in range(3):
        g.step(sc.params, sc.gravity)
        outs = {w: np.empty(n, abi.ATTR_DTYPES[w]) for w in (abi.ATTR_POSITION, abi.ATTR_VELOCITY, abi.ATTR_PHASE, abi.ATTR_VOLUME)}
        for w, o in outs.items():
            g.read_async(w, o)
        if frame < 2:
            g.read_wait()
            for w, o in outs.items():
                assert np.array_equal(o, g.read(w)), (frame, w)
        else:  # enqueue the next frame before waiting: the copies must hold the state of THIS frame
            want = {w: g.read(w) for w in outs}
            g.read_async(abi.ATTR_POSITION, outs[abi.ATTR_POSITION])  # a second read of the same attribute queues behind the first
            g.step(sc.params, sc.gravity)
            g.read_wait()
            for w, o in outs.items():
                assert np.array_equal(o, want[w]), (frame, w)


def test_append_remove_read_roundtrip(oracle):
    sc = scenes.jitter(scenes.sand3(6, 4, 4), 0.01, seed=5)
    from sparkl_b200.pipeline import context_from_scene
    g = context_from_scene(sc)
    n = len(sc.particles)
    back = g.read_all()
    for name in ("position", "velocity", "volume", "status", "phase", "data"):
        assert np.array_equal(getattr(back, name), getattr(sc.particles, name)), name
    g.step(sc.params, sc.gravity)
    extra = sc.particles.take(np.arange(10))
    g.append(extra)
    assert len(g) == n + 10
    st = g.read(abi.ATTR_POSITION)
    assert np.array_equal(st[n:], extra.position)
    g.remove_range(3, 8)
    assert len(g) == n + 5
    assert np.array_equal(g.read(abi.ATTR_POSITION)[n - 5:], extra.position)
    g.step(sc.params, sc.gravity)
    assert np.isfinite(g.read(abi.ATTR_POSITION)["point"]).all()


def test_async_upload_matches_blocking_upload():
    """spk_particles_upload_async: same state as the blocking upload, also when it is enqueued behind a running frame"""
    import torch
    from sparkl_b200.pipeline import context_from_scene
    sc = scenes.jitter(scenes.sand3(10, 6, 6), 0.02, seed=9)
    a, b = context_from_scene(sc), context_from_scene(sc)
    pins = {}
    pc = abi.ParticleComponents(0)
    for name in ("status", "position", "velocity", "volume", "phase", "data"):
        src = getattr(sc.particles, name)
        t = torch.empty(src.nbytes, dtype=torch.uint8).pin_memory()
        arr = t.numpy().view(src.dtype).reshape(src.shape)
        arr[...] = src
        pins[name] = t
        setattr(pc, name, arr)
    for frame in range(3):
        a.upload(sc.particles)
        b.upload_async(pc)                       # queued behind the previous frame of `b`
        a.step(sc.params, sc.gravity)
        b.step(sc.params, sc.gravity, timings=False)
    pa, pb = a.read_all(), b.read_all()
    assert np.array_equal(pa.status, pb.status)  # discrete state: exact; the floating-point fields differ by the order of the P2G atomics
    util.assert_particles_close(pb, pa, sc.cell_width, util.scene_tol(sc, 1e-4), "async vs blocking upload")
    # an out-of-range model index is reported by the next step
    pc.status["model_index"][0] = 1 << 25
    b.upload_async(pc)
    with pytest.raises(abi.SparklError):
        b.step(sc.params, sc.gravity)
