"""CPU tests of the oracle (the CPU restatement of the reference's GPU algorithm).

The reference holds no golden vectors for this path, so the oracle is pinned here on (a) closed-form
known answers of every building block, (b) the invariants of the reference's own (commented-out) sort checker
`check_active_blocks` (src/cuda/cuda_sparse_grid.rs:491-652), (c) conservation laws, and on the GPU box against the
reference's shipped kernels (tests/test_reference_gpu.py)."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes
from tests import util

ORIGIN = 1 << 20
MASK = (1 << 21) - 1


def unpack(k):
    return np.array([k & MASK, (k >> 21) & MASK, (k >> 42) & MASK], np.int64) - ORIGIN


def test_block_key_of_point(oracle):
    """gpu_grid.rs:271-279: cell = round(p/h) (half away from zero), block = floor((cell - 2) / 4)"""
    h = 0.5
    cases = {(1.0, 1.0, 1.0): (0, 0, 0),        # cell 2 -> block 0
             (0.74, 0.76, 0.0): (-1, 0, -1),     # cells 1, 2, 0
             (2.75, -0.25, -0.26): (0, -1, -1),  # 5.5 -> 6 -> block 1? (6-2)/4 = 1
             (-1.25, 3.0, 2.99): (-2, 1, 1)}     # -2.5 -> -3 -> floor(-5/4) = -2 ; 6 -> 1 ; 5.98 -> 6 -> 1
    cases[(2.75, -0.25, -0.26)] = (1, -1, -1)
    for p, want in cases.items():
        got = unpack(oracle.block_key_of_point(p, h))
        assert tuple(got) == want, (p, got, want)
    # NaN positions saturate to cell 0 (Rust `as i64`)
    assert tuple(unpack(oracle.block_key_of_point((float("nan"), 0.0, 0.0), h))) == (-1, -1, -1)


def test_quadratic_weights_partition_of_unity(oracle):
    """kernel.rs:15-21 for x in [0.5, 1.5]: weights sum to 1, first moment reproduces x"""
    import ctypes as C
    l = oracle.lib()
    for x in np.linspace(0.5, 1.5, 41):
        w = np.zeros(3, np.float32)
        l.orc_quadratic_weights(C.c_float(x), w.ctypes.data_as(C.c_void_p))
        assert abs(w.sum() - 1.0) < 3e-7
        assert abs((w * np.arange(3)).sum() - x) < 1e-6
        assert (w >= 0).all()


def test_dt_fixed_point(oracle):
    """gpu_timestep.rs:17-26: u64(dt * 1e12) in f32, saturating, NaN -> 0"""
    l = oracle.lib()
    assert l.orc_dt_from_sec(1.0 / 60.0) == int(np.float32(np.float32(1.0 / 60.0) * np.float32(1e12)))
    assert l.orc_dt_from_sec(float("nan")) == 0 and l.orc_dt_from_sec(-1.0) == 0
    assert l.orc_dt_from_sec(1e30) == 2**64 - 1
    assert l.orc_dt_into_sec(16666667008) == pytest.approx(0.016666667, rel=1e-6)


def test_svd_stand_in(oracle):
    rng = np.random.Generator(np.random.PCG64(0))
    for _ in range(200):
        F = (np.eye(3) + 0.3 * rng.standard_normal((3, 3))).astype(np.float32)
        u, s, vt = oracle.svd3(F)
        assert np.allclose(u @ np.diag(s) @ vt, F, atol=2e-6)
        assert np.allclose(u @ u.T, np.eye(3), atol=2e-6) and np.allclose(vt @ vt.T, np.eye(3), atol=2e-6)
        assert (s >= 0).all()
        assert np.allclose(np.sort(s), np.sort(np.linalg.svd(F.astype(np.float64), compute_uv=False)), atol=2e-6)
    # rank deficient input still returns orthogonal factors
    u, s, vt = oracle.svd3(np.diag([2.0, 0.0, 0.0]).astype(np.float32))
    assert np.allclose(u @ u.T, np.eye(3), atol=1e-6) and np.allclose(np.sort(s), [0, 0, 2])


def test_corotated_stress_known_answer(oracle):
    """elasticity_corotated_linear.rs:31-45 at phase 1: tau = 2 mu (F - R) F^T + lambda (J - 1) J I"""
    m = abi.ParticleModel(abi.CorotatedLinearElasticity(1.0e5, 0.3))
    lam, mu = float(m.constitutive_model.lam), float(m.constitutive_model.mu)
    # pure rotation: zero stress
    c, s = np.cos(0.7), np.sin(0.7)
    R = np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])
    assert np.abs(oracle.kirchhoff_stress(m, R)).max() < 0.05  # f32 noise * 2mu
    # uniform stretch a: tau = (2 mu (a - 1) a + lambda (a^3 - 1) a^3) I
    a = 1.1
    tau = oracle.kirchhoff_stress(m, np.eye(3) * a)
    want = 2 * mu * (a - 1) * a + lam * (a ** 3 - 1) * a ** 3
    assert np.allclose(tau, np.eye(3) * want, rtol=1e-5, atol=1e-2)
    # rotation * stretch: tau = R diag(...) R^T
    S = np.diag([1.2, 0.9, 1.05])
    tau = oracle.kirchhoff_stress(m, R @ S)
    J = np.linalg.det(S)
    want = R @ (2 * mu * (S - np.eye(3)) @ S) @ R.T + lam * (J - 1) * J * np.eye(3)
    assert np.allclose(tau, want, rtol=2e-5, atol=0.05)


def test_neo_hookean_and_monaghan_stress(oracle):
    m = abi.ParticleModel(abi.NeoHookeanElasticity(1.0e5, 0.3))
    lam, mu = float(m.constitutive_model.lam), float(m.constitutive_model.mu)
    F = np.diag([1.2, 0.9, 1.05])
    J = np.linalg.det(F)
    B = F @ F.T
    dev = mu * J ** (-2 / 3) * (B - np.trace(B) / 3 * np.eye(3))
    vol = (2 / 3 * mu + lam) / 2 * (J * J - 1) * np.eye(3)
    assert np.allclose(oracle.kirchhoff_stress(m, F), dev + vol, rtol=2e-5, atol=0.05)  # J >= 1, phase 1
    # Monaghan: tau = -max(p0 ((rho/rho0)^gamma - 1), -max_neg) * Jp * I, rho/rho0 = 1/F00 (quirk 7)
    w = abi.ParticleModel(abi.MonaghanSphEos(1.0e4, 7, 0.0))
    F = np.eye(3); F[0, 0] = 0.98
    tau = oracle.kirchhoff_stress(w, F, jp=0.5)
    assert np.allclose(tau, -(1.0e4 * ((1 / 0.98) ** 7 - 1)) * 0.5 * np.eye(3), rtol=1e-4)
    F[0, 0] = 1.5  # expansion: pressure clamped at -max_neg_pressure = -1
    assert np.allclose(oracle.kirchhoff_stress(w, F, jp=1.0), 1.0 * np.eye(3), rtol=1e-5)


def _data():
    d = np.zeros(1, abi.DATA)
    d["nacc_alpha"], d["plastic_hardening"], d["elastic_hardening"] = -0.01, 1.0, 1.0
    return d


def test_drucker_prager_return_mapping(oracle):
    """plasticity_drucker_prager.rs:36-104"""
    m = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(1e7, 0.2), abi.DruckerPragerPlasticity(1e7, 0.2))
    # expansion (strain trace > 0): singular values reset to 1, F becomes the rotation, hardening += |strain|
    F = np.diag([1.1, 1.05, 1.02])
    F2, jp, d = oracle.plastic_update(m, F, _data())
    assert np.allclose(F2, np.eye(3), atol=1e-6)
    assert float(d["plastic_hardening"][0]) == pytest.approx(1.0 + np.linalg.norm(np.log([1.1, 1.05, 1.02])), rel=1e-5)
    assert jp == pytest.approx(1.1 * 1.05 * 1.02, rel=1e-5)
    assert float(d["log_vol_gain"][0]) == pytest.approx(np.log(1.1 * 1.05 * 1.02), rel=1e-4)
    # pure isotropic compression: deviatoric strain is exactly zero -> also the "reset" branch (quirk 9)
    F2, jp, d = oracle.plastic_update(m, np.eye(3) * 0.5, _data())
    assert np.allclose(F2, np.eye(3), atol=1e-6)
    # compression with small shear stays inside the cone: untouched
    F = np.diag([0.9, 0.895, 0.9])
    F2, jp, d = oracle.plastic_update(m, F, _data())
    assert np.allclose(F2, F, atol=1e-7) and jp == 1.0 and float(d["plastic_hardening"][0]) == 1.0
    # large shear under mild compression: projected onto the cone, volume preserved in log space along the deviator
    F = np.diag([1.3, 0.7, 0.99])
    F2, jp, d = oracle.plastic_update(m, F, _data())
    e, e2 = np.log(np.diag(F)), np.log(np.diag(F2))
    assert np.allclose(e.sum(), e2.sum(), atol=1e-5)          # trace of the Hencky strain unchanged
    dev, dev2 = e - e.mean(), e2 - e2.mean()
    assert np.linalg.norm(dev2) < np.linalg.norm(dev) and np.allclose(dev2 / np.linalg.norm(dev2), dev / np.linalg.norm(dev), atol=1e-4)


def test_snow_rankine_nacc_smoke(oracle):
    F = np.diag([1.2, 0.8, 1.0])
    m = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(1e5, 0.3), abi.SnowPlasticity())
    F2, jp, d = oracle.plastic_update(m, F, _data())
    assert np.allclose(np.diag(F2), [1.0045, 0.975, 1.0], atol=1e-6)
    assert jp == pytest.approx(1.2 * 0.8 / (1.0045 * 0.975), rel=1e-5)
    assert float(d["elastic_hardening"][0]) == pytest.approx(np.exp(10 * (1 - jp)), rel=1e-4)
    m = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(1e5, 0.3), abi.RankinePlasticity(1e5, 0.3, 10.0, 5.0))
    F2, _, d = oracle.plastic_update(m, np.diag([1.00001, 1.0, 1.0]), _data())
    assert np.allclose(F2, np.diag([1.00001, 1.0, 1.0]), atol=1e-7)  # below the tensile strength: elastic
    F2, _, d = oracle.plastic_update(m, np.diag([1.2, 1.0, 1.0]), _data())
    assert F2[0, 0] < 1.2 and float(d["plastic_hardening"][0]) > 1.0
    m = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(1e5, 0.3), abi.NaccPlasticity(1e5, 0.3, 0.5, True, 3.0, 0.6))
    F2, _, d = oracle.plastic_update(m, np.diag([1.3, 1.3, 1.3]), _data())  # p_tr < -beta p0: projected to the tensile tip
    assert np.allclose(F2, np.eye(3) * F2[0, 0], atol=1e-6) and F2[0, 0] < 1.3


def test_cuboid_projection(oracle):
    """parry Aabb::project_local_point, solid = false: outside -> clamp, inside -> nearest face, first axis wins ties"""
    c = abi.CuboidCollider((1.0, 2.0, 3.0), translation=(10.0, 0.0, 0.0))
    p, inside = oracle.cuboid_project(c, (12.0, 0.5, 0.0))
    assert not inside and np.allclose(p, (11.0, 0.5, 0.0))
    p, inside = oracle.cuboid_project(c, (10.9, 0.0, 0.0))
    assert inside and np.allclose(p, (11.0, 0.0, 0.0))
    p, inside = oracle.cuboid_project(c, (10.0, 1.9, 2.5))
    assert inside and np.allclose(p, (10.0, 2.0, 2.5))
    p, inside = oracle.cuboid_project(c, (13.0, 3.0, 4.0))
    assert not inside and np.allclose(p, (11.0, 2.0, 3.0))
    # rotated by 90 degrees about z: local x axis becomes world y
    q = (0.0, 0.0, np.sin(np.pi / 4), np.cos(np.pi / 4))
    c = abi.CuboidCollider((1.0, 2.0, 3.0), rotation=q)
    p, inside = oracle.cuboid_project(c, (0.0, 1.5, 0.0))
    assert not inside and np.allclose(p, (0.0, 1.0, 0.0), atol=1e-6)


def _engine(oracle, sc):
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models)
    e.colliders_upload(sc.colliders)
    e.upload(sc.particles)
    return e


def check_active_blocks(engine, pos, h, oracle):
    """the reference's own sort invariants, cuda_sparse_grid.rs:491-652"""
    vids, per_block, multiset, hdr, ids = util.canonical_sort_state(engine)
    n = len(pos)
    assert len(np.unique(ids)) == n and ids.max(initial=0) < max(n, 1)            # unique in-range sorted ids
    assert len(np.unique(hdr["virtual_id"])) == len(hdr)                          # map <-> active blocks bijection
    assert int(hdr["num_particles"].sum()) == n                                   # every particle visited exactly once
    active = set(int(v) for v in hdr["virtual_id"])
    for h_ in hdr:
        s = ids[h_["first_particle"]: h_["first_particle"] + h_["num_particles"]]
        for pid in s[:: max(1, len(s) // 7)]:                                     # particle's block == header's virtual id
            assert oracle.block_key_of_point(pos[pid], h) == int(h_["virtual_id"])
        if h_["num_particles"]:
            b = unpack(int(h_["virtual_id"])) + ORIGIN                            # its 8 associated blocks are active
            for o in range(8):
                k = (int(b[0]) + (o & 1)) | ((int(b[1]) + ((o >> 1) & 1)) << 21) | ((int(b[2]) + ((o >> 2) & 1)) << 42)
                assert k in active
    # dispatch blocks tile every block's particle range in steps of 128
    want = sorted((int(v["virtual_id"]), 128 * k) for v in hdr for k in range((int(v["num_particles"]) + 127) // 128))
    assert multiset == want
    # first_particle is the exclusive scan of num_particles in id order
    assert np.array_equal(hdr["first_particle"], np.concatenate([[0], np.cumsum(hdr["num_particles"])[:-1]]).astype(np.uint32))


def test_sort_invariants(oracle):
    rng = np.random.Generator(np.random.PCG64(42))
    pts = np.concatenate([rng.uniform(-3, 3, (3000, 3)), rng.uniform(5.0, 5.8, (900, 3))]).astype(np.float32)
    sc = scenes.sand3(1, 1, 1)
    sc.particles = abi.particles_from_positions(pts, 0, 0.05, 2700.0)
    e = _engine(oracle, sc)
    info = e.stage_sort()
    assert info.num_active_blocks == e.num_active_blocks(1)
    check_active_blocks(e, pts, sc.cell_width, oracle)
    assert info.num_dispatch_blocks == sum((int(c) + 127) // 128 for c in e.grid_headers(1)["num_particles"])


def test_empty_set(oracle):
    sc = scenes.sand3(1, 1, 1)
    sc.particles = abi.ParticleComponents(0)
    e = _engine(oracle, sc)
    sim, t = e.step(sc.params, sc.gravity)
    assert sim == pytest.approx(sc.params.dt) and t.num_substeps == 0


def test_mass_momentum_conservation_and_first_substep(oracle):
    """P2G conserves mass exactly up to f32 rounding; the first substep discards initial velocities"""
    sc = scenes.jitter(scenes.sand3(8, 6, 6, with_colliders=False), 0.02, seed=1)
    sc.particles.velocity["vector"][:] = (3.0, 0.0, 0.0)
    e = _engine(oracle, sc)
    e.frame_begin(sc.params)
    e.stage_sort(); e.stage_reset_grid()
    dt = e.stage_estimate_dt(1e-5, 1 / 60)
    e.stage_g2p2g(dt)
    nodes = e.grid_nodes(1)
    m = float(sc.particles.volume["mass"].astype(np.float64).sum())
    assert float(nodes["mass"].astype(np.float64).sum()) == pytest.approx(m, rel=1e-6)
    assert float(nodes["psi_mass"].astype(np.float64).sum()) == pytest.approx(m, rel=1e-6)
    assert np.abs(e.read(abi.ATTR_VELOCITY)["vector"]).max() == 0.0
    assert np.array_equal(e.read(abi.ATTR_POSITION), sc.particles.position)
    e.stage_grid_update(dt, sc.gravity)
    v = e.grid_nodes(0)
    has_mass = v["mass"] > 0
    assert np.allclose(v["momentum_velocity"][has_mass][:, 1], -9.81 * dt, rtol=1e-5)   # v = g dt on every massive node
    assert (v["projection_status_tag"] == abi.PROJ_TOO_FAR).all()                       # no collider
    # second substep: momentum = sum m v, particles pick up g*dt
    e.stage_sort(); e.stage_reset_grid()
    dt2 = e.stage_estimate_dt(1e-5, 1 / 60 - dt)
    e.stage_g2p2g(dt2)
    pv = e.read(abi.ATTR_VELOCITY)["vector"]
    assert np.allclose(pv[:, 1], -9.81 * dt, rtol=1e-4) and np.abs(pv[:, [0, 2]]).max() < 1e-6
    nodes = e.grid_nodes(1)
    assert float(nodes["momentum_velocity"][..., 1].astype(np.float64).sum()) == pytest.approx(m * -9.81 * dt, rel=1e-4)


def test_dt_rules(oracle):
    """timestep.rs:87-112 and cuda_mpm_pipeline.rs:294-299: clamp to min_dt, accumulator is a running min, not reset per substep"""
    sc = scenes.sand3(4, 4, 4, with_colliders=False)
    e = _engine(oracle, sc)
    e.frame_begin(sc.params)
    dt = e.stage_estimate_dt(1e-5, 1 / 60)
    c = sc.models[0].constitutive_model
    bulk, shear = float(c.lam) + 2 * float(c.mu) / 3, float(c.mu)
    want = 0.9 * 0.2 / np.sqrt((bulk + 4 / 3 * shear) / 2700.0)
    assert dt == pytest.approx(want, rel=1e-5)
    assert e.stage_estimate_dt(1e-5, 1e-3) == pytest.approx(1e-3, rel=1e-6)   # max_dt wins
    assert e.stage_estimate_dt(1e-5, 1 / 60) == pytest.approx(1e-3, rel=1e-6) # running minimum (quirk 1)
    e.frame_begin(sc.params)
    assert e.stage_estimate_dt(5e-3, 1 / 60) == pytest.approx(5e-3, rel=1e-6)  # clamped up to min_dt
    # failed particles impose no bound
    sc.particles.status["failed"] = 1
    e = _engine(oracle, sc)
    e.frame_begin(sc.params)
    assert e.stage_estimate_dt(1e-5, 1e-3) == pytest.approx(1 / 60, rel=1e-6)
    # Neo-Hookean: cell_width and hardening swapped (quirk 4): cfl * hardening / max(|v|, sqrt((K + 4/3 G) * h / rho0))
    nh = scenes.neo_hookean_block(3, 3, 3, with_floor=False)
    e = _engine(oracle, nh)
    e.frame_begin(nh.params)
    c = nh.models[0].constitutive_model
    bulk, shear = (float(c.lam) + 2 * float(c.mu) / 3) * 0.2, float(c.mu) * 0.2
    want = 0.5 * 1.0 / np.sqrt((bulk + 4 / 3 * shear) / 1000.0)
    assert e.stage_estimate_dt(1e-7, 1.0) == pytest.approx(min(want, nh.params.dt), rel=1e-5)


def test_free_fall_and_floor(oracle):
    """a block in free fall follows g t^2 / 2 until it lands on the cuboid floor, then rests above it"""
    sc = scenes.sand3(6, 3, 6)
    e = _engine(oracle, sc)
    y0 = float(sc.particles.position["point"][:, 1].mean())
    t = 0.0
    for _ in range(6):
        sim, tm = e.step(sc.params, sc.gravity)
        t += sim
    y = float(e.read(abi.ATTR_POSITION)["point"][:, 1].mean())
    assert y0 - y == pytest.approx(0.5 * 9.81 * t * t, rel=0.12)  # symplectic Euler + first substep without velocity
    for _ in range(60):
        e.step(sc.params, sc.gravity)
    p = e.read(abi.ATTR_POSITION)["point"]
    assert np.isfinite(p).all() and p[:, 1].min() > 2.0  # floor top at y = 2
    st = e.read(abi.ATTR_STATUS)
    assert st["failed"].sum() == 0


def test_velocity_clamp_quirk(oracle):
    """particle_updater.rs:113-121: if ANY |v_i| dt >= h, ALL components become signum(v_i) h / dt (signum(0) = 1)"""
    sc = scenes.sand3(2, 2, 2, with_colliders=False)
    sc.particles.status["kinematic_vel_enabled"] = 1
    sc.particles.status["kinematic_vel"] = np.array([1.0e6, 0.0, -3.0], np.float32)
    e = _engine(oracle, sc)
    e.frame_begin(sc.params)
    e.stage_sort(); e.stage_reset_grid()
    dt = e.stage_estimate_dt(1e-5, 1e-3)
    e.stage_g2p2g(dt)
    v = e.read(abi.ATTR_VELOCITY)["vector"]
    lim = np.float32(0.2) / np.float32(dt)
    assert np.allclose(v, np.array([lim, lim, -lim]), rtol=1e-6)


def test_failed_particles_leave_the_grid(oracle):
    """particle_updater.rs:155-165 (quirk 6): failed particles advect but contribute no mass"""
    sc = scenes.sand3(4, 4, 4, with_colliders=False)
    sc.particles.status["failed"][:10] = 1
    e = _engine(oracle, sc)
    e.frame_begin(sc.params)
    e.stage_sort(); e.stage_reset_grid()
    e.stage_g2p2g(1e-3)
    m = float(e.grid_nodes(1)["mass"].astype(np.float64).sum())
    assert m == pytest.approx(float(sc.particles.volume["mass"][10:].astype(np.float64).sum()), rel=1e-6)
    vol = e.read(abi.ATTR_VOLUME)
    assert np.array_equal(vol["deformation_gradient"][:10], np.tile(np.eye(3, dtype=np.float32).reshape(9), (10, 1)))


def test_collider_cache_is_carried(oracle):
    """quirk 8: projection status / normal computed once per node and carried through K14 while the block stays active"""
    sc = scenes.sand3(4, 2, 4)
    sc.particles.position["point"][:, 1] -= 0.35  # start close to the floor so nodes are within 2h
    e = _engine(oracle, sc)
    e.step(sc.params, sc.gravity)
    k0, n0 = util.nodes_by_key(e, 0)
    assert (n0["projection_status_tag"] != abi.PROJ_NOT_COMPUTED).all()
    near = n0["projection_status_tag"] != abi.PROJ_TOO_FAR
    assert near.any() and np.allclose(np.linalg.norm(n0["collision_normal"][near], axis=-1), 1.0, atol=1e-5)
    e.frame_begin(sc.params)
    e.stage_sort(); e.stage_reset_grid()
    k1, n1 = util.nodes_by_key(e, 1)
    common, i0, i1 = np.intersect1d(k0, k1, return_indices=True)
    for f in ("projection_status_tag", "projection_status_payload", "collision_normal", "projection_scaled_dir"):
        assert np.array_equal(n0[f][i0], n1[f][i1]), f
    assert np.array_equal(n1["prev_mass"][i1], n0["mass"][i0])
    assert (n1["mass"] == 0).all()
