"""CPU suite of the dim2 build: the dim2 oracle (oracle/mpm_oracle2d.cpp) against known answers for every
dimension-dependent piece (block keys, stencil, 2x2 SVD, DIM-dependent model constants, parry2d projections) and against
size-independent properties (mass / momentum conservation, free fall, x <-> y transposition symmetry); the dim2 product library
loads, reports SPK_DIM == 2, exports every symbol of include/sparkl_abi.h and has no CPU fallback."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import binding as ob
from sparkl_b200 import abi, abi2d, scenes2d

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
f32 = np.float32


# ---------------------------------------------------------------------------------------------- ABI / library
def test_header_compiles_for_both_dimensions_and_pins_layouts():
    for dim in (2, 3):  # the static_asserts of the header pin every struct size / offset
        subprocess.check_call(["g++", "-std=c++17", f"-DSPK_DIM={dim}", "-fsyntax-only", "-x", "c++", os.path.join(ROOT, "include", "sparkl_abi.h")])
    assert abi2d.POSITION.itemsize == 8 and abi2d.VOLUME.itemsize == 32 and abi2d.STATUS.itemsize == 24 and abi2d.GRID_NODE.itemsize == 56
    assert abi2d.STATUS.fields["model_index"][1] == 16 and abi2d.GRID_NODE.fields["projection_status_tag"][1] == 24
    assert len(abi2d.CuboidCollider((1, 2)).to_gpu_collider()) == 72


def test_dim2_library_loads_exports_and_has_no_cpu_fallback():
    from sparkl_b200 import _lib
    if not os.path.exists(_lib.LIB2D_PATH):
        from sparkl_b200 import build
        build.build()
    l = _lib.lib(2)
    assert l.spk_dim() == 2 and _lib.lib(3).spk_dim() == 3
    for name in _lib.EXPORTS:
        assert hasattr(l, name), name
    import torch
    if not torch.cuda.is_available():
        h = C.c_void_p()
        rc = l.spk_ctx_create(0, C.c_float(0.1), 1024, C.byref(h))
        assert rc > 0 and not h.value  # a CUDA error code, never a CPU path


def test_dim2_sources_never_touch_the_oracle():
    for f in ("spk_g2p2g2d.cu", "spk_grid2d.cu", "spk_math2d.cuh"):
        assert "oracle" not in open(os.path.join(ROOT, "sparkl_b200", "csrc", f)).read()


# ---------------------------------------------------------------------------------------------- keys / stencil / svd
def test_block_key_dim2():
    """BlockVirtualId::pack (dim2), gpu_grid.rs:90-94: x | y << 32 with origin 2^31; block = floor((round(p / h) - 2) / 4)"""
    h = 0.25
    for p in [(0.0, 0.0), (0.49, 0.51), (-3.2, 7.7), (100.3, -55.1), (0.125, 0.375), (-0.125, -0.375)]:
        cells = [int(np.sign(f32(v) / f32(h)) * np.floor(np.abs(f32(v) / f32(h)) + f32(0.5))) for v in p]  # round half away
        blk = [(c - 2 + (1 << 31) * 4) // 4 for c in cells]
        want = (blk[0] & 0xFFFFFFFF) | ((blk[1] & 0xFFFFFFFF) << 32)
        assert ob.block_key_of_point2d(p, h) == want, p
    xy = abi2d.unpack_key(np.array([ob.block_key_of_point2d((1.0, -1.0), 0.25)], np.uint64))[0] - abi2d.PACK_ORIGIN
    assert tuple(xy) == (0, -2)  # cells (4, -4): blocks floor(2 / 4), floor(-6 / 4)


def test_svd2_reconstructs_and_matches_lapack():
    rng = np.random.default_rng(1)
    for i in range(500):
        F = rng.normal(size=(2, 2)).astype(np.float32)
        if i % 4 == 0:
            F = (np.eye(2) + 1e-3 * F).astype(np.float32)
        if i % 50 == 1:
            F[:, 1] = F[:, 0] * f32(0.5)  # rank one
        U, s, Vt = ob.svd2(F)
        assert np.abs(U @ np.diag(s) @ Vt - F).max() <= 2e-6 * max(1.0, np.abs(F).max())
        assert np.abs(U @ U.T - np.eye(2)).max() <= 1e-6 and np.abs(Vt @ Vt.T - np.eye(2)).max() <= 1e-6 and (s >= 0).all()
        assert np.allclose(np.sort(s), np.sort(np.linalg.svd(F.astype(np.float64), compute_uv=False)), atol=3e-6)


# ---------------------------------------------------------------------------------------------- models (DIM-dependent constants)
E, NU = 1.0e5, 0.3


def _lame():
    return [float(v) for v in abi.lame_lambda_mu(E, NU)]


def test_corotated_stress_dim2_diagonal_F():
    lam, mu = _lame()
    a, b = 1.1, 0.95
    tau = ob.kirchhoff_stress2d(abi.ParticleModel(abi.CorotatedLinearElasticity(E, NU)), np.diag([a, b]))
    j = a * b
    want = np.diag([2 * mu * (a - 1) * a + lam * (j - 1) * j, 2 * mu * (b - 1) * b + lam * (j - 1) * j])
    assert np.abs(tau - want).max() <= 2e-5 * np.abs(want).max()


def test_corotated_stress_dim2_is_rotation_equivariant():
    model = abi.ParticleModel(abi.CorotatedLinearElasticity(E, NU))
    F = np.array([[1.05, 0.02], [-0.03, 0.97]])
    t = 0.7
    R = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]])
    tau, tau_r = ob.kirchhoff_stress2d(model, F), ob.kirchhoff_stress2d(model, R @ F)
    assert np.abs(tau_r - R @ tau @ R.T).max() <= 5e-5 * np.abs(tau).max()


def test_neo_hookean_stress_dim2_uses_dim_in_exponent_and_trace():
    """elasticity_neo_hookean.rs:43-67 with DIM = 2: J^(-2/2) dev_2(F F^T) (trace / 2), k = 2/3 mu + lambda (literal)"""
    lam, mu = _lame()
    F = np.array([[1.2, 0.1], [0.05, 0.9]])
    j = np.linalg.det(F)
    cg = F @ F.T
    dev = mu * j ** (-1.0) * (cg - np.trace(cg) / 2 * np.eye(2))
    vol = (2.0 / 3.0 * mu + lam) / 2 * (j * j - 1) * np.eye(2)
    tau = ob.kirchhoff_stress2d(abi.ParticleModel(abi.NeoHookeanElasticity(E, NU)), F)
    assert np.abs(tau - (dev + vol)).max() <= 2e-5 * np.abs(dev + vol).max()


def test_monaghan_stress_dim2():
    m = abi.ParticleModel(abi.MonaghanSphEos(1.0e4, 7, 0.5))
    F = np.diag([0.9, 1.0])  # fluid: F[0, 0] = J
    G = np.array([[0.3, 0.2], [-0.1, -0.4]])
    tau = ob.kirchhoff_stress2d(m, F, jp=1.0, mass=2.0, volume0=2.0, vel_grad=G)
    p = max(1.0e4 * ((1 / 0.9) ** 7 - 1), -1.0)
    sr = (G + G.T) / 2
    want = -p * np.eye(2) + 2 * 0.5 * (sr - np.trace(sr) / 2 * np.eye(2))
    assert np.abs(tau - want).max() <= 2e-5 * np.abs(want).max()


def _data():
    d = np.zeros(1, abi2d.DATA)
    d["nacc_alpha"], d["plastic_hardening"], d["elastic_hardening"] = -0.01, 1.0, 1.0
    return d


def test_drucker_prager_dim2():
    model = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(E, NU), abi.DruckerPragerPlasticity(E, NU))
    # expansion: projected to the tip, singular values 1, Jp *= det, hardening += |log strain|
    F2, jp, d = ob.plastic_update2d(model, np.diag([1.1, 1.05]), _data())
    assert np.abs(F2 - np.eye(2)).max() <= 1e-6 and abs(jp - 1.1 * 1.05) <= 1e-5
    assert abs(float(d["plastic_hardening"][0]) - (1 + np.hypot(np.log(1.1), np.log(1.05)))) <= 1e-5
    # compression with shear: return mapping of plasticity_drucker_prager.rs:36-62 with d = 2, evaluated in numpy
    lam, mu = _lame()
    s = np.array([0.9, 0.8])
    dp = model.plastic_model
    angle = float(dp.h0) + (float(dp.h1) * 1.0 - float(dp.h3)) * np.exp(-float(dp.h2) * 1.0)
    alpha = np.sqrt(2 / 3) * 2 * np.sin(angle) / (3 - np.sin(angle))
    strain = np.log(s)
    dev = strain - strain.sum() / 2
    gamma = np.linalg.norm(dev) + (2 * lam + 2 * mu) / (2 * mu) * strain.sum() * alpha
    F2, jp, d = ob.plastic_update2d(model, np.diag(s), _data())
    if gamma <= 0:
        assert np.abs(F2 - np.diag(s)).max() <= 1e-6
    else:
        want = np.exp(strain - gamma * dev / np.linalg.norm(dev))
        assert np.abs(np.sort(np.diag(F2)) - np.sort(want)).max() <= 2e-5
        assert abs(float(d["plastic_hardening"][0]) - (1 + gamma)) <= 2e-5


def test_snow_and_rankine_dim2():
    snow = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(E, NU), abi.SnowPlasticity())
    F2, jp, d = ob.plastic_update2d(snow, np.diag([1.2, 0.9]), _data())
    lo, hi = 1 - 2.5e-2, 1 + 4.5e-3
    assert np.abs(np.diag(F2) - [hi, lo]).max() <= 1e-6 and abs(jp - 1.2 * 0.9 / (hi * lo)) <= 1e-5
    assert abs(float(d["elastic_hardening"][0]) - np.exp(10 * (1 - jp))) <= 1e-4
    # Rankine, plasticity_rankine.rs:35-78 with idx = [0, 1, DIM - 1]: pure tension along x beyond the strength
    lam, mu = _lame()
    rk = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(E, NU), abi.RankinePlasticity(E, NU, 1.0e2, 5.0))
    for s in (np.array([1.05, 1.0]), np.array([1.0, 1.05]), np.array([1.04, 1.05]), np.array([1.05, 1.05])):
        e = np.log(np.sort(s)[::-1])      # nalgebra's 2x2 SVD returns the singular values in descending order
        # idx = [0, 1, 1] sorted by eigenvalue -> (e3, e2, e1): (1, 1, 0) when e[0] > e[1], (0, 1, 1) when they are equal
        e1, e2 = (0, 1) if e[0] > e[1] else (1, 1)
        soft = 1.0e2
        F2, _, d = ob.plastic_update2d(rk, np.diag(s), _data())
        if lam * e.sum() + 2 * mu * e[e1] <= soft:
            want = e
        elif (2 * mu + lam) * e[e2] + lam * (e.sum() - e[e1]) <= soft:
            want = e.copy()
            want[e1] = (soft - lam * (e.sum() - e[e1])) / (2 * mu + lam)
        else:
            want = np.full(2, soft / (2 * mu + 3 * lam))
        assert np.abs(np.sort(np.diag(F2)) - np.sort(np.exp(want))).max() <= 2e-5, s
        assert abs(float(d["plastic_hardening"][0]) - min(1 + 5.0 * np.linalg.norm(e - want), 1.0e2)) <= 1e-4


def test_nacc_friction_constant_depends_on_dim():
    a3 = abi.NaccPlasticity(E, NU, 0.04, True, 3.0, np.radians(40.0))
    a2 = abi2d.NaccPlasticity(E, NU, 0.04, True, 3.0, np.radians(40.0))
    sf = np.sin(np.radians(40.0))
    assert abs(float(a2.friction) - np.sqrt(2 / 3) * 2 * sf / (3 - sf) * 2 / np.sqrt(2 / 4)) <= 1e-5
    assert abs(float(a3.friction) - np.sqrt(2 / 3) * 2 * sf / (3 - sf) * 3 / np.sqrt(2 / 3)) <= 1e-5


# ---------------------------------------------------------------------------------------------- colliders (parry2d)
def test_cuboid_projection_dim2_against_brute_force():
    rng = np.random.default_rng(5)
    col = abi2d.CuboidCollider((1.5, 0.5), translation=(0.3, -0.2), rotation=0.6)
    c, s = np.cos(0.6), np.sin(0.6)
    R = np.array([[c, -s], [s, c]])
    # dense sampling of the boundary in local coordinates
    t = np.linspace(-1, 1, 4001)
    edge = np.concatenate([np.stack([1.5 * t, np.full_like(t, 0.5)], 1), np.stack([1.5 * t, np.full_like(t, -0.5)], 1),
                           np.stack([np.full_like(t, 1.5), 0.5 * t], 1), np.stack([np.full_like(t, -1.5), 0.5 * t], 1)])
    world_edge = edge @ R.T + np.array([0.3, -0.2])
    for _ in range(200):
        p = rng.uniform(-3, 3, 2)
        proj, inside = ob.collider_project2d(col, p, 1.0)
        lp = R.T @ (p - np.array([0.3, -0.2]))
        assert inside == bool(abs(lp[0]) <= 1.5 and abs(lp[1]) <= 0.5)
        d = np.linalg.norm(world_edge - p, axis=1).min()
        assert abs(np.linalg.norm(proj - p) - d) <= 2e-3


def test_polyline_projection_dim2_inside_outside():
    rng = np.random.default_rng(6)
    poly = np.array([[0, 0], [2, 0], [2, 1], [1, 0.4], [0, 1]], np.float32)  # counter-clockwise, one reentrant vertex
    col = abi2d.PolylineCollider(poly)
    flipped = abi2d.PolylineCollider(poly, flip_interior=True)
    inside_ref = scenes2d._points_in_polygon
    n_checked = 0
    for _ in range(400):
        p = rng.uniform(-0.5, 2.5, 2).astype(np.float32)
        proj, inside = ob.collider_project2d(col, p, 1.0)
        # distance to the polygon boundary
        d = min(_seg_dist(p, poly[i], poly[(i + 1) % len(poly)]) for i in range(len(poly)))
        assert abs(np.linalg.norm(proj - p) - d) <= 1e-5
        if d > 1e-3:
            assert inside == bool(inside_ref(p[None, :], poly)[0]), p
            assert ob.collider_project2d(flipped, p, 1.0)[1] == (not inside)
            n_checked += 1
    assert n_checked > 300


def _tri_mesh_2d():
    """an L-shaped region made of four triangles (internal edges included) + a detached triangle"""
    v = np.array([[0.03, 0.02], [2.04, 0.05], [2.01, 1.03], [0.02, 1.01], [1.02, 1.02], [1.03, 2.05], [0.01, 2.02], [3.1, 3.2], [4.2, 3.1], [3.6, 4.3]], np.float32)
    t = np.array([[0, 1, 2], [0, 2, 3], [3, 4, 5], [3, 5, 6], [7, 8, 9]], np.uint32)
    return v, t


def test_trimesh_projection_dim2_against_brute_force():
    """dim2 TriMesh (parry2d: a soup of solid triangles; the closest triangle decides, a point inside a triangle goes to that triangle's
    nearest edge and is flagged inside) against an independent f64 computation"""
    v, t = _tri_mesh_2d()
    col = abi2d.TriMeshCollider(v, t, translation=(0.4, -0.3), rotation=0.35)
    flipped = abi2d.TriMeshCollider(v, t, translation=(0.4, -0.3), rotation=0.35, flip_interior=True)
    rng = np.random.Generator(np.random.PCG64(17))
    c, s_ = np.cos(np.float32(0.35)), np.sin(np.float32(0.35))
    V = v.astype(np.float64)

    def seg_dist(p, a, b):
        ab, ap = b - a, p - a
        u = np.clip(ab @ ap / (ab @ ab), 0.0, 1.0)
        return np.linalg.norm(p - (a + u * ab)), a + u * ab

    def line_dist(p, a, b):
        ab, ap = b - a, p - a
        u = ab @ ap / (ab @ ab)
        return np.linalg.norm(p - (a + u * ab)), a + u * ab

    checked_flags = 0
    for _ in range(600):
        lp = rng.uniform([-0.8, -0.8], [4.8, 4.8])
        per_tri = []
        for tri in t:
            a, b, cc = V[tri[0]], V[tri[1]], V[tri[2]]
            m = np.array([b - a, cc - a]).T
            l1, l2 = np.linalg.solve(m, lp - a)
            inside = l1 > 0 and l2 > 0 and l1 + l2 < 1
            edges = [(a, b), (b, cc), (cc, a)]
            if inside:
                d, q = min((line_dist(lp, *e) for e in edges), key=lambda x: x[0])
            else:
                d, q = min((seg_dist(lp, *e) for e in edges), key=lambda x: x[0])
            per_tri.append((d, q, inside))
        order = sorted(range(len(t)), key=lambda k: per_tri[k][0])
        d0, q0, in0 = per_tri[order[0]]
        wp = np.array([c * lp[0] - s_ * lp[1] + 0.4, s_ * lp[0] + c * lp[1] - 0.3])
        got = ob.collider_project2d(col, wp.astype(np.float32), 1.0)
        assert got is not None
        gq, gin = got
        wq = np.array([c * q0[0] - s_ * q0[1] + 0.4, s_ * q0[0] + c * q0[1] - 0.3])
        assert abs(np.linalg.norm(gq - wp) - d0) < 2e-5, (lp, gq, wq)       # the distance is that of the closest triangle
        if per_tri[order[1]][0] - d0 > 1e-4:                               # an unambiguous winner (not a shared edge): point and flag too
            assert np.abs(gq - wq).max() < 5e-5, (lp, gq, wq)
            assert gin == bool(in0)
            assert ob.collider_project2d(flipped, wp.astype(np.float32), 1.0)[1] == (not in0)
            checked_flags += 1
    assert checked_flags > 300


def test_heightfield_projection_dim2_against_brute_force():
    """parry2d HeightField (examples2d/basic2.rs ground): closest point on the polyline of samples within max_dist, inside = under
    the surface and within the x footprint (gpu_collider.rs:73-99)"""
    rng = np.random.default_rng(16)
    g = scenes2d.basic2_ground()
    flipped = abi2d.HeightFieldCollider(g.heights, g.scale, translation=g.translation, flip_interior=True)
    n = len(g.heights)
    xs = (-0.5 + np.arange(n) / (n - 1)) * g.scale[0] + g.translation[0]
    ys = g.heights.astype(np.float64) * g.scale[1] + g.translation[1]
    pts = np.stack([xs, ys], axis=1)
    hits = 0
    for _ in range(400):
        px = rng.uniform(-0.8, 1.8)
        p = np.array([px, np.interp(px, xs, ys) + rng.uniform(-0.2, 0.2)])  # around the surface (and beyond both ends of the field)
        max_dist = 0.08
        r = ob.collider_project2d(g, p, max_dist)
        d = min(_seg_dist(p, pts[i], pts[i + 1]) for i in range(n - 1))
        if d > max_dist * (1 + 1e-4) or p[0] < xs[0] - max_dist or p[0] > xs[-1] + max_dist:
            assert r is None or abs(d - max_dist) <= 1e-4, (p, d)
            continue
        if d > max_dist * (1 - 1e-4):
            continue
        assert r is not None, (p, d)
        proj, inside = r
        hits += 1
        assert abs(np.linalg.norm(proj - p) - d) <= 2e-6
        surface_y = np.interp(p[0], xs, ys)
        if abs(p[1] - surface_y) > 1e-3 and xs[0] + 1e-3 < p[0] < xs[-1] - 1e-3 and abs(p[1] - proj[1]) > 1e-4:
            assert inside == bool(p[1] <= proj[1])  # the reference compares with the PROJECTION's y, not the surface height at p.x
            assert ob.collider_project2d(flipped, p, max_dist)[1] == bool(p[1] >= proj[1])
    assert hits > 40


def _seg_dist(p, a, b):
    ab, ap = b - a, p - a
    u = np.clip(np.dot(ab, ap) / np.dot(ab, ab), 0, 1)
    return float(np.linalg.norm(p - (a + u * ab)))


# ---------------------------------------------------------------------------------------------- whole-step properties
def _oracle(scene):
    o = ob.OracleEngine2D(scene.cell_width)
    o.models_upload(scene.models)
    o.colliders_upload(scene.colliders)
    o.upload(scene.particles)
    return o


def test_sort_invariants_dim2():
    sc = scenes2d.jitter(scenes2d.elasticity2(nblocks=2, scale=0.3), 0.01, seed=2)
    o = _oracle(sc)
    info = o.stage_sort()
    hdr = o.grid_headers(1)
    ids = o.read(abi.ATTR_SORTED_IDS)
    assert np.array_equal(np.sort(ids), np.arange(len(sc.particles)))
    assert np.array_equal(hdr["first_particle"], np.concatenate([[0], np.cumsum(hdr["num_particles"])[:-1]]))
    keys = np.array([ob.block_key_of_point2d(p, sc.cell_width) for p in sc.particles.position["point"]], np.uint64)
    for hrow in hdr:
        own = ids[hrow["first_particle"]: hrow["first_particle"] + hrow["num_particles"]]
        assert (keys[own] == hrow["virtual_id"]).all()
    # active set = particle blocks + {0,1}^2 (gpu_grid.rs:293-304)
    xy = abi2d.unpack_key(np.unique(keys))
    want = {(int(x + ox), int(y + oy)) for x, y in xy for ox in (0, 1) for oy in (0, 1)}
    got = {tuple(int(v) for v in r) for r in abi2d.unpack_key(hdr["virtual_id"])}
    assert got == want and info.num_active_blocks == len(want)
    assert info.num_dispatch_blocks == int(sum((int(n) + 63) // 64 for n in hdr["num_particles"]))  # G2P2G_THREADS = 64 in dim2


def test_p2g_conserves_mass_and_momentum_dim2():
    sc = scenes2d.jitter(scenes2d.model_block2("corotated", 10, 9), 0.02, seed=4)
    sc.colliders = []
    o = _oracle(sc)
    o.frame_begin(sc.params)
    o.substep(sc.params, sc.gravity, 1.0 / 60)           # first substep: velocities come from the (empty) grid
    o.stage_sort(); o.stage_reset_grid()
    dt = o.stage_estimate_dt(0.0, 1.0e-3)
    pv = o.read(abi.ATTR_VELOCITY)["vector"].astype(np.float64)
    o.stage_g2p2g(dt)
    nodes = o.grid_nodes(1)
    m = sc.particles.volume["mass"].astype(np.float64)
    assert abs(nodes["mass"].astype(np.float64).sum() - m.sum()) <= 1e-6 * m.sum()
    # APIC: grid momentum after P2G == particle momentum after G2P (affine and stress terms sum to zero over the stencil)
    v_after = o.read(abi.ATTR_VELOCITY)["vector"].astype(np.float64)
    mom_p = (v_after * m[:, None]).sum(axis=0)
    mom_g = nodes["momentum_velocity"].astype(np.float64).reshape(-1, 2).sum(axis=0)
    assert np.abs(mom_g - mom_p).max() <= 1e-4 * max(np.abs(mom_p).max(), m.sum() * 1e-3), (mom_g, mom_p, pv.mean(axis=0))


def test_free_fall_dim2():
    sc = scenes2d.elasticity2(nblocks=2, scale=0.3, with_colliders=False)
    o = _oracle(sc)
    y0 = sc.particles.position["point"][:, 1].astype(np.float64).mean()
    T = 0.0
    for _ in range(12):
        sim, _t = o.step(sc.params, sc.gravity)
        T += sim
    v = o.read(abi.ATTR_VELOCITY)["vector"].astype(np.float64)
    y = o.read(abi.ATTR_POSITION)["point"][:, 1].astype(np.float64).mean()
    assert abs(v[:, 1].mean() + 9.81 * T) <= 2e-3 * 9.81 * T
    assert abs((y - y0) + 0.5 * 9.81 * T * T) <= 0.02 * 0.5 * 9.81 * T * T


def _transpose_scene(sc, fluid=False):
    """swap the x and y axes of a scene (positions, velocities, F -> P F P, colliders, gravity); a fluid particle keeps its
    F untouched: F[0, 0] is the volume ratio J there (particle.rs:89-91), not a tensor component"""
    import copy
    t = copy.deepcopy(sc)
    t.particles.position["point"] = sc.particles.position["point"][:, ::-1].copy()
    t.particles.velocity["vector"] = sc.particles.velocity["vector"][:, ::-1].copy()
    F = sc.particles.volume["deformation_gradient"].reshape(-1, 2, 2)  # [n, col, row]
    if not fluid:
        t.particles.volume["deformation_gradient"] = F[:, ::-1, ::-1].reshape(-1, 4).copy()
    t.colliders = [abi2d.CuboidCollider(c.half_extents[::-1], c.translation[::-1], 0.0, c.friction, c.penalty_stiffness, c.grid_boundary_handling)
                   for c in sc.colliders]
    t.gravity = sc.gravity[::-1]
    return t


@pytest.mark.parametrize("kind", ["corotated", "sand", "fluid", "nacc"])
def test_transposition_symmetry_dim2(kind):
    """the dim2 pipeline must not care which axis is called x: running the x <-> y mirrored scene gives the mirrored state.
    Catches index mix-ups in keys, quadrants, tile layout, NBH order and the 2x2 algebra."""
    sc = scenes2d.jitter(scenes2d.model_block2(kind, 9, 8), 0.01, seed=9)
    a, b = _oracle(sc), _oracle(_transpose_scene(sc, fluid=kind == "fluid"))
    for _ in range(2):
        a.step(sc.params, sc.gravity)
        b.step(sc.params, sc.gravity[::-1])
    pa, pb = a.read_all(), b.read_all()
    assert np.abs(pa.position["point"] - pb.position["point"][:, ::-1]).max() <= 2e-5 * sc.cell_width
    va = pa.velocity["vector"]
    assert np.abs(va - pb.velocity["vector"][:, ::-1]).max() <= 1e-3 * max(np.abs(va).max(), 1e-2)
    Fa = pa.volume["deformation_gradient"].reshape(-1, 2, 2)
    Fb = pb.volume["deformation_gradient"].reshape(-1, 2, 2)[:, ::-1, ::-1]
    if kind != "fluid":
        assert np.abs(Fa - Fb).max() <= 5e-5
    assert np.array_equal(pa.status["failed"], pb.status["failed"])


def test_stop_after_one_substep_and_zero_dt_dim2():
    sc = scenes2d.model_block2("corotated", 6, 6)
    o = _oracle(sc)
    p = abi.SolverParameters(stop_after_one_substep=1)
    sim, t = o.step(p, sc.gravity)
    assert t.num_substeps == 1 and 0 < sim < p.dt
    sim, t = o.step(abi.SolverParameters(dt=0.0), sc.gravity)
    assert sim == 0.0 and t.num_substeps == 0


# ---------------------------------------------------------------------------------------------- the DEVICE math, compiled for the host
# sparkl_b200/csrc/spk_math2d.cuh is what the dim2 CUDA kernel inlines (one shared SVD, DevModel parameter blocks).  It is compiled
# here with g++ (tests/host/math2d_host.cpp) and compared with the oracle, so the kernel's arithmetic is checked without a GPU.
from tests.host.emu2d import DEVMODEL, devmodel as _devmodel  # noqa: E402


@pytest.fixture(scope="module")
def hostmath():
    from tests.host.emu2d import math_lib
    return math_lib()


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _random_F(rng, i):
    F = np.eye(2) + rng.uniform(-0.15, 0.15, (2, 2)) * np.array([[1.0, 0.4], [0.4, 1.0]])
    if i % 7 == 0:
        t = rng.uniform(0, 2 * np.pi)
        F = np.array([[np.cos(t), -np.sin(t)], [np.sin(t), np.cos(t)]]) @ F
    if i % 11 == 0:
        F = np.diag(rng.uniform(0.8, 1.2, 2))
    return F.astype(np.float32)


def test_device_math_svd2(hostmath):
    rng = np.random.default_rng(3)
    for i in range(1000):
        F = _random_F(rng, i) if i % 3 else rng.normal(size=(2, 2)).astype(np.float32)
        if i % 97 == 5:
            F[:, 1] = F[:, 0] * f32(-0.25)  # rank one
        f = np.ascontiguousarray(F.T)
        u, s, v = np.zeros(4, np.float32), np.zeros(2, np.float32), np.zeros(4, np.float32)
        hostmath.hm_svd2(_vp(f), _vp(u), _vp(s), _vp(v))
        U, V = u.reshape(2, 2).T, v.reshape(2, 2).T
        assert s[0] >= s[1] >= 0  # nalgebra's svd_ordered2 order
        assert np.abs(U @ np.diag(s) @ V.T - F).max() <= 4e-6 * max(1.0, np.abs(F).max())
        assert np.abs(U @ U.T - np.eye(2)).max() <= 2e-6 and np.abs(V @ V.T - np.eye(2)).max() <= 2e-6
        assert np.allclose(s, ob.svd2(F)[1], atol=4e-6 * max(1.0, np.abs(F).max()))


@pytest.mark.parametrize("kind", ["sand", "nacc", "rankine", "snow"])
def test_device_math_plastic_update_matches_oracle(hostmath, kind):
    rng = np.random.default_rng(8)
    model = scenes2d.model_block2(kind, 2, 2).models[0]
    dm = _devmodel(model)
    changed = 0
    for i in range(600):
        F = _random_F(rng, i)
        d = _data()
        d["plastic_hardening"] = 1.0 + (rng.uniform(0, 0.5) if i % 2 else 0.0)
        d["log_vol_gain"] = rng.uniform(-0.02, 0.02) if i % 3 == 0 else 0.0
        d["nacc_alpha"] = rng.uniform(-0.05, 0.02)
        jp0 = float(rng.uniform(0.9, 1.1))
        Fo, jpo, do = ob.plastic_update2d(model, F, d, jp=jp0)
        st = np.concatenate([np.ascontiguousarray(F.T).ravel(), [jp0, 1.0, d["elastic_hardening"][0], d["plastic_hardening"][0], d["log_vol_gain"][0],
                                                                  d["nacc_alpha"][0]]]).astype(np.float32)
        changed += hostmath.hm_plastic(_vp(dm), _vp(st))
        Fd = st[:4].reshape(2, 2).T
        sv = np.linalg.svd(F.astype(np.float64), compute_uv=False)
        if kind in ("rankine", "snow") and abs(sv[0] - sv[1]) < 2e-3 * sv[0]:
            continue  # coalescing singular values: U, V ill-defined, f32 vs f64 recompositions differ (tests/util.py ROBUST_MODELS)
        assert np.abs(Fd - Fo).max() <= 3e-5, (i, F, Fd, Fo)
        assert abs(st[4] - jpo) <= 3e-5 and abs(st[6] - do["elastic_hardening"][0]) <= 3e-4 * max(1.0, abs(do["elastic_hardening"][0]))
        assert abs(st[7] - do["plastic_hardening"][0]) <= 3e-5 * max(1.0, abs(do["plastic_hardening"][0]))
        assert abs(st[8] - do["log_vol_gain"][0]) <= 3e-5 and abs(st[9] - do["nacc_alpha"][0]) <= 1e-4
    assert changed > 100  # the return mappings actually fired


@pytest.mark.parametrize("kind", ["corotated", "neo_hookean", "failure_split"])
def test_device_math_stress_matches_oracle(hostmath, kind):
    rng = np.random.default_rng(12)
    model = scenes2d.model_block2(kind, 2, 2).models[0]
    dm = _devmodel(model)
    for i in range(600):
        F = _random_F(rng, i)
        phase = [1.0, 0.0, 0.5][i % 3]
        hard = float(rng.uniform(0.8, 1.5))
        want = ob.kirchhoff_stress2d(model, F, phase=phase, hardening=hard)
        out = np.zeros(4, np.float32)
        hostmath.hm_stress(_vp(dm), _vp(np.ascontiguousarray(F.T)), phase, hard, _vp(out))
        got = out.reshape(2, 2).T
        lam_mu = float(model.constitutive_model.lam) + 2 * float(model.constitutive_model.mu)
        assert np.abs(got - want).max() <= 5e-5 * np.abs(want).max() + 2e-6 * lam_mu, (i, got, want)


def test_device_math_dt_candidate_matches_oracle(hostmath):
    rng = np.random.default_rng(13)
    for kind in ("corotated", "neo_hookean", "fluid"):
        sc = scenes2d.model_block2(kind, 1, 1)
        dm = _devmodel(sc.models[0])
        for i in range(50):
            pc = sc.particles.take(np.array([0]))
            v = rng.uniform(-30, 30, 2).astype(np.float32) * (0.0 if i == 0 else 1.0)
            pc.velocity["vector"][0] = v
            pc.volume["deformation_gradient"][0, 0] = f32(rng.uniform(0.9, 1.1))
            pc.data["elastic_hardening"][0] = f32(rng.uniform(0.8, 1.3))
            o = ob.OracleEngine2D(sc.cell_width)
            o.models_upload(sc.models); o.upload(pc)
            o.set_dt_acc(2 ** 63)
            o.stage_estimate_dt(0.0, 1.0e6)
            cand = hostmath.hm_dt_candidate(_vp(dm), sc.cell_width, float(pc.volume["mass"][0]), float(pc.volume["volume0"][0]),
                                            float(pc.volume["deformation_gradient"][0, 0]), float(pc.data["elastic_hardening"][0]), float(v[0]), float(v[1]))
            want = o.dt_acc() / 1e12
            if np.isnan(cand):
                cand = 1.0e6  # Rust's f32::min ignores NaN (the kernel maps NaN to "no bound" the same way)
            assert abs(min(cand, 1.0e6) - want) <= 2e-5 * want, (kind, i, cand, want)


def test_device_math_eigen_extrema(hostmath):
    rng = np.random.default_rng(14)
    for _ in range(200):
        a = rng.normal(size=(2, 2)).astype(np.float32)
        a[0, 1] = 123.0  # upper triangle is ignored (lower triangle read, like nalgebra SymmetricEigen)
        mn, mx = C.c_float(), C.c_float()
        hostmath.hm_eig_minmax(_vp(np.ascontiguousarray(a.T)), C.byref(mn), C.byref(mx))
        sym = np.array([[a[0, 0], a[1, 0]], [a[1, 0], a[1, 1]]], np.float64)
        ev = np.linalg.eigvalsh(sym)
        assert abs(mn.value - ev[0]) <= 1e-5 * max(1, np.abs(ev).max()) and abs(mx.value - ev[1]) <= 1e-5 * max(1, np.abs(ev).max())


# ---------------------------------------------------------------------------------------------- the DEVICE kernels, emulated on the host
# tests/host/kernels2d_host.cpp compiles the product's dim2 kernel sources (G2P2G particle body, reset, grid update with colliders,
# export) with g++; tests/host/emu2d.py adds the numpy restatement of the sort and of the host sequencing.  The bodies of the GPU
# parity tests (tests/test_dim2_gpu.py) are then run with that emulation in place of the CUDA context: same comparisons, same
# tolerances, no GPU.  What this leaves to the GPU run: the launch glue (TMA staging / bulk-reduce flush, barriers, dispatch
# indexing) and the CUDA sort kernels, which the dim2 library shares with the GPU-verified 3D build.
def _host_engines(scene):
    from tests.host.emu2d import Emu2D
    g = Emu2D(scene.cell_width)
    g.models_upload(scene.models)
    g.colliders_upload(scene.colliders)
    g.upload(scene.particles)
    return g, _oracle(scene)


@pytest.fixture()
def gpu_test_bodies(monkeypatch):
    import tests.test_dim2_gpu as T
    monkeypatch.setattr(T, "_engines", _host_engines)
    return T


def test_device_kernels_on_host_elasticity2(gpu_test_bodies):
    gpu_test_bodies.test_elasticity2_lockstep()
    gpu_test_bodies.test_elasticity2_on_the_ground()


@pytest.mark.parametrize("kind", ["corotated", "neo_hookean", "fluid", "sand", "nacc", "failure", "failure_split"])
def test_device_kernels_on_host_models(gpu_test_bodies, kind):
    gpu_test_bodies.test_model_block2_lockstep(kind)


@pytest.mark.parametrize("kind", ["rankine", "snow"])
def test_device_kernels_on_host_clamping_models(gpu_test_bodies, kind):
    gpu_test_bodies.test_model_block2_lockstep_clamping_models(kind)


def test_device_kernels_on_host_fluid_and_polyline(gpu_test_bodies):
    gpu_test_bodies.test_fluids2_dam_break_lockstep()
    gpu_test_bodies.test_polyline_trough_dim2()


def test_device_kernels_on_host_basic2_heightfield_ground(gpu_test_bodies):
    gpu_test_bodies.test_basic2_on_the_heightfield_ground()


def test_device_kernels_on_host_roundtrip_and_kinematic():
    """pack / unpack mirror + kinematic and static particles through the emulated kernel"""
    sc = scenes2d.jitter(scenes2d.model_block2("corotated", 10, 8), 0.01, seed=5)
    sc.particles.status["kinematic_vel_enabled"][:10] = 1
    sc.particles.status["kinematic_vel"][:10] = (0.5, 0.25)
    sc.particles.status["is_static"][10:20] = 1
    sc.particles.status["failed"][20:24] = 1
    g, o = _host_engines(sc)
    back = g.read_all()
    for name in ("position", "velocity", "volume", "status", "phase", "data"):
        assert np.array_equal(getattr(back, name), getattr(sc.particles, name)), name
    import tests.test_dim2_gpu as T
    g.frame_begin(sc.params); o.frame_begin(sc.params)
    rem = float(f32(sc.params.dt))
    for s in range(4):
        rem -= T._substep_both(g, o, sc, rem, T.RTOL_1 * (1 + 3 * s), f"substep {s}")
    assert np.array_equal(g.read(abi.ATTR_STATUS)["failed"], o.read(abi.ATTR_STATUS)["failed"])
