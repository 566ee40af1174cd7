"""HeightField colliders: the oracle's restatement of parry3d 0.13's
`HeightField::project_local_point_with_max_dist` + sparkl's inside test (src_kernels/gpu_collider.rs:73-99) against an
independent float64 brute force over ALL triangles of the field.  parry3d is not under /root/reference and its device
layout (`CudaHeightFieldPtr`) was not recovered from the PTX, so this shape is NOT pinned against the reference's own
kernels: what is pinned here is the geometry the published algorithm defines — the closest point of the
triangulated surface within max_dist, None beyond it."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes


def _brute_force(hf: abi.HeightFieldCollider, pt):
    """closest point on the triangulated surface in float64 (identity rotation), scanning every triangle"""
    H = hf.heights.astype(np.float64)
    nr, nc = H.shape
    sx, sy, sz = (float(v) for v in hf.scale)
    lp = np.asarray(pt, np.float64) - np.asarray(hf.translation, np.float64)
    zs = (-0.5 + np.arange(nr) / (nr - 1)) * sz
    xs = (-0.5 + np.arange(nc) / (nc - 1)) * sx
    best, bestq = np.inf, None
    for j in range(nc - 1):
        for i in range(nr - 1):
            p00 = np.array([xs[j], H[i, j] * sy, zs[i]]); p10 = np.array([xs[j], H[i + 1, j] * sy, zs[i + 1]])
            p01 = np.array([xs[j + 1], H[i, j + 1] * sy, zs[i]]); p11 = np.array([xs[j + 1], H[i + 1, j + 1] * sy, zs[i + 1]])
            for a, b, c in ((p00, p10, p01), (p10, p11, p01)):
                q = _closest_on_triangle(lp, a, b, c)
                d = np.linalg.norm(lp - q)
                if d < best:
                    best, bestq = d, q
    return best, bestq + np.asarray(hf.translation, np.float64), lp


def _closest_on_segment(p, a, b):
    ab = b - a
    t = np.clip(np.dot(p - a, ab) / max(np.dot(ab, ab), 1e-300), 0.0, 1.0)
    return a + t * ab


def _closest_on_triangle(p, a, b, c):
    """plane projection + barycentric test, else the best of the three edges (not the Voronoi-region walk of the oracle)"""
    n = np.cross(b - a, c - a)
    nn = np.dot(n, n)
    if nn > 0:
        q = p - n * (np.dot(p - a, n) / nn)
        u = np.dot(np.cross(b - a, q - a), n); v = np.dot(np.cross(c - b, q - b), n); w = np.dot(np.cross(a - c, q - c), n)
        if u >= 0 and v >= 0 and w >= 0:
            return q
    cands = [_closest_on_segment(p, a, b), _closest_on_segment(p, b, c), _closest_on_segment(p, c, a)]
    return min(cands, key=lambda q: np.linalg.norm(p - q))


def test_heightfield_projection_matches_brute_force(oracle):
    rng = np.random.Generator(np.random.PCG64(31))
    heights = rng.uniform(-1.0, 1.0, (7, 9)).astype(np.float32)
    hf = abi.HeightFieldCollider(heights, (8.0, 1.5, 6.0), translation=(0.5, 2.0, -0.25))
    max_dist = 0.4
    npts, nfound = 400, 0
    for _ in range(npts):
        pt = np.array([rng.uniform(-4.5, 5.5), rng.uniform(0.2, 3.8), rng.uniform(-3.6, 3.1)], np.float32)
        want_d, want_q, lp = _brute_force(hf, pt)
        got = oracle.collider_project(hf, pt, max_dist)
        if abs(want_d - max_dist) < 1e-4:
            continue  # on the max_dist threshold: either answer is legal in f32
        if want_d > max_dist:
            assert got is None, (pt, want_d, got)
            continue
        assert got is not None, (pt, want_d)
        nfound += 1
        q, inside = got
        assert np.abs(q.astype(np.float64) - want_q).max() <= 2e-5, (pt, q, want_q)
        in_footprint = abs(lp[0]) <= 4.0 and abs(lp[2]) <= 3.0
        if abs(lp[1] - (want_q[1] - hf.translation[1])) > 1e-4 and abs(abs(lp[0]) - 4.0) > 1e-4 and abs(abs(lp[2]) - 3.0) > 1e-4:
            assert inside == (in_footprint and lp[1] <= want_q[1] - hf.translation[1]), (pt, inside)
    assert nfound > 60  # the sample really exercises the surface


def test_heightfield_far_point_and_outside_footprint(oracle):
    hf = scenes.sand3_heightfield_ground()
    assert oracle.collider_project(hf, (0.0, 5.0, 0.0), 0.4) is None          # 5 above the trough bottom
    assert oracle.collider_project(hf, (25.0, 0.0, 0.0), 0.4) is None         # beyond the x extent (+-20)
    q, inside = oracle.collider_project(hf, (1.0, -0.1, 0.0), 0.4)            # just below the trough bottom (y = 0 at z = 0)
    assert inside and abs(q[1]) < 2e-2
    q, inside = oracle.collider_project(hf, (1.0, 0.15, 0.0), 0.4)
    assert not inside and abs(q[1]) < 2e-2


def test_heightfield_scene_runs_on_the_oracle(oracle):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, ground="heightfield", y0=0.45), 0.02, seed=5)
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    for _ in range(2):
        e.step(sc.params, sc.gravity)
    nodes = e.grid_nodes(0)
    tags = np.bincount(nodes["projection_status_tag"].ravel().astype(int), minlength=4)
    assert tags[abi.PROJ_INSIDE if hasattr(abi, "PROJ_INSIDE") else 1] > 0 and tags[2] > 0 and tags[3] > 0
    pos = e.read(abi.ATTR_POSITION)["point"]
    assert np.isfinite(pos).all() and pos[:, 1].min() > 0.15, "the sand must rest on the heightfield, not fall through it"
