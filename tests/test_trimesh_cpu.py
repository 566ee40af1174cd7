"""TriMesh colliders: the oracle's restatement of parry3d's TriMesh point projection (closest point over all
triangles, surface semantics without pseudo-normals) against an independent f64 brute force, rotation / translation, flip_interior."""
import numpy as np

from oracle import binding as ob
from sparkl_b200 import abi
from tests.test_device_grid3d_cpu import ramp_mesh


def _closest_on_triangle_f64(p, a, b, c):
    """independent route: minimise |a + u (b - a) + v (c - a) - p|^2 over the triangle by checking the interior solution and the three edges"""
    ab, ac = b - a, c - a
    m = np.array([[ab @ ab, ab @ ac], [ab @ ac, ac @ ac]])
    rhs = np.array([(p - a) @ ab, (p - a) @ ac])
    cands = []
    if abs(np.linalg.det(m)) > 1e-14:
        u, v = np.linalg.solve(m, rhs)
        if u >= 0 and v >= 0 and u + v <= 1:
            cands.append(a + u * ab + v * ac)
    for s, e in ((a, b), (b, c), (c, a)):
        d = e - s
        t = np.clip(((p - s) @ d) / (d @ d), 0, 1)
        cands.append(s + t * d)
    return min(cands, key=lambda q: np.linalg.norm(q - p))


def test_trimesh_projection_against_brute_force():
    rng = np.random.default_rng(3)
    verts, tris = ramp_mesh()
    t = 0.4
    quat = (0.0, np.sin(t / 2), 0.0, np.cos(t / 2))  # rotation about y
    R = np.array([[np.cos(t), 0, np.sin(t)], [0, 1, 0], [-np.sin(t), 0, np.cos(t)]])
    tr = np.array([0.3, -0.2, 0.1])
    col = abi.TriMeshCollider(verts, tris, translation=tuple(tr), rotation=quat)
    flipped = abi.TriMeshCollider(verts, tris, translation=tuple(tr), rotation=quat, flip_interior=True)
    V = verts.astype(np.float64)
    for _ in range(150):
        p = rng.uniform([-2, 1, -2], [4, 4, 3])
        proj, inside = ob.collider_project(col, p, 0.4)  # max_dist is ignored for a TriMesh (the reference passes Real::MAX)
        lp = R.T @ (p - tr)
        best = min((_closest_on_triangle_f64(lp, V[i], V[j], V[k]) for i, j, k in tris), key=lambda q: np.linalg.norm(q - lp))
        want = R @ best + tr
        assert abs(np.linalg.norm(proj - p) - np.linalg.norm(want - p)) <= 2e-5
        assert np.linalg.norm(proj - want) <= 2e-3  # the closest point itself is ill-conditioned near edges shared by two faces
        assert inside is False and ob.collider_project(flipped, p, 0.4)[1] is True
    # a point exactly on a vertex of the mesh lies on the surface: "inside" (and not, when flipped)
    on = abi.TriMeshCollider(verts, tris)
    assert ob.collider_project(on, verts[7], 1.0)[1] is True
    assert ob.collider_project(abi.TriMeshCollider(verts, tris, flip_interior=True), verts[7], 1.0)[1] is False


def test_trimesh_upload_validation_in_the_oracle():
    verts, tris = ramp_mesh()
    bad = tris.copy()
    bad[0, 0] = len(verts)  # index out of range
    e = ob.OracleEngine(0.2)
    import pytest
    with pytest.raises(abi.SparklError):
        e.colliders_upload([abi.TriMeshCollider(verts, bad)])
    e.colliders_upload([abi.TriMeshCollider(verts, tris)])
