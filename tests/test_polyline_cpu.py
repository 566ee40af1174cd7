"""Polyline colliders: the oracle's restatement of `polyline_project_point`
(src_kernels/gpu_collider.rs:105-201, in-tree) against an independent float64 check: closest point on the closed polygon,
inside = point-in-polygon (counter-clockwise interior), z of the query preserved."""
import numpy as np

from sparkl_b200 import abi, scenes


def _closest_on_polygon(poly, p):
    best, bestq = np.inf, None
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        ab = b - a
        t = np.clip(np.dot(p - a, ab) / np.dot(ab, ab), 0.0, 1.0)
        q = a + t * ab
        d = np.linalg.norm(p - q)
        if d < best:
            best, bestq = d, q
    return best, bestq


def _point_in_polygon(poly, p):
    inside = False
    n = len(poly)
    for i in range(n):
        a, b = poly[i], poly[(i + 1) % n]
        if (a[1] > p[1]) != (b[1] > p[1]) and p[0] < (b[0] - a[0]) * (p[1] - a[1]) / (b[1] - a[1]) + a[0]:
            inside = not inside
    return inside


def test_polyline_projection_and_inside_test(oracle):
    # a counter-clockwise L-shaped polygon (one re-entrant vertex) in the z = 0 plane, untransformed (so that the reference's
    # world-minus-local quirk in the inside test is neutral)
    poly = np.array([[0, 0], [4, 0], [4, 1.5], [1.5, 1.5], [1.5, 3], [0, 3]], np.float64)
    col = abi.PolylineCollider(poly.astype(np.float32))
    rng = np.random.Generator(np.random.PCG64(17))
    checked = 0
    for _ in range(600):
        p = np.array([rng.uniform(-1, 5), rng.uniform(-1, 4), 0.0])
        d, q = _closest_on_polygon(poly, p[:2])
        if d < 1e-3:
            continue  # on the boundary: the side is decided by round-off
        got = oracle.collider_project(col, p.astype(np.float32), 0.4)
        assert got is not None
        pt, inside = got
        assert np.abs(pt[:2].astype(np.float64) - q).max() <= 1e-5 and pt[2] == np.float32(p[2])
        assert inside == _point_in_polygon(poly, p[:2]), (p, inside)
        checked += 1
    assert checked > 500
    flipped = abi.PolylineCollider(poly.astype(np.float32), flip_interior=True)
    pt, inside = oracle.collider_project(flipped, np.array([1.0, 1.0, 0.0], np.float32), 0.4)
    assert not inside
    # the query's z is kept and the polygon acts as a prism along z
    pt, inside = oracle.collider_project(col, np.array([1.0, 1.0, 7.5], np.float32), 0.4)
    assert pt[2] == np.float32(7.5)


def test_polyline_scene_runs_on_the_oracle(oracle):
    """sand in a polygonal trough: the polyline is the outline of the U-shaped container wall"""
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=0.5), 0.02, seed=6)
    trough = np.array([[-1, 0.3], [-1, 6], [-1.5, 6], [-1.5, -0.5], [3.5, -0.5], [3.5, 6], [3, 6], [3, 0.3]], np.float32)  # counter-clockwise: the U-shaped wall is the solid
    sc.colliders = [abi.PolylineCollider(trough)]
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    for _ in range(2):
        e.step(sc.params, sc.gravity)
    pos = e.read(abi.ATTR_POSITION)["point"]
    tags = np.bincount(e.grid_nodes(0)["projection_status_tag"].ravel().astype(int), minlength=4)
    assert np.isfinite(pos).all() and pos[:, 1].min() > 0.3 and tags[1] > 0 and tags[2] > 0
