"""The dim3 DEVICE grid-update kernel (sparkl_b200/csrc/spk_grid.cu: Cuboid / HeightField / Polyline projections, SDF gradient,
Stick / Friction / FrictionZUp boundary handling, projection cache), compiled for the host, against the CPU oracle on GPU-less CI.

Per substep the oracle's grid after P2G is converted to the product's device layout (block-major `acc`, packed projection status,
cache arrays), the product's kernel runs thread by thread (tests/host/grid3d_host.cpp), and its export is compared node by node
with the oracle's grid after `grid_update`."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from sparkl_b200 import abi, scenes
from tests import util
from tests.host.emu2d import DevCollider, GridBuffers, _build

vp = C.c_void_p


@pytest.fixture(scope="module")
def hk3():
    l = _build("grid3d_host")
    assert l.hk3_sizeof_gridbuffers() == C.sizeof(GridBuffers) and l.hk3_sizeof_devcollider() == C.sizeof(DevCollider)
    l.hk3_grid_update.argtypes = [GridBuffers, C.c_uint32, vp, C.c_uint32] + [C.c_float] * 5
    l.hk3_grid_export.argtypes = [GridBuffers, C.c_uint32, vp]
    return l


def _p(a):
    return a.ctypes.data_as(vp)


def _dev_colliders(colliders):
    """mirror of spk_colliders_upload (sparkl_b200/csrc/spk_api.cu), dim3"""
    cols = (DevCollider * max(len(colliders), 1))()
    keep = []
    for i, c in enumerate(colliders):
        d = cols[i]
        if isinstance(c, abi.HeightFieldCollider):
            d.shape, d.nrows, d.ncols, d.hf_flags = 1, c.nrows, c.ncols, 1 | (2 if c.flip_interior else 0)
            data = np.ascontiguousarray(c.heights.T.ravel()).astype(np.float32)  # column-major nrows x ncols
            keep.append(data)
            d.heights = data.ctypes.data
            for a in range(3):
                d.scale[a] = np.float32(c.scale[a])
        elif isinstance(c, abi.TriMeshCollider):
            data = np.concatenate([c.vertices.ravel(), c.indices.ravel().view(np.float32)]).astype(np.float32)  # [vertices | index bits]
            keep.append(data)
            d.shape, d.nrows, d.ncols, d.heights, d.hf_flags = 2, len(c.vertices), len(c.indices), data.ctypes.data, 2 if c.flip_interior else 0
        elif isinstance(c, abi.PolylineCollider):
            data = np.ascontiguousarray(c.vertices, np.float32)
            keep.append(data)
            d.shape, d.nrows, d.heights, d.hf_flags = 3, len(data), data.ctypes.data, 2 if c.flip_interior else 0
        else:
            d.shape = 0
            for a in range(3):
                d.half_extents[a] = np.float32(c.half_extents[a])
        for a in range(4):
            d.rot[a] = np.float32(c.rotation[a])
        for a in range(3):
            d.tr[a] = np.float32(c.translation[a])
        d.friction, d.handling = np.float32(c.friction), c.grid_boundary_handling
    return cols, keep


def _to_device_layout(hdr, nodes):
    nb = len(hdr)
    g = {"key": hdr["virtual_id"].astype(np.uint64).copy(), "acc": np.zeros((max(nb, 1), 6, 64), np.float32), "vel": np.zeros((max(nb, 1) * 64, 4), np.float32),
         "proj": np.zeros(max(nb, 1) * 64, np.uint32), "cacheN": np.zeros((max(nb, 1) * 64, 4), np.float32), "cacheD": np.zeros((max(nb, 1) * 64, 4), np.float32)}
    g["acc"][:nb, 0] = nodes["mass"]
    for a in range(3):
        g["acc"][:nb, 1 + a] = nodes["momentum_velocity"][..., a]
    g["acc"][:nb, 4], g["acc"][:nb, 5] = nodes["psi_momentum_velocity"], nodes["psi_mass"]
    cell = np.arange(64)
    g["acc"][:, :, cell ^ ((cell >> 4) & 2)] = g["acc"][:, :, cell].copy()  # SPK_ACC_SWZ (spk_internal.cuh): position of a cell inside a channel chunk
    tag, payload = nodes["projection_status_tag"].astype(np.uint32), nodes["projection_status_payload"].astype(np.uint32)
    g["proj"][:nb * 64] = (tag | np.where((tag == 1) | (tag == 2), payload << 2, 0)).ravel()
    g["cacheN"][:nb * 64, :3] = nodes["collision_normal"].reshape(-1, 3)
    g["cacheN"][:nb * 64, 3] = nodes["prev_mass"].ravel()
    g["cacheD"][:nb * 64, :3] = nodes["projection_scaled_dir"].reshape(-1, 3)
    blob = GridBuffers()
    blob.cap_blocks = blob.nab_host = nb
    blob.blk_key, blob.vel, blob.acc, blob.proj, blob.cacheN, blob.cacheD = _p(g["key"]), _p(g["vel"]), _p(g["acc"]), _p(g["proj"]), _p(g["cacheN"]), _p(g["cacheD"])
    return g, blob


def _run(hk3, sc, nsub, rtol=2e-5):
    o = ob.OracleEngine(sc.cell_width)
    o.models_upload(sc.models); o.colliders_upload(sc.colliders); o.upload(sc.particles)
    cols, keep = _dev_colliders(sc.colliders)
    p = sc.params
    o.frame_begin(p)
    rem = float(np.float32(p.dt))
    tags = np.zeros(4, int)
    for s in range(nsub):
        o.stage_sort(); o.stage_reset_grid()
        dt = o.stage_estimate_dt(float(np.float32(p.dt) / np.float32(p.max_num_substeps)), float(min(np.float32(rem), np.float32(p.max_substep_dt))))
        o.stage_g2p2g(dt)
        hdr, nodes = o.grid_headers(1), o.grid_nodes(1)
        g, blob = _to_device_layout(hdr, nodes)
        hk3.hk3_grid_update(blob, len(hdr), C.cast(cols, vp), len(sc.colliders), sc.cell_width, dt, *[float(x) for x in sc.gravity])
        blob.has_velocity = 1
        out = np.zeros((len(hdr), 64), abi.GRID_NODE)
        hk3.hk3_grid_export(blob, len(hdr), _p(out))
        o.stage_grid_update(dt, sc.gravity)
        want = o.grid_nodes(0)
        assert np.array_equal(o.grid_headers(0)["virtual_id"], hdr["virtual_id"])
        util.assert_nodes_close(out, want, util.scene_tol(sc, rtol), f"{sc.name} substep {s}: product grid_update (host-compiled) vs oracle", sc, hdr["virtual_id"], ob,
                                velocity_field=True)
        tags += np.bincount(want["projection_status_tag"].ravel().astype(int), minlength=4)[:4]
        rem = float(np.float32(rem) - np.float32(dt))
        if rem <= 1e-11:
            o.frame_begin(p); rem = float(np.float32(p.dt))
    return tags


def test_grid_update_cuboid_floor_and_walls(hk3):
    sc = scenes.jitter(scenes.sand3(10, 5, 8), 0.02, seed=11)
    tags = _run(hk3, sc, 6)
    assert tags[1] > 0 and tags[2] > 0


def test_grid_update_rotated_cuboid_and_stick(hk3):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False), 0.02, seed=12)
    q = np.array([0.0, 0.0, np.sin(0.15), np.cos(0.15)], np.float32)  # rotation about z
    sc.colliders = [abi.CuboidCollider((6.0, 0.4, 6.0), (0.8, 2.05, 0.6), tuple(q), friction=0.3),
                    abi.CuboidCollider((0.3, 3.0, 3.0), (-0.55, 3.0, 0.5), grid_boundary_handling=abi.BoundaryHandling.Stick),
                    abi.CuboidCollider((0.3, 3.0, 3.0), (2.35, 3.0, 0.5), grid_boundary_handling=abi.BoundaryHandling.FrictionZUp),
                    abi.CuboidCollider((9.0, 9.0, 9.0), (0.0, 0.0, 0.0), grid_boundary_handling=abi.BoundaryHandling.NONE)]
    _run(hk3, sc, 6)


def test_grid_update_heightfield_ground(hk3):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, ground="heightfield", y0=0.9), 0.02, seed=13)
    tags = _run(hk3, sc, 5)
    assert tags[1] + tags[2] > 0 and tags[3] > 0  # heightfield projections are limited to 2 h: TooFar occurs


def test_grid_update_polyline_trough(hk3):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=0.5), 0.02, seed=6)
    trough = np.array([[-1.03, 0.31], [-1.03, 6.1], [-1.52, 6.1], [-1.52, -0.53], [3.51, -0.53], [3.51, 6.1], [3.07, 6.1], [3.07, 0.31]], np.float32)
    sc.colliders = [abi.PolylineCollider(trough)]
    _run(hk3, sc, 5)


def ramp_mesh(nx=6, nz=5, x0=-1.0, x1=3.0, z0=-1.0, z1=2.0, y_lo=2.05, y_hi=2.9):
    """a tilted, slightly bumpy surface triangulated on an nx x nz lattice (vertices off the grid nodes)"""
    xs, zs = np.linspace(x0, x1, nx), np.linspace(z0, z1, nz)
    verts = np.array([[x, y_lo + (y_hi - y_lo) * (x - x0) / (x1 - x0) + 0.07 * np.sin(3.1 * x + 1.7 * z), z] for x in xs for z in zs], np.float32)
    tris = []
    for i in range(nx - 1):
        for j in range(nz - 1):
            a, b, c, d = i * nz + j, (i + 1) * nz + j, (i + 1) * nz + j + 1, i * nz + j + 1
            tris += [[a, b, c], [a, c, d]]
    return verts, np.array(tris, np.uint32)


def test_grid_update_trimesh_ramp(hk3):
    sc = scenes.jitter(scenes.sand3(8, 4, 6, with_colliders=False, y0=2.4), 0.02, seed=15)
    verts, tris = ramp_mesh()
    sc.colliders = [abi.TriMeshCollider(verts, tris, friction=0.4), abi.CuboidCollider((9.0, 0.5, 9.0), (0.0, 1.0, 0.0))]
    tags = _run(hk3, sc, 6)
    assert tags[2] > 0
