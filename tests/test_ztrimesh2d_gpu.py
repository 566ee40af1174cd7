"""dim2 TriMesh colliders (parry2d's TriMesh: solid 2D triangles, SURVEY §8 f-1 x f-3): product vs the dim2 oracle.  The same body runs
on the CUDA emulation in the CPU suite (tests/test_emulated_cuda_cpu.py); the oracle itself is checked against an f64 brute force in
tests/test_dim2_cpu.py."""
import numpy as np
import pytest

from sparkl_b200 import abi, abi2d, scenes2d
from tests import test_dim2_gpu as D2

pytestmark = pytest.mark.gpu


def _ramp_scene(flip=False):
    """sand dropped on a ramp that is ONE large triangle (a node deep inside a single triangle has an unambiguous projection; at an edge
    shared by two triangles the winner, and with it the inside flag, depends on round-off in the reference too) + a detached second
    triangle so that the search runs over more than one element; no vertex or edge passes through a grid node (multiples of h = 0.2)"""
    sc = scenes2d.model_block2("sand", 12, 10)
    v = np.array([[-3.03, 2.96], [1.01, -1.34], [5.07, 1.88], [8.3, 0.1], [9.4, 0.2], [8.8, 1.3]], np.float32)
    t = np.array([[0, 1, 2], [3, 4, 5]], np.uint32)
    sc.colliders = [abi2d.TriMeshCollider(v, t, flip_interior=flip)]
    sc.particles.velocity["vector"][:, 1] = -2.0
    return sc


def test_sand_on_a_trimesh_ramp_dim2():
    g, o = D2._run(_ramp_scene(), 8)
    tags = np.bincount(o.grid_nodes(0)["projection_status_tag"].ravel().astype(int), minlength=4)
    assert tags[1] > 0 and tags[2] > 0  # nodes inside the ramp and nodes above it both got a projection


def test_trimesh_argument_errors_dim2():
    from sparkl_b200.pipeline import SingleGpuMpmContext
    g = SingleGpuMpmContext(0.2, dim=2)
    v = np.array([[0, 0], [1, 0], [0, 1]], np.float32)
    with pytest.raises(abi.SparklError) as e:
        g.colliders_upload([abi2d.TriMeshCollider(v, np.array([[0, 1, 3]], np.uint32))])  # index past the vertex list
    assert e.value.code == abi.EINVAL
    g.colliders_upload([abi2d.TriMeshCollider(v, np.array([[0, 1, 2]], np.uint32))])
