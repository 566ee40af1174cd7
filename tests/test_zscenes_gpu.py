"""The two remaining example scenes of the reference, product vs oracle in lock-step: examples3d/sand_penetration3.rs (ROTATED heightfield
colliders) and examples2d/l_panel2.rs (stiff concrete panels, one with a maximum-stress failure model).  The same bodies run on the CUDA
emulation in the CPU suite (tests/test_emulated_cuda_cpu.py)."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes, scenes2d
from tests import test_dim2_gpu as D2
from tests import test_parity_gpu as P3

pytestmark = pytest.mark.gpu


def test_sand_penetration3_against_the_rotated_heightfield_wall(oracle):
    """a reduced lattice placed against the wall at z = 5 (a heightfield turned by -pi / 2 about x: its bumps reach z = 4 at y = 14),
    clear of the floor and of the ceiling: rotated-heightfield projection, inside flag and friction in the product vs the oracle"""
    sc = scenes.jitter(scenes.sand_penetration3(8, 6, 6, origin=(0.0, 13.7, 3.52)), 0.02, seed=12)
    g, o = P3._engines(sc, oracle)
    P3._run_frames(g, o, sc, 8, lambda s: P3.RTOL_1 * (1 + 3 * s))
    nodes = o.grid_nodes(0)
    tags = np.bincount(nodes["projection_status_tag"].ravel().astype(int), minlength=4)
    assert tags[1] > 0 and tags[2] > 0 and tags[3] > 0        # inside / outside / too far all occur
    hit = nodes["projection_status_tag"].ravel() != 3
    assert set(np.unique(nodes["projection_status_payload"].ravel()[hit & (nodes["projection_status_tag"].ravel() != 0)])) <= {2}  # only the z = +5 wall (collider 2) is in reach


def test_l_panel2_lockstep():
    sc = scenes2d.jitter(scenes2d.l_panel2(0.25), 0.0002, seed=3)
    D2._run(sc, 5)
