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


def test_several_contexts_in_one_call(oracle):
    """`CudaMpmPipeline::step(.., contexts: &mut [SingleGpuMpmContext])` with more than one context (the reference's in-process multi-GPU
    mode, cuda_mpm_pipeline.rs:262-659): the particles of a scene are dealt out to three contexts (interleaved: every block is shared),
    one spk_step call advances them all; the union equals the single-context oracle run.  (Contexts on one device here.)"""
    from sparkl_b200.pipeline import CudaMpmPipeline, SingleGpuMpmContext
    from tests import util
    sc = scenes.jitter(scenes.sand3(10, 5, 6), 0.02, seed=5)
    n = len(sc.particles)
    parts = [np.arange(n)[k::3].copy() for k in range(3)]
    ctxs = []
    for ids in parts:
        c = SingleGpuMpmContext(sc.cell_width)
        c.models_upload(sc.models); c.colliders_upload(sc.colliders)
        c.upload(sc.particles.take(ids))
        ctxs.append(c)
    o = oracle.OracleEngine(sc.cell_width)
    o.models_upload(sc.models); o.colliders_upload(sc.colliders); o.upload(sc.particles)
    pipe = CudaMpmPipeline()
    for _ in range(2):
        sim, t = pipe.step(sc.params, sc.gravity, ctxs)
        so, to = o.step(sc.params, sc.gravity)
        assert t.num_substeps == to.num_substeps and abs(sim - so) < 1e-7
    ref = o.read_all()
    pos = np.zeros(n, abi.POSITION); vel = np.zeros(n, abi.VELOCITY); vol = np.zeros(n, abi.VOLUME)
    for ids, c in zip(parts, ctxs):
        pc = c.read_all()
        pos[ids], vel[ids], vol[ids] = pc.position, pc.velocity, pc.volume
    assert util.rel_err(pos["point"], ref.position["point"]) < 2e-5
    assert util.rel_err(vel["vector"], ref.velocity["vector"], floor=1e-2) < 2e-3
    assert util.rel_err(vol["deformation_gradient"], ref.volume["deformation_gradient"], floor=1.0) < 2e-4
    # a context cannot be listed twice
    with pytest.raises(abi.SparklError) as e:
        pipe.step(sc.params, sc.gravity, [ctxs[0], ctxs[0]])
    assert e.value.code == abi.EINVAL


def test_several_contexts_in_one_call_dim2():
    """the same in the dim2 build (96 floats per shared block): two contexts vs the single-context dim2 oracle"""
    from oracle import binding as ob
    from sparkl_b200 import abi2d
    from sparkl_b200.pipeline import CudaMpmPipeline, SingleGpuMpmContext
    from tests import util
    sc = scenes2d.jitter(scenes2d.elasticity2(2, scale=0.3), 0.002, seed=9)
    n = len(sc.particles)
    parts = [np.arange(n)[k::2].copy() for k in range(2)]
    ctxs = []
    for ids in parts:
        c = SingleGpuMpmContext(sc.cell_width, dim=2)
        c.models_upload(sc.models); c.colliders_upload(sc.colliders)
        c.upload(sc.particles.take(ids))
        ctxs.append(c)
    o = ob.OracleEngine2D(sc.cell_width)
    o.models_upload(sc.models); o.colliders_upload(sc.colliders); o.upload(sc.particles)
    pipe = CudaMpmPipeline()
    for _ in range(2):
        sim, t = pipe.step(sc.params, sc.gravity, ctxs)
        so, to = o.step(sc.params, sc.gravity)
        assert t.num_substeps == to.num_substeps
    ref = o.read_all()
    pos = np.zeros(n, abi2d.POSITION); vel = np.zeros(n, abi2d.VELOCITY)
    for ids, c in zip(parts, ctxs):
        pc = c.read_all()
        pos[ids], vel[ids] = pc.position, pc.velocity
    assert util.rel_err(pos["point"], ref.position["point"]) < 2e-5
    assert util.rel_err(vel["vector"], ref.velocity["vector"], floor=1e-2) < 2e-3
