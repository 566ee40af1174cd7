"""Fluids-volume recomputation: `spk_stage_recompute_fluids_volume` against the oracle's restatement of the reference's
pass (src/dynamics/solver/fluids_volume.rs:9-116; GPU kernels src_kernels/cuda/fluids_volume.rs:9-105) and a size-independent
property, in dim3 and dim2.  The same bodies run on the CUDA emulation in the CPU suite (tests/test_emulated_cuda_cpu.py)."""
import numpy as np
import pytest

from sparkl_b200 import abi, scenes, scenes2d
from tests import test_dim2_gpu as D2
from tests import test_parity_gpu as P3
from tests import util

pytestmark = pytest.mark.gpu


def _fluid_next_to_solid():
    """a Monaghan fluid block touching a corotated block: the solid's particles scatter mass too but keep their F"""
    from sparkl_b200.abi import CorotatedLinearElasticity, MonaghanSphEos, ParticleComponents, ParticleModel, cube_particles
    h, r = 0.8, 0.1
    models = [ParticleModel(MonaghanSphEos(1.0e6, 7, 1.01e-3)), ParticleModel(CorotatedLinearElasticity(1.0e6, 0.3))]
    pc = ParticleComponents.concatenate([cube_particles((1.6, 1.6, 1.6), 14, 12, 10, 0, r, 1000.0),
                                         cube_particles((1.6 + 14 * 0.2, 1.6, 1.6), 6, 12, 10, 1, r, 2000.0)])
    return scenes.jitter(scenes.Scene("fluid_next_to_solid", h, pc, models, []), 0.03, seed=8)


def test_fluids_volume_recomputation(oracle):
    """the fluids-volume recomputation (src_kernels/cuda/fluids_volume.rs, sequenced like
    src/dynamics/solver/fluids_volume.rs:9-116) as the explicit stage between reset_grid and g2p2g, against the oracle"""
    sc = _fluid_next_to_solid()
    g, o = P3._engines(sc, oracle)
    with pytest.raises(abi.SparklError) as e:
        g.stage_recompute_fluids_volume()  # needs the sort's particle -> block map
    assert e.value.code == abi.ESTATE
    P3._run_frames(g, o, sc, 2, lambda s: 5e-5 * (1 + 3 * s))
    p = sc.params
    for rep in range(2):  # twice: the pass leaves the grid and the sort state as it found them
        g.stage_sort(); o.stage_sort()
        g.stage_reset_grid(); o.stage_reset_grid()
        before = g.read(abi.ATTR_VOLUME).copy()
        g.stage_recompute_fluids_volume(); o.stage_recompute_fluids_volume()
        vg, vo = g.read(abi.ATTR_VOLUME), o.read(abi.ATTR_VOLUME)
        fluid = g.read(abi.ATTR_STATUS)["model_index"] == 0
        Fg, Fo = vg["deformation_gradient"], vo["deformation_gradient"]
        assert np.allclose(Fg[fluid, 0], Fo[fluid, 0], rtol=2e-5, atol=0), np.abs(Fg[fluid, 0] / Fo[fluid, 0] - 1).max()
        assert np.array_equal(Fg[~fluid], before["deformation_gradient"][~fluid])  # only the fluids change (fluids_volume.rs:75-77)
        assert np.array_equal(Fg[:, 1:], before["deformation_gradient"][:, 1:])    # and only F[0][0]
        if rep == 0:
            assert (np.abs(Fg[fluid, 0] - before["deformation_gradient"][fluid, 0]) > 1e-2).any()  # the pass did something
        for name in ("mass", "volume0", "radius0", "plastic_deformation_gradient_det"):
            assert np.array_equal(vg[name], before[name]), name
        assert not g.grid_nodes(1)["mass"].any()  # "Reset the modified cells", fluids_volume.rs:98-113
        # the substep goes on in lock-step from the recomputed volumes
        min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
        g.frame_begin(p); o.frame_begin(p)
        g.stage_estimate_dt(min_dt, float(p.max_substep_dt))
        dt = o.stage_estimate_dt(min_dt, float(p.max_substep_dt))
        g.set_dt_acc(o.dt_acc())
        g.stage_g2p2g(dt); o.stage_g2p2g(dt)
        tol = util.scene_tol(sc, 2e-4)
        util.assert_particles_close(g.read_all(), o.read_all(), sc.cell_width, tol, f"after g2p2g following the recomputation {rep}")
        g.stage_grid_update(dt, sc.gravity); o.stage_grid_update(dt, sc.gravity)


def test_fluids_volume_of_a_rest_lattice_is_its_rest_volume():
    """property: in the interior of a lattice at rest spacing the grid density equals the rest density (partition of unity of the
    quadratic B-splines), so the recomputed volume is volume0 (F[0][0] = 1); at the free surface the volume grows"""
    from sparkl_b200.pipeline import context_from_scene
    sc = scenes.fluids3(32, 32, 32, with_colliders=False)
    g = context_from_scene(sc)
    g.stage_sort(); g.stage_reset_grid()
    g.stage_recompute_fluids_volume()
    F00 = g.read(abi.ATTR_VOLUME)["deformation_gradient"][:, 0].reshape(32, 32, 32)
    assert np.abs(F00[13:-13, 13:-13, 13:-13] - 1.0).max() < 1e-4   # more than two stencil radii (2 x 1.5 h = 12 particles) away from the surface
    assert F00[0, 0, 0] > 4.0 and (F00 > 1.0 - 1e-4).all()


def test_fluids_volume_recomputation_dim2():
    """fluids-volume recomputation (src_kernels/cuda/fluids_volume.rs with the `dim2` feature) against the dim2 oracle,
    and the rest-lattice property (interior F[0][0] = 1)"""
    sc = scenes2d.jitter(scenes2d.fluids2(40, 30, with_colliders=False), 0.004, seed=6)
    g, o = D2._engines(sc)
    with pytest.raises(abi.SparklError) as e:
        g.stage_recompute_fluids_volume()
    assert e.value.code == abi.ESTATE
    g.stage_sort(); o.stage_sort()
    g.stage_reset_grid(); o.stage_reset_grid()
    before = g.read(abi.ATTR_VOLUME).copy()
    g.stage_recompute_fluids_volume(); o.stage_recompute_fluids_volume()
    Fg, Fo = g.read(abi.ATTR_VOLUME)["deformation_gradient"], o.read(abi.ATTR_VOLUME)["deformation_gradient"]
    assert np.allclose(Fg[:, 0], Fo[:, 0], rtol=2e-5, atol=0), np.abs(Fg[:, 0] / Fo[:, 0] - 1).max()
    assert np.array_equal(Fg[:, 1:], before["deformation_gradient"][:, 1:])
    assert not g.grid_nodes(1)["mass"].any()
    rest = scenes2d.fluids2(32, 32, with_colliders=False)
    g2, _ = D2._engines(rest)
    g2.stage_sort(); g2.stage_reset_grid(); g2.stage_recompute_fluids_volume()
    F00 = g2.read(abi.ATTR_VOLUME)["deformation_gradient"][:, 0].reshape(32, 32)
    assert np.abs(F00[8:-8, 8:-8] - 1.0).max() < 1e-4 and F00[0, 0] > 2.0
