"""The custom particle-kernel extension point (`ParticleUpdater`, `g2p2g_generic`, `CudaParticleKernelsLauncher`,
`GpuConstitutiveModel::Custom(u32)`) on the GPU.  examples/custom_updater instantiates the library's own G2P2G / dt kernel
templates with a user-defined updater and registers its launch functions through spk_set_particle_kernels_launcher."""
import ctypes as C
import os

import numpy as np
import pytest

from sparkl_b200 import abi, scenes
from tests import util

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLE = os.path.join(ROOT, "examples", "custom_updater", "libcustom_updater.so")


@pytest.fixture(scope="module")
def example():
    return _load_example()


def _load_example():
    if not os.path.exists(EXAMPLE):
        from sparkl_b200 import build
        build.build()
    l = C.CDLL(EXAMPLE)
    l.cu_launcher_create.argtypes = [C.c_void_p, C.c_uint32, C.c_float, C.POINTER(C.c_void_p)]
    l.cu_launcher_vtable.argtypes = [C.c_void_p]
    l.cu_launcher_vtable.restype = C.c_void_p
    l.cu_launcher_calls.argtypes = [C.c_void_p, C.c_int]
    l.cu_launcher_destroy.argtypes = [C.c_void_p]
    from tests.host.emu2d import DEVMODEL
    assert l.cu_sizeof_devmodel() == DEVMODEL.itemsize
    return l


def _launcher(example, models, damping=0.0):
    from tests.host.emu2d import devmodel
    table = np.concatenate([devmodel(m) for m in models])
    h = C.c_void_p()
    assert example.cu_launcher_create(table.ctypes.data_as(C.c_void_p), len(table), damping, C.byref(h)) == 0
    return h


def _scene(custom: bool):
    sc = scenes.jitter(scenes.neo_hookean_block(12, 10, 8), 0.002, seed=4)
    rng = np.random.Generator(np.random.PCG64(5))  # seeded pre-strain: the constitutive model acts from the first substep on
    n = len(sc.particles)
    F = np.tile(np.eye(3, dtype=np.float32), (n, 1, 1)) + rng.uniform(-0.05, 0.05, (n, 3, 3)).astype(np.float32) * np.float32(0.5)
    sc.particles.volume["deformation_gradient"] = F.transpose(0, 2, 1).reshape(n, 9)
    real = sc.models[0]
    if custom:
        sc.models = [abi.ParticleModel(abi.CustomConstitutiveModel(0))]
    return sc, real


def test_custom_updater_reproduces_the_builtin_model(example):
    """Custom(0) resolved by the user's updater to a Neo-Hookean table entry == the built-in Neo-Hookean model: same substep
    count, state equal up to the summation order of the P2G atomics"""
    from sparkl_b200.pipeline import context_from_scene
    sc_ref, real = _scene(False)
    sc_cus, _ = _scene(True)
    g_ref, g_cus = context_from_scene(sc_ref), context_from_scene(sc_cus)
    h = _launcher(example, [real])
    assert g_cus._lib.spk_set_particle_kernels_launcher(g_cus._h, example.cu_launcher_vtable(h)) == 0
    nsub = 0
    for _ in range(2):
        (s1, t1), (s2, t2) = g_ref.step(sc_ref.params, sc_ref.gravity), g_cus.step(sc_cus.params, sc_cus.gravity)
        assert t1.num_substeps == t2.num_substeps and abs(s1 - s2) <= 1e-7
        nsub += t2.num_substeps
    a, b = g_cus.read_all(), g_ref.read_all()
    a.status["model_index"] = b.status["model_index"]
    util.assert_particles_close(a, b, sc_ref.cell_width, util.scene_tol(sc_ref, 2e-5 * nsub), "custom updater vs built-in model")
    assert np.abs(b.velocity["vector"] - b.velocity["vector"].mean(axis=0)).max() > 1e-2  # the stress did act (not a rigid free fall)
    assert example.cu_launcher_calls(h, 0) == 2          # timestep_started: once per spk_step
    assert example.cu_launcher_calls(h, 2) == nsub       # launch_g2p2g: once per substep
    assert example.cu_launcher_calls(h, 1) == nsub       # estimate: a REGISTERED estimate launcher is called every substep (step_generic, cuda_mpm_pipeline.rs:381-392)
    g_cus.close()
    example.cu_launcher_destroy(h)


def test_default_launcher_treats_custom_as_zero_stress_and_null_restores_it(example):
    """without a launcher `Custom(id)` is the reference's default: zero stress, no dt bound (gpu_constitutive_model.rs:83,108)"""
    from sparkl_b200.pipeline import context_from_scene
    sc, real = _scene(True)
    sc.colliders = []
    g = context_from_scene(sc)
    h = _launcher(example, [real])
    vt = example.cu_launcher_vtable(h)
    assert g._lib.spk_set_particle_kernels_launcher(g._h, vt) == 0
    assert g._lib.spk_set_particle_kernels_launcher(g._h, None) == 0
    params = abi.SolverParameters()  # the scene's own parameters cap the substep; here nothing but the frame length should
    for _ in range(3):
        _, t = g.step(params, sc.gravity)
        assert t.num_substeps <= 2  # no constitutive dt bound: only the velocity bound / frame length limit the substep
    assert example.cu_launcher_calls(h, 2) == 0 and example.cu_launcher_calls(h, 0) == 0
    F = g.read(abi.ATTR_VOLUME)["deformation_gradient"]
    v = g.read(abi.ATTR_VELOCITY)["vector"]
    assert np.abs(v[:, 0]).max() <= 1e-4 and np.abs(v[:, 1] - v[:, 1].mean()).max() <= 1e-3  # rigid free fall: no stress acts
    assert np.isfinite(F).all()
    example.cu_launcher_destroy(h)


def test_custom_updater_damping_and_stage_calls(example):
    """the updater's own physics (velocity damping) acts, and the stage-level entry points go through the launcher too"""
    from sparkl_b200.pipeline import context_from_scene
    sc, real = _scene(True)
    sc.colliders = []
    runs = {}
    for damping in (0.0, 20.0):
        g = context_from_scene(sc)
        h = _launcher(example, [real], damping)
        assert g._lib.spk_set_particle_kernels_launcher(g._h, example.cu_launcher_vtable(h)) == 0
        g.frame_begin(sc.params)
        rem = float(np.float32(sc.params.dt))
        for _ in range(5):
            rem -= g.substep(sc.params, sc.gravity, rem)
        assert example.cu_launcher_calls(h, 2) == 5
        runs[damping] = g.read(abi.ATTR_VELOCITY)["vector"].copy()
        g.close()
        example.cu_launcher_destroy(h)
    vy0, vy1 = runs[0.0][:, 1].mean(), runs[20.0][:, 1].mean()
    assert vy0 < 0 and vy0 < vy1 < 0 and abs(vy1) < 0.99 * abs(vy0)
