"""The product's CUDA sources on a GPU-less machine.

tests/host/emu/ compiles sparkl_b200/csrc/*.cu — every kernel AND the host sequencing of spk_api.cu — for the host with g++ on top
of a small CUDA execution model (each CUDA thread is a fiber; __syncthreads / warp collectives are scheduling points; TMA bulk
copies become loops; see tests/host/emu/cuda_emu.h).  The resulting libsparkl_b200_emu*.so export the same C ABI, so the bodies of
the GPU test modules run against them unchanged: same scenes, same comparisons with the oracle and with the committed outputs of
the reference's own kernels, same tolerances.  What the emulation cannot show: anything about performance, the asynchronous
overlap of streams, real TMA / mbarrier behaviour and the NVLink path — those stay with `pytest -m gpu`.

The emulated libraries are test infrastructure: they live in tests/host/, nothing under sparkl_b200/ knows them, and the product
still fails loudly without a GPU (tests/test_abi_cpu.py::test_no_cpu_fallback runs with the real library)."""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests", "host", "emu"))


@pytest.fixture()
def emulated(monkeypatch):
    """route sparkl_b200._lib.lib(dim) to the emulated libraries for the duration of one test"""
    import build_emu
    from sparkl_b200 import _lib
    kw = {}
    if os.environ.get("SPK_EMU_ASAN") == "1":  # AddressSanitizer build: run pytest with LD_PRELOAD=$(gcc -print-file-name=libasan.so)
        kw = dict(tag="_asan", extra_flags=["-fsanitize=address", "-fno-omit-frame-pointer", "-g"])
    monkeypatch.setattr(_lib, "LIB_PATH", build_emu.build(3, **kw))
    monkeypatch.setattr(_lib, "LIB2D_PATH", build_emu.build(2, **kw))
    monkeypatch.setattr(_lib, "_LIBS", {})
    monkeypatch.setattr(_lib, "_LIB", None)
    yield _lib
    # monkeypatch restores the real paths and an empty cache: the next lib() call loads the real library again


def test_the_emulation_is_not_a_product_path():
    from sparkl_b200 import _lib
    for attr in ("LIB_PATH", "LIB2D_PATH"):
        assert "emu" not in os.path.basename(getattr(_lib, attr))
    for root, _, files in os.walk(os.path.join(ROOT, "sparkl_b200")):
        for f in files:
            if f.endswith(".py"):
                assert "libsparkl_b200_emu" not in open(os.path.join(root, f)).read(), f


# ---------------------------------------------------------------------------------------------- dim3: product vs oracle, stage by stage
@pytest.mark.parametrize("name", ["test_sand_first_substeps", "test_fluid_dam_break", "test_neo_hookean_block", "test_sand_on_heightfield_ground",
                                  "test_sand_in_a_polyline_trough", "test_sand_on_a_trimesh_ramp", "test_cube_through_sand_kinematic",
                                  "test_full_step_matches_oracle", "test_empty_and_single_particle", "test_ragged_blocks_and_chunk_boundaries",
                                  "test_append_remove_read_roundtrip"])
def test_dim3_parity_bodies_on_the_emulation(emulated, oracle, name):
    import tests.test_parity_gpu as T
    getattr(T, name)(oracle)


@pytest.mark.parametrize("kind", ["nacc", "rankine", "snow", "failure", "failure_split"])
def test_dim3_other_models_on_the_emulation(emulated, oracle, kind):
    import tests.test_parity_gpu as T
    T.test_other_plasticity_and_failure_models(oracle, kind)


@pytest.mark.parametrize("name", ["test_async_readback_matches_blocking_reads", "test_invalid_collider_shapes_are_rejected"])  # the async-upload test pins host memory through torch, which needs a driver
def test_dim3_api_bodies_on_the_emulation(emulated, name):
    import tests.test_parity_gpu as T
    getattr(T, name)()


@pytest.mark.parametrize("name", ["test_negative_far_and_nan_coordinates", "test_failed_and_static_particles", "test_model_table_with_unused_entries",
                                  "test_block_ids_follow_the_morton_curve"])
# (test_hash_map_grows_when_blocks_outnumber_the_estimate also passes on the emulation but takes a minute there: GPU suite only)
def test_dim3_edge_cases_on_the_emulation(emulated, oracle, name):
    import tests.test_edge_gpu as T
    getattr(T, name)(oracle)


def test_fluids_volume_recomputation_on_the_emulation(emulated, oracle):
    import tests.test_volume_recomputation_gpu as T
    T.test_fluids_volume_recomputation(oracle)
    T.test_fluids_volume_of_a_rest_lattice_is_its_rest_volume()
    T.test_fluids_volume_recomputation_dim2()


def test_device_side_migration_on_the_emulation(emulated):
    import tests.test_zmigration_gpu as T
    for k in (0, 3):
        T.test_migrate_out_is_a_stable_partition_and_in_appends(k)
    T.test_simulation_goes_on_from_a_migrated_set()
    T.test_migrate_in_grows_on_the_device()
    T.test_truncate_and_capacity()
    T.test_migration_argument_errors()


def _emulated_migration_worker(rank, world, port, nframes, out_dir):
    """one rank of a 2-process run of the PRODUCT (emulated library) over gloo: host-staged halo exchange through the stage calls
    (DistributedMpmPipeline) + device-side migration (migrate_device) after every frame"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import build_emu
    from sparkl_b200 import _lib
    _lib.LIB_PATH, _lib.LIB2D_PATH, _lib._LIBS, _lib._LIB = build_emu.build(3), build_emu.build(2), {}, None
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparkl_b200 import abi
        from sparkl_b200.distributed import DistributedMpmPipeline, SlabPartition, migrate_device
        from sparkl_b200.pipeline import SingleGpuMpmContext
        from tests.test_distributed_cpu import _migration_scene
        sc = _migration_scene()
        part = SlabPartition([-(1 << 40), -1, 1 << 40])
        owner = part.owner(sc.particles.position["point"][:, 0], sc.cell_width)
        gids = np.nonzero(owner == rank)[0].astype(np.uint32)
        e = SingleGpuMpmContext(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders)
        e.upload(sc.particles.take(gids))
        e.set_gids(gids)
        cpu = torch.device("cpu")
        pipe = DistributedMpmPipeline(e, dist, cpu)
        moved = 0
        for _ in range(nframes):
            pipe.step(sc.params, sc.gravity)
            sent, recv = migrate_device(e, part, dist, cpu)
            moved += sent
            assert (part.owner(e.read(abi.ATTR_POSITION)["point"][:, 0], sc.cell_width) == rank).all()
        pc = e.read_all()
        np.savez(os.path.join(out_dir, f"mig{rank}.npz"), ids=e.read_gids(), pos=pc.position, vel=pc.velocity, vol=pc.volume, moved=moved)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_with_device_side_migration_on_the_emulation(tmp_path, oracle):
    """world_size-2 gloo run of the product's own code (emulated): the particles that cross the slab boundary are re-homed by
    spk_particles_migrate_out / _in, and the union of the two ranks equals the single-rank oracle run"""
    import numpy as np
    import torch.multiprocessing as mp
    import build_emu
    from sparkl_b200 import abi
    from tests import util
    from tests.test_distributed_cpu import _free_port, _migration_scene
    build_emu.build(3); build_emu.build(2)  # before the workers race to build them
    nframes = 4
    mp.spawn(_emulated_migration_worker, args=(2, _free_port(), nframes, str(tmp_path)), nprocs=2, join=True)
    sc = _migration_scene()
    e = oracle.OracleEngine(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    for _ in range(nframes):
        e.step(sc.params, sc.gravity)
    ref = e.read_all()
    r = [np.load(tmp_path / f"mig{k}.npz") for k in range(2)]
    assert int(r[0]["moved"]) + int(r[1]["moved"]) > 0
    ids = np.concatenate([r[0]["ids"], r[1]["ids"]])
    assert np.array_equal(np.sort(ids), np.arange(len(ref)))
    pos = np.zeros(len(ref), abi.POSITION); vel = np.zeros(len(ref), abi.VELOCITY); vol = np.zeros(len(ref), abi.VOLUME)
    for k in range(2):
        pos[r[k]["ids"]], vel[r[k]["ids"]], vol[r[k]["ids"]] = r[k]["pos"], r[k]["vel"], r[k]["vol"]
    assert util.rel_err(pos["point"], ref.position["point"]) < 2e-5
    assert util.rel_err(vel["vector"], ref.velocity["vector"], floor=1e-2) < 2e-3
    assert util.rel_err(vol["deformation_gradient"], ref.volume["deformation_gradient"], floor=1.0) < 2e-4


# ---------------------------------------------------------------------------------------------- product vs the reference's own kernels (fixtures)
def _golden_cases():
    from tests import util
    return util.golden_cases()


@pytest.mark.parametrize("case", _golden_cases())
def test_product_matches_reference_fixture_on_the_emulation(emulated, case):
    import tests.test_golden_gpu as T
    T.test_product_matches_reference_fixture(case)


# ---------------------------------------------------------------------------------------------- dim2
@pytest.mark.parametrize("name", ["test_upload_read_roundtrip_dim2", "test_elasticity2_lockstep", "test_elasticity2_on_the_ground", "test_fluids2_dam_break_lockstep",
                                  "test_polyline_trough_dim2", "test_basic2_on_the_heightfield_ground", "test_full_step_matches_oracle_dim2",
                                  "test_momentum_conservation_and_free_fall_dim2", "test_unsupported_dim2_features_fail_loudly"])
def test_dim2_bodies_on_the_emulation(emulated, name):
    import tests.test_dim2_gpu as T
    getattr(T, name)()


@pytest.mark.parametrize("kind", ["corotated", "neo_hookean", "fluid", "sand", "nacc", "failure", "failure_split"])
def test_dim2_models_on_the_emulation(emulated, kind):
    import tests.test_dim2_gpu as T
    T.test_model_block2_lockstep(kind)


def test_streaming_upload_and_async_readback_on_the_emulation(emulated):
    """the e2e loop of bench.py (stage the next state, commit, step, asynchronous read-back) through the same entry points, with plain
    numpy arrays standing in for pinned memory: same state as the blocking calls"""
    import numpy as np
    from sparkl_b200 import abi, scenes
    from sparkl_b200.pipeline import context_from_scene
    sc = scenes.jitter(scenes.sand3(8, 4, 6), 0.02, seed=5)
    a, b = context_from_scene(sc), context_from_scene(sc)
    pc = sc.particles
    for frame in range(2):
        a.upload(pc)
        b.upload_async(pc, stage_only=True)
        b.commit_staged()
        (sa, ta), (sb, tb) = a.step(sc.params, sc.gravity), b.step(sc.params, sc.gravity, timings=False)
        assert ta.num_substeps > 0 and abs(sa - sb) < 1e-7
        out = {w: np.empty(len(pc), abi.ATTR_DTYPES[w]) for w in (abi.ATTR_POSITION, abi.ATTR_VELOCITY, abi.ATTR_PHASE)}
        for w, buf in out.items():
            b.read_async(w, buf)
        b.read_wait()
        for w, buf in out.items():
            assert np.array_equal(buf, a.read(w)), w  # the emulation is deterministic: bit-identical
        pc = a.read_all()


def test_g2p2g_mode0_on_the_emulation(emulated, oracle, monkeypatch):
    """the A/B variant of the fused kernel (SPK_G2P2G_MODE=0: per-particle shared atomics instead of the cell-binned register
    accumulation) gives the same physics"""
    import tests.test_parity_gpu as T
    monkeypatch.setenv("SPK_G2P2G_MODE", "0")
    T.test_sand_first_substeps(oracle)
    T.test_cube_through_sand_kinematic(oracle)


def test_cpp_host_example_builds_and_matches_the_python_host(emulated, tmp_path):
    """include/sparkl_b200.hpp (the C++ mirror of the reference's host API) + examples/cpp_host/sand_column.cpp: linked against the
    emulated library it must reproduce, bit for bit, what the Python mirror computes through the same library; linked against the
    real library on a GPU-less machine it must fail loudly (no CPU fallback)"""
    import subprocess
    import numpy as np
    import build_emu
    from sparkl_b200 import abi, scenes
    from sparkl_b200.pipeline import context_from_scene
    src = os.path.join(ROOT, "examples", "cpp_host", "sand_column.cpp")
    inc = os.path.join(ROOT, "include")
    emu_lib = build_emu.build(3)
    exe = str(tmp_path / "sand_column_emu")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + inc, src, "-o", exe, emu_lib, "-Wl,-rpath," + os.path.dirname(emu_lib)])
    out_bin = str(tmp_path / "pos.bin")
    out = subprocess.check_output([exe, "8", "4", "6", "2", out_bin]).decode()
    assert "particles 384" in out and "ERROR" not in out, out
    pos_cpp = np.fromfile(out_bin, np.float32).reshape(-1, 3)
    # the same scene through the Python mirror (scenes.sand3 builds exactly this lattice) and the same library
    sc = scenes.sand3(8, 4, 6, with_colliders=False)
    ext = 0.1
    sc.colliders = [abi.CuboidCollider((8 * ext / 2 + 6.0, 1.0, 6 * ext / 2 + 6.0), (8 * ext / 2, 1.0, 6 * ext / 2))]
    g = context_from_scene(sc)
    for _ in range(2):
        g.step(sc.params, sc.gravity)
    assert np.array_equal(g.read(abi.ATTR_POSITION)["point"], pos_cpp)
    # the heightfield builder of the C++ mirror against the Python one
    out = subprocess.check_output([exe, "8", "4", "6", "2", out_bin, "heightfield"]).decode()
    assert "particles 384" in out, out
    pos_cpp = np.fromfile(out_bin, np.float32).reshape(-1, 3)
    col, row = np.meshgrid(np.arange(41), np.arange(41), indexing="xy")  # heights[row, col]
    heights = (((col * 3 + row * 5) % 7).astype(np.float32) * np.float32(0.001)).astype(np.float32)
    y0 = np.float32(0.2) * np.float32(3.0) + np.float32(2.0)
    sc.colliders = [abi.HeightFieldCollider(heights, (40.0, 10.0, 40.0), (8 * ext / 2, float(y0 - np.float32(0.3)), 6 * ext / 2))]
    g = context_from_scene(sc)
    for _ in range(2):
        g.step(sc.params, sc.gravity)
    assert np.array_equal(g.read(abi.ATTR_POSITION)["point"], pos_cpp)
    assert (g.grid_nodes(0)["projection_status_tag"] == 1).any()  # some nodes are inside the ground: the collider acts
    # against the real library: builds, and without a GPU exits with the "no usable CUDA device" code
    import torch
    real = os.path.join(ROOT, "sparkl_b200", "libsparkl_b200.so")
    if os.path.exists(real) and not torch.cuda.is_available():
        exe2 = str(tmp_path / "sand_column")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I" + inc, src, "-o", exe2, real, "-Wl,-rpath," + os.path.dirname(real)])
        assert subprocess.run([exe2, "4", "2", "2", "1"], capture_output=True).returncode == 10


@pytest.mark.parametrize("name", ["F1M_fluids3", "S4M_sand3", "C16M_cube_through_sand3", "N8M_neo_hookean_slab"])
def test_size_independent_properties_on_the_emulation(emulated, monkeypatch, name):
    """the property checks of tests/test_fullsize_gpu.py (bit-exact block membership recomputed on the host, active set, exclusive-scan
    headers, sort idempotence, P2G mass / momentum conservation, update invariants) on the same scenes at sizes the emulation handles"""
    import tests.test_fullsize_gpu as T
    from sparkl_b200 import scenes
    small = {"F1M_fluids3": lambda: scenes.fluids3(24, 12, 12), "S4M_sand3": lambda: scenes.sand3(16, 8, 8),
             "C16M_cube_through_sand3": lambda: scenes.cube_through_sand3(20, 10, 10, 5), "N8M_neo_hookean_slab": lambda: scenes.neo_hookean_block(14, 14, 14)}
    monkeypatch.setattr(T, "CONFIGS", small)
    T.test_full_size_properties(name)


def test_custom_particle_kernels_on_the_emulation(emulated, monkeypatch):
    """SURVEY §8 f-4 without a GPU: examples/custom_updater (a user's updater type instantiating the library's G2P2G / dt kernel
    templates in its own shared object, plugged in through spk_set_particle_kernels_launcher) built for the emulation; the bodies of
    tests/test_custom_updater_gpu.py run on it"""
    import build_emu
    import tests.test_custom_updater_gpu as T
    monkeypatch.setattr(T, "EXAMPLE", build_emu.build_example())
    example = T._load_example()
    T.test_custom_updater_reproduces_the_builtin_model(example)
    T.test_default_launcher_treats_custom_as_zero_stress_and_null_restores_it(example)
    T.test_custom_updater_damping_and_stage_calls(example)


def test_dim2_trimesh_on_the_emulation(emulated):
    import tests.test_ztrimesh2d_gpu as T
    T.test_sand_on_a_trimesh_ramp_dim2()
    T.test_trimesh_argument_errors_dim2()


def test_remaining_example_scenes_on_the_emulation(emulated, oracle):
    import tests.test_zscenes_gpu as T
    T.test_sand_penetration3_against_the_rotated_heightfield_wall(oracle)
    T.test_l_panel2_lockstep()


def test_several_contexts_in_one_step_call_on_the_emulation(emulated, oracle):
    import tests.test_zscenes_gpu as T
    T.test_several_contexts_in_one_call(oracle)
    T.test_several_contexts_in_one_call_dim2()


def _dim2_scene():
    from sparkl_b200 import scenes2d
    return scenes2d.jitter(scenes2d.elasticity2(2, scale=0.3), 0.002, seed=9)


def _emulated_dim2_worker(rank, world, port, nframes, out_dir):
    """one rank of a 2-process run of the dim2 PRODUCT (emulated libsparkl_b200_2d) over gloo through the portable halo protocol"""
    import numpy as np
    import torch
    import torch.distributed as dist
    import build_emu
    from sparkl_b200 import _lib
    _lib.LIB_PATH, _lib.LIB2D_PATH, _lib._LIBS, _lib._LIB = build_emu.build(3), build_emu.build(2), {}, None
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from sparkl_b200.distributed import DistributedMpmPipeline
        from sparkl_b200.pipeline import SingleGpuMpmContext
        sc = _dim2_scene()
        ids = np.arange(len(sc.particles))[rank::world].copy()  # interleaved ownership: every block is shared (any ownership is valid)
        e = SingleGpuMpmContext(sc.cell_width, dim=2)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders)
        e.upload(sc.particles.take(ids))
        pipe = DistributedMpmPipeline(e, dist, torch.device("cpu"))
        assert pipe.payload_floats == 96
        halo = 0
        for _ in range(nframes):
            pipe.step(sc.params, sc.gravity)
            halo = max(halo, pipe.halo_blocks_last)
        pc = e.read_all()
        np.savez(os.path.join(out_dir, f"d2_{rank}.npz"), ids=ids, pos=pc.position, vel=pc.velocity, vol=pc.volume, halo=halo)
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_ranks_dim2_on_the_emulation(tmp_path):
    """the dim2 build on the portable multi-process path (spk_halo_* of spk_grid2d.cu, 96 floats per shared block): two gloo ranks
    with interleaved particle ownership against the single-rank dim2 oracle run"""
    import numpy as np
    import torch.multiprocessing as mp
    import build_emu
    from oracle import binding as ob
    from sparkl_b200 import abi2d
    from tests import util
    from tests.test_distributed_cpu import _free_port
    build_emu.build(3); build_emu.build(2)
    nframes = 2
    mp.spawn(_emulated_dim2_worker, args=(2, _free_port(), nframes, str(tmp_path)), nprocs=2, join=True)
    sc = _dim2_scene()
    e = ob.OracleEngine2D(sc.cell_width)
    e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
    for _ in range(nframes):
        e.step(sc.params, sc.gravity)
    ref = e.read_all()
    r = [np.load(tmp_path / f"d2_{k}.npz") for k in range(2)]
    assert int(r[0]["halo"]) > 0 and int(r[1]["halo"]) > 0
    pos = np.zeros(len(ref), abi2d.POSITION); vel = np.zeros(len(ref), abi2d.VELOCITY); vol = np.zeros(len(ref), abi2d.VOLUME)
    for k in range(2):
        pos[r[k]["ids"]], vel[r[k]["ids"]], vol[r[k]["ids"]] = r[k]["pos"], r[k]["vel"], r[k]["vol"]
    assert util.rel_err(pos["point"], ref.position["point"]) < 2e-5
    assert util.rel_err(vel["vector"], ref.velocity["vector"], floor=1e-2) < 2e-3
    assert util.rel_err(vol["deformation_gradient"], ref.volume["deformation_gradient"], floor=1.0) < 2e-4


@pytest.mark.timeout(600)
def test_no_kernel_depends_on_zeroed_allocations():
    """cudaMalloc returns garbage on a GPU but calloc'ed memory here: re-run a cross-section of the bodies above in a subprocess whose
    emulated cudaMalloc fills every allocation with 0xff bytes (NaN floats, all-ones integers).  scripts/emu_sanitize.sh runs the whole
    module this way (two poison bytes) and under AddressSanitizer."""
    import subprocess
    if os.environ.get("SPK_EMU_POISON"):
        pytest.skip("already running poisoned")
    picks = ("test_sand_first_substeps or test_fluid_dam_break or test_sand_on_heightfield_ground or test_elasticity2_lockstep or "
             "fluids_volume or device_side_migration_on or test_ragged_blocks or sand3_jitter")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-k", picks, "-p", "no:cacheprovider"],
                       env=dict(os.environ, SPK_EMU_POISON="0xff"), capture_output=True, text=True, cwd=ROOT, timeout=580)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert " passed" in r.stdout and "no tests ran" not in r.stdout
