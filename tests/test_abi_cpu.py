"""The C-ABI library loads on a GPU-less host, exports every symbol include/sparkl_abi.h declares, and refuses to
compute without a CUDA device (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

from sparkl_b200 import abi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sparkl_abi.h")).read()
    return sorted(set(re.findall(r"^\s*(?:int|const char\*)\s+(spk_[a-z0-9_]+)\s*\(", src, re.M)))


def test_library_exports_every_declared_symbol():
    from sparkl_b200 import _lib
    l = _lib.lib()
    declared = _declared_symbols()
    assert len(declared) >= 30
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(l, name), name
    assert l.spk_abi_version() == 1


def test_header_layouts_compile_in_c_and_cpp(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "sparkl_abi.h"\nint main(void){return sizeof(SpkGpuParticleModel)==96 && sizeof(SpkGpuCollider)==336 && sizeof(SpkGpuGridNode)==72 ? 0 : 1;}\n')
    for cc, std in (("gcc", "-std=c11"), ("g++", "-std=c++17")):
        exe = tmp_path / ("t_" + cc)
        args = [cc, std, "-x", "c" if cc == "gcc" else "c++", str(src), "-I", os.path.join(ROOT, "include"), "-o", str(exe)]
        subprocess.check_call(args)
        assert subprocess.call([str(exe)]) == 0


def test_model_and_collider_images():
    m = abi.ParticleModel.with_plasticity(abi.CorotatedLinearElasticity(1e7, 0.2), abi.DruckerPragerPlasticity(1e7, 0.2))
    b = np.frombuffer(m.to_gpu_model(), np.uint8)
    assert len(b) == 96
    assert np.frombuffer(b, "<u4", 1, 0)[0] == 0 and np.frombuffer(b, "<u4", 1, 32)[0] == 0 and np.frombuffer(b, "<u4", 1, 80)[0] == 2
    lam, mu = abi.lame_lambda_mu(1e7, 0.2)
    assert np.frombuffer(b, "<f4", 2, 16).tolist() == [lam, mu]            # corotated lambda, mu @16, @20
    assert np.frombuffer(b, "<f4", 1, 40)[0] == np.float32(np.float32(35.0) * np.float32(np.float32(np.pi) / np.float32(180.0)))
    assert np.frombuffer(b, "<f4", 1, 68)[0] == 1.0                        # volume_correction
    e = abi.ParticleModel(abi.NeoHookeanElasticity(1e6, 0.3))
    b = np.frombuffer(e.to_gpu_model(), np.uint8)
    assert np.frombuffer(b, "<u4", 1, 0)[0] == 1 and np.frombuffer(b, "<u4", 1, 32)[0] == 5      # plastic None = 5
    c = abi.CuboidCollider((1, 2, 3), (4, 5, 6), friction=0.25, grid_boundary_handling=abi.BoundaryHandling.Stick)
    b = np.frombuffer(c.to_gpu_collider(), np.uint8)
    assert len(b) == 336 and np.frombuffer(b, "<f4", 3, 8).tolist() == [1, 2, 3]
    assert np.frombuffer(b, "<f4", 4, 296).tolist() == [0, 0, 0, 1] and np.frombuffer(b, "<f4", 3, 312).tolist() == [4, 5, 6]
    assert np.frombuffer(b, "<f4", 1, 324)[0] == 0.25 and np.frombuffer(b, "<u4", 1, 332)[0] == 0


def test_no_cpu_fallback():
    """without a CUDA device every compute entry point must fail loudly (never route through the oracle)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from sparkl_b200.pipeline import SingleGpuMpmContext
    with pytest.raises(abi.SparklError) as ei:
        SingleGpuMpmContext(0.2)
    assert ei.value.code > 0  # a cudaError_t, not an SPK_E* code
    # the product package never imports, links or loads the oracle
    for root, _, files in os.walk(os.path.join(ROOT, "sparkl_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(root, f)).read()
                for needle in ("import oracle", "from oracle", "liboracle", "mpm_oracle", "oracle/"):
                    assert needle not in text, (root, f, needle)


def test_lame_parameters_f32():
    lam, mu = abi.lame_lambda_mu(1.0e7, 0.2)
    assert mu == np.float32(np.float32(1.0e7) / (np.float32(2.0) * (np.float32(1.0) + np.float32(0.2))))
    assert float(lam) == pytest.approx(1e7 * 0.2 / (1.2 * 0.6), rel=1e-6)


def test_custom_kernel_launcher_abi_and_example_build():
    """the launcher entry points are exported by both libraries, misuse is an SPK_E* code (no GPU needed), the
    launcher struct is four pointers, and the example that instantiates the library's kernel template with a user updater builds"""
    import ctypes as C
    import subprocess
    from sparkl_b200 import _lib, build
    for dim in (2, 3):
        l = _lib.lib(dim)
        assert l.spk_set_particle_kernels_launcher(None, None) == abi.EINVAL
        assert l.spk_default_launch_g2p2g(None, C.c_float(0.1)) == abi.EINVAL
    src = '#include <stdio.h>\n#include "sparkl_abi.h"\nint main(void) { printf("%zu", sizeof(SpkParticleKernelsLauncher)); return 0; }\n'
    exe = os.path.join(ROOT, "tests", "host", "_launcher_size")
    subprocess.run(["gcc", "-x", "c", "-I" + os.path.join(ROOT, "include"), "-o", exe, "-"], input=src.encode(), check=True)
    assert subprocess.check_output([exe]).decode() == "32"
    os.remove(exe)
    build.build()
    ex = C.CDLL(build.EXAMPLE_LIB)
    for name in ("cu_launcher_create", "cu_launcher_vtable", "cu_launcher_calls", "cu_launcher_destroy"):
        assert hasattr(ex, name)
    from tests.host.emu2d import DEVMODEL
    assert ex.cu_sizeof_devmodel() == DEVMODEL.itemsize


def test_abi_header_is_plain_c_in_both_dimensions(tmp_path):
    """include/sparkl_abi.h is the drop-in boundary: it must compile as C11 (no C++-isms) for the dim3 and the dim2 build, and the wire
    record of the migration entry points is 128 bytes in dim3"""
    src = tmp_path / "abi.c"
    src.write_text('#include "sparkl_abi.h"\n#include <stdio.h>\nint main(void) { printf("%d %d %d\\n", (int)sizeof(SpkParticleRecord), (int)sizeof(SpkGpuCollider), '
                   '(int)sizeof(SpkGpuGridNode)); return 0; }\n')
    inc = os.path.join(ROOT, "include")
    for dim, want in ((3, "128 336 72"), (2, "104 72 56")):
        exe = str(tmp_path / f"abi{dim}")
        subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", f"-DSPK_DIM={dim}", "-I" + inc, str(src), "-o", exe])
        assert subprocess.check_output([exe]).decode().strip() == want
