"""In-tree build of libsparkl_b200.so (hand-written CUDA for sm_100a + the C ABI).  `python -m sparkl_b200.build`."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
SOURCES = ["spk_api.cu", "spk_sort.cu", "spk_g2p2g.cu", "spk_grid.cu", "spk_dist.cu", "spk_fluids.cu", "spk_migrate.cu"]
# the dim2 library: same host sequencing / sort / ABI sources compiled with -DSPK_DIM=2, its own stencil kernels
SOURCES_2D = ["spk_api.cu", "spk_sort.cu", "spk_g2p2g2d.cu", "spk_grid2d.cu", "spk_dist.cu", "spk_fluids.cu", "spk_migrate.cu"]
HEADERS = ["spk_internal.cuh", "spk_device.cuh", "spk_math.cuh", "spk_math2d.cuh", "spk_g2p2g2d_body.cuh", "spk_updater.cuh", "spk_g2p2g_kernel.cuh", "spk_g2p2g2d_kernel.cuh", "../../include/sparkl_abi.h"]
LIB = os.path.join(HERE, "libsparkl_b200.so")
LIB_2D = os.path.join(HERE, "libsparkl_b200_2d.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# IEEE division / sqrt and no flush-to-zero are REQUIRED: block membership depends on div.rn + round-half-away
# (SURVEY.md §3.3), so no --use_fast_math.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v",
         "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"] + os.environ.get("SPK_NVCC_DEFS", "").split()


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


EXAMPLE_SRC = os.path.join(os.path.dirname(HERE), "examples", "custom_updater", "custom_updater.cu")
EXAMPLE_LIB = os.path.join(os.path.dirname(HERE), "examples", "custom_updater", "libcustom_updater.so")


def build(force: bool = False, verbose: bool = False) -> str:
    """builds both libraries (and the custom-updater example that instantiates the library's kernel template); returns the path of
    the 3D one"""
    _build_one(SOURCES, ".o", [], LIB, force, verbose)
    _build_one(SOURCES_2D, ".2d.o", ["-DSPK_DIM=2"], LIB_2D, force, verbose)
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    if os.path.exists(EXAMPLE_SRC) and (force or _stale(EXAMPLE_LIB, [EXAMPLE_SRC] + hdrs)):
        cmd = [NVCC] + [f for f in FLAGS if f not in ("-Xptxas", "-v")] + ["-shared", "-o", EXAMPLE_LIB, EXAMPLE_SRC]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError("nvcc failed on the custom-updater example")
    return LIB


def _build_one(sources, obj_suffix, defs, lib, force, verbose) -> str:
    hdrs = [os.path.join(CSRC, h) for h in HEADERS]
    objs = []
    for s in sources:
        src = os.path.join(CSRC, s)
        obj = os.path.join(CSRC, s.replace(".cu", obj_suffix))
        if force or _stale(obj, [src] + hdrs):
            cmd = [NVCC] + FLAGS + defs + ["-c", src, "-o", obj]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if verbose or r.returncode:
                sys.stderr.write(r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError(f"nvcc failed on {s}")
        objs.append(obj)
    if force or _stale(lib, objs):
        # -Bsymbolic: the two libraries export the same names (like the sparkl2d / sparkl3d crates); calls between the objects of
        # one library must bind inside it even when both are loaded into one process
        cmd = [NVCC, "-shared", "-Xlinker", "-Bsymbolic", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
                                                     "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return lib


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
