"""TEST INFRASTRUCTURE: builds libsparkl_b200_emu.so / libsparkl_b200_emu_2d.so — the PRODUCT'S OWN sources (sparkl_b200/csrc/*.cu,
*.cuh: kernels and host sequencing) compiled for the host with g++ on top of the CUDA emulation shim of this directory (cuda_emu.h).
The only source transformation is the launch syntax: `kernel<<<grid, block, smem, stream>>>(args);` becomes
`emu_launch(grid, block, [&] { kernel(args); });`.  Never used by the product (sparkl_b200/_lib.py does not know these files)."""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(HERE)))
CSRC = os.path.join(ROOT, "sparkl_b200", "csrc")
OUT_DIR = os.path.dirname(HERE)  # tests/host
LAUNCH = re.compile(r"([A-Za-z_][\w:]*(?:<[^<>;{}()]*>)?)\s*<<<(.*?)>>>\s*\((.*?)\);", re.S)


def _split_top(args: str):
    out, depth, cur = [], 0, ""
    for ch in args:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    out.append(cur.strip())
    return out


def transform(text: str) -> str:
    def repl(m):
        cfg = _split_top(m.group(2))
        return f"emu_launch((unsigned)({cfg[0]}), (unsigned)({cfg[1]}), [&] {{ {m.group(1)}({m.group(3)}); }});"
    text = LAUNCH.sub(repl, text)
    return text.replace('#include "../../include/sparkl_abi.h"', f'#include "{os.path.join(ROOT, "include", "sparkl_abi.h")}"')


def lib_path(dim: int, tag: str = "") -> str:
    return os.path.join(OUT_DIR, f"libsparkl_b200_emu{tag}.so" if dim == 3 else f"libsparkl_b200_emu{tag}_2d.so")


def build(dim: int = 3, force: bool = False, tag: str = "", extra_flags=()) -> str:
    """`tag` / `extra_flags`: instrumented variants (e.g. tag="_asan", extra_flags=["-fsanitize=address", "-g"])"""
    extra_flags = list(extra_flags)
    sys.path.insert(0, ROOT)
    from sparkl_b200 import build as pb
    sources = pb.SOURCES if dim == 3 else pb.SOURCES_2D
    out = lib_path(dim, tag)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))] + [os.path.join(HERE, f) for f in ("cuda_emu.h", "cuda_emu.cpp", "build_emu.py")]
    deps.append(os.path.join(ROOT, "include", "sparkl_abi.h"))
    if not force and os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(d) for d in deps):
        return out
    src_dir = os.path.join(HERE, f"_src{dim}{tag}")
    os.makedirs(src_dir, exist_ok=True)
    for f in os.listdir(CSRC):
        if f.endswith((".cu", ".cuh")):
            with open(os.path.join(CSRC, f)) as fh:
                text = transform(fh.read())
            with open(os.path.join(src_dir, f.replace(".cu", ".cpp") if f.endswith(".cu") else f), "w") as fh:
                fh.write(text)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    flags = ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-U_FORTIFY_SOURCE", "-DSPK_EMU", f"-DSPK_DIM={dim}", "-I" + cuda_inc, "-I" + HERE, "-include", "cuda_emu.h"]
    flags += extra_flags
    objs = []
    for s in list(sources) + ["cuda_emu"]:
        src = os.path.join(HERE, "cuda_emu.cpp") if s == "cuda_emu" else os.path.join(src_dir, s.replace(".cu", ".cpp"))
        obj = os.path.join(src_dir, os.path.basename(src) + ".o")
        r = subprocess.run(["g++"] + flags + ["-c", src, "-o", obj], capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"emulation build failed on {s}")
        objs.append(obj)
    subprocess.check_call(["g++", "-shared", "-Wl,-Bsymbolic", "-o", out] + [f for f in extra_flags if f.startswith("-fsanitize")] + objs)
    return out


def build_example() -> str:
    """examples/custom_updater/custom_updater.cu (a user's own instantiation of the library's kernel templates) for the emulation:
    compiled against the transformed headers of the dim3 build, with its own copy of the CUDA stand-ins"""
    lib3 = build(3)
    src = os.path.join(ROOT, "examples", "custom_updater", "custom_updater.cu")
    out = os.path.join(OUT_DIR, "libcustom_updater_emu.so")
    if os.path.exists(out) and os.path.getmtime(out) >= max(os.path.getmtime(src), os.path.getmtime(lib3), os.path.getmtime(os.path.abspath(__file__))):
        return out
    src_dir = os.path.join(HERE, "_src3")
    with open(src) as fh:
        text = transform(fh.read()).replace('#include "../../sparkl_b200/csrc/spk_g2p2g_kernel.cuh"', '#include "spk_g2p2g_kernel.cuh"')
    cpp = os.path.join(src_dir, "custom_updater.cpp")
    with open(cpp, "w") as fh:
        fh.write(text)
    cuda_inc = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
    flags = ["-O1", "-std=c++17", "-fPIC", "-ffp-contract=off", "-w", "-U_FORTIFY_SOURCE", "-DSPK_EMU", "-DSPK_DIM=3", "-I" + cuda_inc, "-I" + HERE, "-I" + src_dir, "-include", "cuda_emu.h"]
    # like the CUDA build of the example (its own .so with its own CUDA runtime), the emulated one carries its own copy of the CUDA
    # stand-ins, bound with -Bsymbolic: a real libcudart that some other module (torch) loaded into the process must not capture them
    subprocess.check_call(["g++"] + flags + ["-shared", "-Wl,-Bsymbolic", "-o", out, cpp, os.path.join(HERE, "cuda_emu.cpp")])
    return out


if __name__ == "__main__":
    for d in (3, 2):
        print(build(d, force="--force" in sys.argv))
