"""Throughput of the dim2 build on a B200: particle-substeps/s through spk_step on the reference's 2D scenes at benchmark size
(fluids2 at its own 300 x 300 and at 1000 x 1000; elasticity2 stars).  Usage: python scripts/dim2_perf.py [out.json]"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from sparkl_b200 import abi, scenes2d  # noqa: E402
from sparkl_b200.pipeline import context_from_scene  # noqa: E402


def run(scene, nsub, warm):
    g = context_from_scene(scene)
    p = abi.SolverParameters(stop_after_one_substep=1, max_num_substeps=scene.params.max_num_substeps, max_substep_dt=scene.params.max_substep_dt)
    for _ in range(warm):
        g.step(p, scene.gravity)
    g.synchronize()
    ms = g2p2g = 0.0
    t0 = time.perf_counter()
    for _ in range(nsub):
        _, t = g.step(p, scene.gravity)
        ms += t.total_ms
        g2p2g += t.g2p2g_ms
    wall = time.perf_counter() - t0
    n = len(scene.particles)
    return {"scene": scene.name, "particles": n, "substeps": nsub, "ms_per_substep": ms / nsub, "g2p2g_ms": g2p2g / nsub,
            "particle_substeps_per_s": n * nsub / (ms * 1e-3), "wall_ms_per_substep": wall * 1e3 / nsub, "active_blocks": t.num_active_blocks,
            "kernel_launches_per_substep": t.kernel_launches}


if __name__ == "__main__":
    out = [run(scenes2d.fluids2(300, 300), 60, 10), run(scenes2d.fluids2(1000, 1000), 40, 10), run(scenes2d.elasticity2(nblocks=5), 60, 10)]
    txt = json.dumps(out, indent=1)
    print(txt)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(txt)
