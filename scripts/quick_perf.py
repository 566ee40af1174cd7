"""Quick per-stage timing of the sand scene (development aid; bench.py is the contract)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene

ni, nj, nk = (int(x) for x in (sys.argv[1:4] if len(sys.argv) >= 4 else (200, 100, 100)))
frames = int(sys.argv[4]) if len(sys.argv) > 4 else 6
scene_name = sys.argv[5] if len(sys.argv) > 5 else "sand3"
t0 = time.time()
sc = scenes.SCENES[scene_name](ni, nj, nk)
print("scene", scene_name, sc.num_particles, "particles, built in %.1fs" % (time.time() - t0), flush=True)
ctx = context_from_scene(sc)
for f in range(frames):
    t0 = time.time()
    sim, t = ctx.step(sc.params, sc.gravity)
    wall = time.time() - t0
    ns = max(t.num_substeps, 1)
    if f < 3 or f % 8 == 0 or f == frames - 1: print(f"frame {f}: substeps {t.num_substeps} dt {t.last_dt:.3e} nab {t.num_active_blocks} ndisp {t.num_dispatch_blocks} | per-substep ms: sort {t.grid_resize_and_sort_ms/ns:.3f} "
          f"reset {t.reset_grid_ms/ns:.3f} dt {t.estimate_timestep_ms/ns:.3f} g2p2g {t.g2p2g_ms/ns:.3f} grid {t.grid_update_ms/ns:.3f} total {t.total_ms/ns:.3f} wall {wall*1e3/ns:.3f} "
          f"| {sc.num_particles*ns/(t.total_ms*1e-3)/1e9:.3f} G p-substeps/s", flush=True)
p = ctx.read(abi.ATTR_POSITION)["point"]
print("finite", np.isfinite(p).all(), "ymin", p[:, 1].min(), "ymean", p[:, 1].mean())
