"""Development aid: per-stage CUDA-event times of the S4M sand scene in the benchmark state (120 settle substeps) for the library
selected by SPK_LIB_PATH / SPK_G2P2G_MODE.  bench.py is the contract; this prints one short line per run.
  python scripts/g2p2g_perf.py [label] [steps] [settle] [ni nj nk]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from sparkl_b200 import abi
from sparkl_b200.pipeline import context_from_scene

label = sys.argv[1] if len(sys.argv) > 1 else "run"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
settle = int(sys.argv[3]) if len(sys.argv) > 3 else 120
dims = tuple(int(x) for x in sys.argv[4:7]) if len(sys.argv) > 6 else bench.FULL
sc, _ = bench.build_scene(0, 1, dims)
ctx = context_from_scene(sc, device=0)
params = abi.SolverParameters(stop_after_one_substep=1)
out = {"label": label, "mode": os.environ.get("SPK_G2P2G_MODE", "default"), "lib": os.path.basename(os.environ.get("SPK_LIB_PATH", "default"))}
for phase, nsettle in (("calm", 3), ("settled", settle)):
    for _ in range(nsettle):
        ctx.step(params, sc.gravity, timings=False)
    st = np.zeros(6)
    for _ in range(steps):
        t = ctx.step(params, sc.gravity, timings=True)[1]
        st += [t.grid_resize_and_sort_ms, t.reset_grid_ms, t.estimate_timestep_ms, t.g2p2g_ms, t.grid_update_ms, t.total_ms]
    st /= steps
    out[phase] = dict(zip(["sort", "reset", "dt", "g2p2g", "grid", "total"], st.round(4).tolist()))
    out[phase]["ndispatch"] = int(t.num_dispatch_blocks)
    out[phase]["frac"] = round(bench.A_G2P2G_SAND * sc.num_particles / (st[3] * 1e-3) / 1e9 / bench._peaks()[0], 4)
p = ctx.read(abi.ATTR_POSITION)["point"]
out["finite"] = bool(np.isfinite(p).all())
out["ymean"] = float(p[:, 1].mean())
print(json.dumps(out), flush=True)
