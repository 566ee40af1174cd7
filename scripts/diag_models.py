"""dev aid: product vs the reference's kernels, per-substep error distribution of F for one model_block kind"""
import sys, numpy as np
sys.path.insert(0, ".")
from sparkl_b200 import abi, scenes
from sparkl_b200.pipeline import context_from_scene
from oracle import binding as ob
kind = sys.argv[1]
sc = scenes.jitter(scenes.model_block(kind), 0.01, seed=3)
g = context_from_scene(sc)
r = ob.RefEngine(sc.cell_width); r.models_upload(sc.models); r.colliders_upload(sc.colliders); r.upload(sc.particles)
o = ob.OracleEngine(sc.cell_width); o.models_upload(sc.models); o.colliders_upload(sc.colliders); o.upload(sc.particles)
E = {"product": g, "oracle": o, "reference": r}
p = sc.params
for e in E.values(): e.frame_begin(p)
rem = np.float32(p.dt); min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
for s in range(4):
    for e in E.values(): e.stage_sort(); e.stage_reset_grid()
    dts = {k: e.stage_estimate_dt(min_dt, float(min(rem, np.float32(p.max_substep_dt)))) for k, e in E.items()}
    acc = r.dt_acc(); dt = dts["reference"]
    Fin = {k: e.read(abi.ATTR_VOLUME)["deformation_gradient"].copy() for k, e in E.items()}
    for e in E.values(): e.set_dt_acc(acc); e.stage_g2p2g(dt)
    F = {k: e.read(abi.ATTR_VOLUME)["deformation_gradient"] for k, e in E.items()}
    for k in ("product", "oracle"):
        d = np.abs(F[k].astype(np.float64) - F["reference"]).max(axis=1)
        i = int(d.argmax())
        sv = np.linalg.svd(F["reference"][i].reshape(3, 3).astype(np.float64), compute_uv=False)
        svin = np.linalg.svd(Fin["reference"][i].reshape(3, 3).astype(np.float64), compute_uv=False)
        print(f"substep {s} {k:8s}: max dF {d.max():.2e} median {np.median(d):.1e} #>1e-5 {(d > 1e-5).sum()} #>1e-4 {(d > 1e-4).sum()} worst particle {i} sv_in {svin} sv_out {sv} dFin {np.abs(Fin[k][i]-Fin['reference'][i]).max():.1e}")
    for e in E.values(): e.stage_grid_update(dt, sc.gravity)
    rem = np.float32(rem - np.float32(dt))
