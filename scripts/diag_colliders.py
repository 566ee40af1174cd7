"""dev aid (GPU): per-node comparison product / oracle vs the reference's kernels after ONE grid_update, for the collider scenes whose
three-way tests disagree: how many nodes differ, in which field, and what those nodes look like."""
import sys, numpy as np
sys.path.insert(0, ".")
from oracle import binding as ob
from sparkl_b200 import abi, scenes
from tests import util
which = sys.argv[1] if len(sys.argv) > 1 else "hf3"
from tests.test_reference_gpu import _three
if which == "hf3":
    sc = scenes.jitter(scenes.sand_penetration3(8, 6, 6, origin=(0.0, 13.7, 3.52)), 0.02, seed=12)
else:
    import tests.test_dim2_gpu as T2
    sc = T2._elasticity2_on_the_ground_scene()
E = _three(sc, ob)
p = sc.params
for e in E.values():
    e.frame_begin(p); e.stage_sort(); e.stage_reset_grid()
min_dt = float(np.float32(p.dt) / np.float32(p.max_num_substeps))
dts = {k: e.stage_estimate_dt(min_dt, float(min(np.float32(p.dt), np.float32(p.max_substep_dt)))) for k, e in E.items()}
dt = dts["reference"]
for e in E.values():
    e.set_dt_acc(E["reference"].dt_acc()); e.stage_g2p2g(dt)
for e in E.values():
    e.stage_grid_update(dt, sc.gravity)
tab = {k: util.nodes_by_key(e, 0) for k, e in E.items()}
vr, nr = tab["reference"]
for k in ("product", "oracle"):
    v, n = tab[k]
    assert np.array_equal(v, vr)
    for f in ("collision_normal", "momentum_velocity", "projection_scaled_dir"):
        d = np.abs(n[f].astype(np.float64) - nr[f]).max(axis=-1)
        bad = np.argwhere(d > 1e-3)
        print(k, f, "nodes differing > 1e-3:", len(bad), "of", d.size, "max", d.max())
        for b, c in bad[:6]:
            print("   block", int(v[b]), "cell", c, "ours", n[f][b, c], "ref", nr[f][b, c], "tag ours/ref", n["projection_status_tag"][b, c], nr["projection_status_tag"][b, c],
                  "payload", n["projection_status_payload"][b, c], nr["projection_status_payload"][b, c], "mass", nr["mass"][b, c], "dir ref", nr["projection_scaled_dir"][b, c])
    t = n["projection_status_tag"] != nr["projection_status_tag"]
    print(k, "tags differing:", int(t.sum()))
if which != "hf3":
    v, n = tab["product"]
    d = np.abs(n["momentum_velocity"].astype(np.float64) - nr["momentum_velocity"]).max(axis=-1)
    order = np.argsort(d.ravel())[::-1][:10]
    for o in order:
        b, c = np.unravel_index(o, d.shape)
        print("node", int(v[b]), c, "diff", d[b, c], "ours", n["momentum_velocity"][b, c], "ref", nr["momentum_velocity"][b, c], "tag", nr["projection_status_tag"][b, c], "normal", nr["collision_normal"][b, c],
              "dir", nr["projection_scaled_dir"][b, c], "mass", nr["mass"][b, c])
    print("gravity", sc.gravity, "dt", dt, "h", sc.cell_width, [ (c.friction, c.grid_boundary_handling) for c in sc.colliders])
