"""dev aid: where does the pipelined end-to-end frame time go"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import bench
from sparkl_b200 import abi
from sparkl_b200.pipeline import context_from_scene
sc = bench.build_scene(0, 1, bench.FULL)
ctx = context_from_scene(sc)
host = sc.particles
pins = [bench.pinned_like(a) for a in (host.status, host.position, host.velocity, host.volume, host.phase, host.data)]
pc = abi.ParticleComponents(0)
pc.status, pc.position, pc.velocity, pc.volume, pc.phase, pc.data = [p[0] for p in pins]
outs = {w: bench.pinned_like(np.empty(len(host), abi.ATTR_DTYPES[w])) for w in (abi.ATTR_POSITION, abi.ATTR_VELOCITY, abi.ATTR_PHASE)}
frame = abi.SolverParameters()
def run(name, fn, n=6):
    fn(); ctx.read_wait(); ctx.synchronize(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for f in range(n): fn()
    ctx.read_wait(); ctx.synchronize(); torch.cuda.synchronize()
    print(f"{name:50s} {(time.perf_counter()-t0)/n*1e3:8.2f} ms/frame", flush=True)
run("step only (6 substeps)", lambda: ctx.step(frame, sc.gravity, timings=False))
run("blocking upload only", lambda: ctx.upload(pc))
run("upload_async (stage+commit) only", lambda: ctx.upload_async(pc))
run("upload_async + step", lambda: (ctx.upload_async(pc), ctx.step(frame, sc.gravity, timings=False)))
def pipelined():
    ctx.commit_staged(); ctx.upload_async(pc, stage_only=True); ctx.step(frame, sc.gravity, timings=False)
ctx.upload_async(pc, stage_only=True)
run("commit + stage(next) + step", pipelined)
def pipelined_reads():
    ctx.commit_staged(); ctx.upload_async(pc, stage_only=True); ctx.step(frame, sc.gravity, timings=False)
    for w, o in outs.items(): ctx.read_async(w, o[0])
run("commit + stage(next) + step + async reads", pipelined_reads)
def blocking_all():
    ctx.upload(pc); ctx.step(frame, sc.gravity, timings=False)
    for w, o in outs.items(): ctx.read(w, out=o[0])
run("blocking upload + step + blocking reads", blocking_all)
