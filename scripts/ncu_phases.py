"""dev aid: per-phase shared-memory wavefronts / instructions / stall samples of one ncu report (phases = SASS ranges between barriers)
  python scripts/ncu_phases.py gpurun_out/s_prof.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
def op(r):
    t = r[ix["Source"]].split()
    return (t[1] if t and t[0].startswith("@") and len(t) > 1 else (t[0] if t else ""))
bounds = [i for i, r in enumerate(data) if op(r).startswith("BAR.SYNC") or "TRYWAIT" in op(r) or op(r) == "EXIT"]
tot_s = sum(f(r, "# Samples") for r in data); tot_i = sum(f(r, "Instructions Executed") for r in data); tot_w = sum(f(r, "L1 Wavefronts Shared") for r in data)
print(f"total: {tot_i/1e6:.1f}M warp instr, {tot_w/1e6:.1f}M shared wavefronts, {int(tot_s)} samples")
a = 0
for b in bounds:
    sub = data[a:b + 1]
    ins = sum(f(r, "Instructions Executed") for r in sub); wf = sum(f(r, "L1 Wavefronts Shared") for r in sub); sm = sum(f(r, "# Samples") for r in sub)
    ops = {}
    for r in sub:
        w = f(r, "L1 Wavefronts Shared")
        if w: ops[op(r)] = ops.get(op(r), 0) + w
    if ins > 0.002 * tot_i or sm > 0.01 * tot_s:
        print(f"{a:5d}-{b:5d}  instr {100*ins/tot_i:5.1f}%  samples {100*sm/tot_s:5.1f}%  wf {wf/1e6:6.1f}M ", {k: round(v / 1e6, 1) for k, v in sorted(ops.items(), key=lambda x: -x[1])[:5]})
    a = b + 1
