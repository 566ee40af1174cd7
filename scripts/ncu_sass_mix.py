"""instruction mix (by SASS opcode) and per-segment totals of the top kernel in an ncu report: where the issue slots go."""
import collections, csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = None; mix = collections.Counter(); tot = 0; thr = 0; smp = collections.Counter()
for r in rows:
    if r and r[0] == "Address": hdr = {h: i for i, h in enumerate(r)}; continue
    if hdr is None or len(r) < 10 or not r[0].startswith("0x"): continue
    ins = int(r[hdr["Instructions Executed"]]); t = int(r[hdr["Thread Instructions Executed"]])
    op = r[1].split()
    op = [o for o in op if not o.startswith("@")][0].split(".")[0] if op else "?"
    mix[op] += ins; tot += ins; thr += t; smp[op] += int(r[hdr["# Samples"]])
n = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
print(f"warp instr {tot}  thread instr {thr}  avg active {thr/tot:.1f}  per particle (n={n:g}): {tot*32/n:.0f} slots")
ts = sum(smp.values())
for op, c in mix.most_common(40):
    print(f"{op:12s} {c/tot*100:6.2f}%  {c*32/n:8.1f} slots/particle   samples {smp[op]/ts*100:5.1f}%")
