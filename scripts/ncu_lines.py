"""per-CUDA-source-line stall samples / instruction share from an ncu report (needs -lineinfo)."""
import collections, csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.OrderedDict(); ts = ti = 0; cur = None; ci = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No":
        ci = {}
        for i, h in enumerate(r): ci.setdefault(h, i)
        continue
    if ci is None or len(r) < 8 or r[2] != "-": continue
    try: smp, ins = int(r[ci["# Samples"]]), int(r[ci["Instructions Executed"]])
    except ValueError: continue
    st = {h.replace("stall_", ""): int(r[i]) for h, i in ci.items() if h.startswith("stall_") and "Not Issued" not in h and r[i].isdigit()}
    a = agg.setdefault((cur, int(r[0]), r[1].strip()[:95]), [0, 0, collections.Counter()])
    a[0] += smp; a[1] += ins; a[2].update(st); ts += smp; ti += ins
print("total samples", ts, "warp instructions", ti)
for k, a in sorted(agg.items(), key=lambda x: -x[1][0])[:top]:
    print(f"{k[0]:14s}:{k[1]:4d} smp {a[0]/ts*100:5.1f}% ins {a[1]/ti*100:5.1f}% {','.join(f'{n}:{c}' for n,c in a[2].most_common(2)):28s}| {k[2]}")
