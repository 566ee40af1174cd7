"""dev aid: executed warp instructions / stall samples per SOURCE LINE of one kernel: joins the SASS rows of an ncu report with the
line table of the same kernel in a library (nvdisasm -g), instruction by instruction.
  python scripts/ncu_lines.py gpurun_out/s_prof.ncu-rep sparkl_b200/libsparkl_b200.so <kernel substring> [top N]"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, lib, pat = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
def f(r, k):
    try: return float(r[ix[k]])
    except Exception: return 0.0
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, stdout=subprocess.DEVNULL)
lines = None
for fn in os.listdir(tmp):
    dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, fn)], capture_output=True, text=True).stdout
    on, cur, out, stack = False, None, [], None
    for line in dis.splitlines():
        if line.startswith("//--------------------- .text."):
            if on and len(out) == len(data): break
            on = pat in line
            out = []
            continue
        if not on: continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
        if m:
            cur = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        if re.match(r"\s+/\*[0-9a-f]+\*/", line): out.append(cur)
    if on and len(out) == len(data): lines = out
    if lines: break
assert lines, "kernel not found / instruction count mismatch (is the library the profiled build?)"
agg = collections.defaultdict(lambda: [0.0, 0.0])
for r, l in zip(data, lines):
    agg[l][0] += f(r, "Instructions Executed"); agg[l][1] += f(r, "# Samples")
ti = sum(v[0] for v in agg.values()); ts = sum(v[1] for v in agg.values())
byfile = collections.defaultdict(lambda: [0.0, 0.0])
for (fn, ln), v in agg.items():
    byfile[fn][0] += v[0]; byfile[fn][1] += v[1]
for fn, v in sorted(byfile.items(), key=lambda x: -x[1][0]): print(f"{fn:32s} instr {100*v[0]/ti:5.1f}%  samples {100*v[1]/ts:5.1f}%")
print()
for (fn, ln), v in sorted(agg.items(), key=lambda x: -x[1][0])[:top]: print(f"{fn}:{ln:<5d} instr {100*v[0]/ti:5.2f}%  samples {100*v[1]/ts:5.2f}%")
