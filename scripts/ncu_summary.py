"""Turn ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.

  python scripts/ncu_summary.py launches gpurun_out/launches.csv profiles/r1_launches_<tag>.md
  python scripts/ncu_summary.py kernel   gpurun_out/prof.ncu-rep  profiles/r1_<kernel>_<tag>  [--json profiles/g2p2g_ncu_summary.json]
"""
import collections
import csv
import json
import re
import subprocess
import sys

KEY = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
       "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread", "launch__occupancy_limit_registers",
       "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
       "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
       "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed_op_shared_atom.sum",
       "lts__t_bytes.sum", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "sm__inst_executed_pipe_fma.sum",
       "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"]


def _lib_sha16():
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    return bench._lib_sha16()


def launches(src, dst):
    rows = list(csv.reader(open(src)))
    for i, r in enumerate(rows):
        if r and r[0] == "ID":
            hdr, start = r, i + 1
            break
    ci = {h: i for i, h in enumerate(hdr)}
    agg, order = collections.OrderedDict(), []
    for r in rows[start:]:
        if len(r) < len(hdr):
            continue
        name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void <unnamed>::", "").replace("<unnamed>::", "")
        v = float(r[ci["Metric Value"]])
        v = v / 1000.0 if r[ci["Metric Unit"]] == "ns" else v * 1000.0 if r[ci["Metric Unit"]] == "ms" else v
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += v
        order.append((r[ci["ID"]], name, r[ci["Grid Size"]], r[ci["Block Size"]], v))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}) — `--metrics gpu__time_duration.sum --clock-control none` (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total us | avg us | share |\n|---|---:|---:|---:|---:|\n")
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| {k} | {a[0]} | {a[1]:.1f} | {a[1] / a[0]:.2f} | {a[1] / tot * 100:.1f}% |\n")
        f.write("\n## last 24 launches (one steady-state substep and a half)\n\n| id | kernel | grid | block | us |\n|---|---|---|---|---:|\n")
        for o in order[-24:]:
            f.write(f"| {o[0]} | {o[1]} | {o[2]} | {o[3]} | {o[4]:.2f} |\n")
    print(open(dst).read())


def kernel(src, dst, json_out=None):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, unit, vals = rows[0], rows[1], rows[2]
    m = {h: (vals[i], unit[i]) for i, h in enumerate(hdr)}
    srcpage = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(srcpage.splitlines()))
    shdr = srows[1]
    ci = {h: i for i, h in enumerate(shdr)}
    data = srows[2:]
    tot_i = sum(int(r[ci["Instructions Executed"]]) for r in data)
    tot_s = sum(int(r[ci["# Samples"]]) for r in data)
    stall_cols = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
    segs, cur = [], {"inst": 0, "samp": 0, "start": 0, "st": collections.Counter()}
    for k, r in enumerate(data):
        cur["inst"] += int(r[ci["Instructions Executed"]])
        cur["samp"] += int(r[ci["# Samples"]])
        for s in stall_cols:
            cur["st"][s] += int(r[ci[s]])
        if "BAR.SYNC" in r[ci["Source"]] or "SYNCS.PHASECHK" in r[ci["Source"]]:
            cur["end"], cur["marker"] = k, r[ci["Source"]].strip().split()[0]
            segs.append(cur)
            cur = {"inst": 0, "samp": 0, "start": k + 1, "st": collections.Counter()}
    cur["end"], cur["marker"] = len(data), "EXIT"
    segs.append(cur)
    stalls = collections.Counter()
    for s in segs:
        stalls.update(s["st"])
    with open(dst + ".md", "w") as f:
        f.write(f"# ncu --set full summary: {m['Kernel Name'][0] if 'Kernel Name' in m else src}\n\nsource report: {src} (scratch, not tracked)\n\n| metric | value | unit |\n|---|---:|---|\n")
        for k in KEY:
            if k in m:
                f.write(f"| {k} | {m[k][0]} | {m[k][1]} |\n")
        f.write("\n## instruction / stall-sample share per kernel phase (SASS segments delimited by barriers)\n\n| SASS range | instr share | sample share | ends at | top stall reasons |\n|---|---:|---:|---|---|\n")
        for s in segs:
            top = ", ".join(f"{a.replace('stall_', '')} {b}" for a, b in s["st"].most_common(3))
            f.write(f"| {s['start']}-{s['end']} | {s['inst'] / tot_i * 100:.1f}% | {s['samp'] / tot_s * 100:.1f}% | {s['marker']} | {top} |\n")
        f.write("\n## warp stall samples, whole kernel\n\n" + ", ".join(f"{a.replace('stall_', '')}: {b / tot_s * 100:.1f}%" for a, b in stalls.most_common(10)) + "\n")
    print(open(dst + ".md").read())
    if json_out:
        def num(k):
            v, u = m[k]
            v = float(v)
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}.get(u, 1)
        json.dump({"kernel": m.get("Kernel Name", ("?",))[0], "dram_bytes_per_launch": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"),
                   "duration_us": float(m["gpu__time_duration.sum"][0]), "source": dst + ".md", "lib_sha16": _lib_sha16(),
                   "note": "lib_sha16 = sha256 over the G2P2G kernel's sources and nvcc flags at capture time (bench.py: _lib_sha16); bench.py reports `traffic` only for the same build"},
                  open(json_out, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3], sys.argv[5] if len(sys.argv) > 5 and sys.argv[4] == "--json" else None)
