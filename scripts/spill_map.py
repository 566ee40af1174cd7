"""dev aid: where ptxas put local-memory traffic (spills) in one kernel of a .cu — source line -> STL/LDL count, plus SASS size
  python scripts/spill_map.py sparkl_b200/csrc/spk_g2p2g.cu g2p2g_cell_kernel [extra nvcc flags...]"""
import collections, re, subprocess, sys
src, pat, extra = sys.argv[1], sys.argv[2], sys.argv[3:]
subprocess.check_call(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-cubin", "-o", "/tmp/_spill.cubin", src] + extra)
txt = subprocess.run(["nvdisasm", "-g", "/tmp/_spill.cubin"], capture_output=True, text=True).stdout
on, cur, cnt, n = False, None, collections.Counter(), 0
for line in txt.splitlines():
    if line.startswith("//--------------------- .text."):
        on = pat in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]+\*/", line):
        n += 1
        if re.search(r"\b(STL|LDL)", line):
            cnt[cur] += 1
print("SASS instructions", n, "local-memory instructions", sum(cnt.values()))
for k, v in sorted(cnt.items()):
    print(f"  {k[0]}:{k[1]}  {v}")
