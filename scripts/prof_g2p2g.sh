#!/bin/bash
# one `ncu --set full` capture of the fused G2P2G kernel on the settled 4M sand scene -> gpurun_out/$1.ncu-rep
tag=${1:-prof_g2p2g}
ncu --set full --clock-control none --import-source on --kernel-name regex:g2p2g --launch-skip 125 --launch-count 1 -f -o gpurun_out/$tag \
  python bench.py --steps 2 --warmup 3 --settle 120 --no-ref-gpu --no-cpu-baseline --e2e-frames 1 > gpurun_out/$tag.log 2>&1
tail -c 300 gpurun_out/$tag.log
