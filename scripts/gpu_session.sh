#!/bin/bash
# one GPU session (run under gpurun): GPU test suite, per-variant G2P2G timings, one ncu capture of the current kernel
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s_tests.log 2>&1
( SPK_G2P2G_MODE=1 python scripts/g2p2g_perf.py mode1 20 2>&1 | tail -1
  python scripts/g2p2g_perf.py mode2 20 2>&1 | tail -1
  scripts/variants.sh run 20
  python scripts/g2p2g_perf.py mode2_again 20 2>&1 | tail -1 ) > gpurun_out/s_perf.log 2>&1
if [ "$1" != noprof ]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:g2p2g_cell --launch-skip 127 --launch-count 1 -f -o gpurun_out/s_prof \
  python scripts/g2p2g_perf.py ncu 3 120 > gpurun_out/s_prof.log 2>&1
fi
cat gpurun_out/s_tests.log gpurun_out/s_perf.log; tail -3 gpurun_out/s_prof.log
