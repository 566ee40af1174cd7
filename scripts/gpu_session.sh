#!/bin/bash
# one GPU session (run under gpurun): [tests], per-variant G2P2G timings, [one ncu capture of the current kernel]
#   bash scripts/gpu_session.sh [tests] [prof] [env VAR=val ...]   (every "VAR=val" argument adds one extra perf run with that environment)
mkdir -p gpurun_out
: > gpurun_out/s_tests.log
if [[ " $* " == *" tests "* ]]; then ( timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > gpurun_out/s_tests.log 2>&1; fi
( python scripts/g2p2g_perf.py default 20 2>&1 | tail -1
  for a in "$@"; do if [[ "$a" == *=* ]]; then env $a python scripts/g2p2g_perf.py "$a" 20 2>&1 | tail -1; fi; done
  [ -d build/variants ] && scripts/variants.sh run 20
  python scripts/g2p2g_perf.py default_again 20 2>&1 | tail -1 ) > gpurun_out/s_perf.log 2>&1
if [[ " $* " == *" prof "* ]]; then
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:g2p2g_cell --launch-skip 127 --launch-count 1 -f -o gpurun_out/s_prof \
  python scripts/g2p2g_perf.py ncu 3 120 > gpurun_out/s_prof.log 2>&1
tail -3 gpurun_out/s_prof.log
fi
cat gpurun_out/s_tests.log gpurun_out/s_perf.log
