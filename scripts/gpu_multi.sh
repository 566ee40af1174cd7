#!/bin/bash
# multi-GPU session (run under gpurun --gpus N): bench lines for the given configs at N ranks  ->  gpurun_out/multi_<config>_n<N>.json
N=$1; shift
mkdir -p gpurun_out
for cfg in "$@"; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps 30 --warmup 5 \
     > gpurun_out/multi_${cfg}_n$N.json 2> gpurun_out/multi_${cfg}_n$N.err
  echo "$cfg n=$N rc=$?"; tail -c 2500 gpurun_out/multi_${cfg}_n$N.json; tail -5 gpurun_out/multi_${cfg}_n$N.err
done
