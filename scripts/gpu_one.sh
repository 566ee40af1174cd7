#!/bin/bash
# run one command on the GPU box and keep its output under gpurun_out/<name>.log:  scripts/gpu_one.sh name cmd...
name=$1; shift
mkdir -p gpurun_out
"$@" > gpurun_out/$name.log 2>&1
echo "rc=$?" >> gpurun_out/$name.log
tail -30 gpurun_out/$name.log
