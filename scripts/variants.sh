#!/bin/bash
# dev aid: build A/B variants of libsparkl_b200.so with extra nvcc defines into build/variants/<name>.so
#   scripts/variants.sh build name "-DSPK_G2P2G_MIN_CTAS=2" ...   |   scripts/variants.sh run   (on the GPU box: g2p2g ms per variant)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
if [ "$1" = build ]; then
  name=$2; shift 2
  objs=""
  for f in spk_api spk_sort spk_g2p2g spk_grid spk_dist; do
    /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC "$@" -c sparkl_b200/csrc/$f.cu -o build/variants/${name}_$f.o &
    objs="$objs build/variants/${name}_$f.o"
  done
  wait
  /usr/local/cuda/bin/nvcc -shared -o build/variants/$name.so $objs -gencode arch=compute_100a,code=sm_100a
  rm -f $objs
  echo built build/variants/$name.so
else
  for so in build/variants/*.so; do
    echo -n "$(basename $so .so): "
    SPK_LIB_PATH=$PWD/$so python bench.py --no-ref-gpu --no-cpu-baseline --e2e-frames 1 --steps 20 2>&1 | grep -o '"kernel_ms_per_launch": [0-9.]*\|"ms_per_step": [0-9.]*' | tr '\n' ' '
    echo
  done
fi
