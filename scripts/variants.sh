#!/bin/bash
# dev aid: build A/B variants of libsparkl_b200.so with extra nvcc defines into build/variants/<name>.so
#   scripts/variants.sh build name "-DSPK_G2P2G_MIN_CTAS=2" ...     only the G2P2G translation unit is recompiled (the other objects are the in-tree ones)
#   scripts/variants.sh buildall name "-DSPK_CELL_THREADS=256" ...  every translation unit is recompiled (defines that the host code sees too)
#   scripts/variants.sh run [steps]                                  (on the GPU box: one line per variant)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC"
if [ "$1" = build ] || [ "$1" = buildall ]; then
  mode=$1; name=$2; shift 2
  objs=""
  for f in spk_g2p2g spk_api spk_sort spk_grid spk_dist spk_fluids spk_migrate; do
    if [ $mode = buildall ] || [ $f = spk_g2p2g ]; then
      if [ $f = spk_g2p2g ]; then $NV -Xptxas -v "$@" -c sparkl_b200/csrc/$f.cu -o build/variants/${name}_$f.o 2>&1 | grep -A2 "g2p2g_cell_kernelI23DefaultParticleUpdaterTILj785EELb0" | grep -E "stack|Used" || true
      else $NV "$@" -c sparkl_b200/csrc/$f.cu -o build/variants/${name}_$f.o; fi
      objs="$objs build/variants/${name}_$f.o"
    else objs="$objs sparkl_b200/csrc/$f.o"; fi
  done
  /usr/local/cuda/bin/nvcc -shared -Xlinker -Bsymbolic -o build/variants/$name.so $objs -gencode arch=compute_100a,code=sm_100a
  rm -f build/variants/${name}_*.o
  echo built build/variants/$name.so
else
  for so in build/variants/*.so; do
    SPK_LIB_PATH=$PWD/$so python scripts/g2p2g_perf.py "$(basename $so .so)" ${2:-20} 2>&1 | tail -1
  done
fi
