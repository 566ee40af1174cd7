#!/bin/bash
# dev aid: build A/B variants of libsparkl_b200.so that differ in the G2P2G translation unit only (extra nvcc defines) into
# build/variants/<name>.so; the other objects are the ones of the in-tree build.
#   scripts/variants.sh build name "-DSPK_G2P2G_MIN_CTAS=2" ...   |   scripts/variants.sh run [steps]   (on the GPU box: one line per variant)
set -e
cd "$(dirname "$0")/.."
mkdir -p build/variants
if [ "$1" = build ]; then
  name=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v "$@" -c sparkl_b200/csrc/spk_g2p2g.cu -o build/variants/${name}_g2p2g.o 2>&1 | grep -A2 "g2p2g_cell_kernel" | grep -E "stack|Used" || true
  objs=""
  for f in spk_api spk_sort spk_grid spk_dist spk_fluids spk_migrate; do objs="$objs sparkl_b200/csrc/$f.o"; done
  /usr/local/cuda/bin/nvcc -shared -Xlinker -Bsymbolic -o build/variants/$name.so build/variants/${name}_g2p2g.o $objs -gencode arch=compute_100a,code=sm_100a
  rm -f build/variants/${name}_g2p2g.o
  echo built build/variants/$name.so
else
  for so in build/variants/*.so; do
    SPK_LIB_PATH=$PWD/$so python scripts/g2p2g_perf.py "$(basename $so .so)" ${2:-20} 2>&1 | tail -1
  done
fi
