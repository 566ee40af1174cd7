#!/bin/bash
# Everything that was written after the round-1 GPU budget was spent and has so far only run on the CUDA emulation / host-compiled
# sources of the CPU suite.  First GPU call of the next session:
#   gpurun --timeout 900 -- 'bash scripts/pending_gpu_checks.sh > gpurun_out/pending.log 2>&1'
# (about 3 minutes on one B200).  Each block is independent; the log ends with one PASS / FAIL line per block.
cd "$(dirname "$0")/.."
declare -A status
run() { local name=$1; shift; echo "=== $name: $*"; if "$@"; then status[$name]=PASS; else status[$name]=FAIL; fi; }
# 1. reference-harness device images of HeightField / Polyline colliders, dim3 and dim2 (layouts recovered from the PTX):
#    three-way lock-step
SPK_RUN_UNVERIFIED=1 run ref_colliders python -m pytest tests/test_reference_gpu.py tests/test_dim2_gpu.py -q -k unverified
# 2. fluids-volume recomputation stage (spk_fluids.cu), dim3 + dim2
run fluids_volume python -m pytest tests/test_volume_recomputation_gpu.py -q -x
# 3. device-side particle migration (spk_migrate.cu), incl. two ranks on the peer-memory path sharing cuda:0
run migration python -m pytest tests/test_zmigration_gpu.py -q -x
# 3b. dim2 TriMesh colliders (spk_grid2d.cu)
run trimesh2d python -m pytest tests/test_ztrimesh2d_gpu.py -q -x
# 3c. the remaining example scenes (rotated heightfields of sand_penetration3, l_panel2)
run scenes python -m pytest tests/test_zscenes_gpu.py -q -x
# 4. the C++ host example against the real library
run cpp_host bash -c 'g++ -std=c++17 -O2 -Iinclude examples/cpp_host/sand_column.cpp -Lsparkl_b200 -lsparkl_b200 -Wl,-rpath,$PWD/sparkl_b200 -o /tmp/sand_column && /tmp/sand_column 40 20 20 3'
# 5. the whole GPU suite, smoke and a bench line (unchanged kernels: SASS-identical to the last measured state)
run gpu_suite python -m pytest tests -q -x -m gpu
run smoke python -c 'import __graft_entry__ as g; g.smoke()'
run bench python bench.py
for k in "${!status[@]}"; do echo "RESULT $k ${status[$k]}"; done
