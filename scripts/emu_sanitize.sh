#!/bin/bash
# Memory checks of the product's CUDA sources without a GPU (compute-sanitizer's memcheck / initcheck need a device): the CUDA
# emulation of tests/host/emu/ built with AddressSanitizer ("device memory" is the host heap, so an out-of-bounds global load / store of
# any kernel aborts the run), and a second pass with every cudaMalloc filled with a poison byte instead of zeros (reads of
# uninitialised device memory then change results).  ~12 minutes on 8 cores; not part of the default CPU suite.
cd "$(dirname "$0")/.."
set -e
echo "== AddressSanitizer"
SPK_EMU_ASAN=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    python -m pytest tests/test_emulated_cuda_cpu.py -x -q -k "not two_ranks and not cpp_host"
for b in 0xff 0x7b; do
    echo "== poisoned allocations ($b)"
    SPK_EMU_POISON=$b python -m pytest tests/test_emulated_cuda_cpu.py -x -q
done
echo "== reversed thread / CTA schedule (a read that misses a barrier or a __syncwarp passes in one order and fails in the other)"
SPK_EMU_SCHEDULE=reverse python -m pytest tests/test_emulated_cuda_cpu.py -x -q
