#!/bin/bash
# Bounds check of the kernels WITHOUT a GPU (compute-sanitizer is closed on the GPU pool): the host-emulation build of
# the same kernel sources (tests/emu/cuda_emu.h, one OS thread per CUDA thread) compiled with AddressSanitizer and
# driven through the parity cases of every kernel family.  Device buffers are heap blocks there and the dynamic
# shared memory of a CTA is one heap block, so an access past either is reported.  usage: bash tools/asan_emu.sh
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${RDP_ASAN_DIR:-/tmp/rdp_asan}
mkdir -p "$OUT"
g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++17 -DRDP_EMU -I "$ROOT/tests/emu" -x c++ -fPIC -shared -pthread \
    -Wno-format-truncation -o "$OUT/librdp_emu.so" "$ROOT/fpga-fmcw-radar-processor_b200/csrc/rdp_api.cu"
ASAN_OPTIONS=detect_leaks=0:halt_on_error=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) RDP_ASAN_EMU="$OUT/librdp_emu.so" \
    python "$ROOT/tools/asan_emu_cases.py"
