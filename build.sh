#!/bin/bash
# Build libasr_b200.so (sm_100a only) in-tree.  Usage: ./build.sh [extra nvcc flags]
set -e
cd "$(dirname "$0")"
SRC=chinese_asr_b200/csrc
OUT=chinese_asr_b200/libasr_b200.so
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
$NVCC -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xcompiler -fPIC -shared -Xptxas -v "$@" \
    $SRC/api.cu $SRC/features.cu $SRC/gemm.cu $SRC/gemm_tc.cu $SRC/encoder.cu $SRC/encoder_tc.cu $SRC/encoder_tc3.cu $SRC/decoder.cu \
    -o $OUT 2> build.log || { cat build.log; exit 1; }
grep -E "error|warning" build.log | grep -v "ptxas info" | head -20 || true
echo "built $OUT"
