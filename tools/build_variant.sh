#!/bin/bash
# Development: link a variant of the library in which the files named by SRCS (default: backward_h) are
# compiled with extra -D flags.   [SRCS="backward_h sample_advect"] tools/build_variant.sh <name> [-DFLAG ...]
#   ->  nvfi_b200/_variants/lib<name>.so      (NVFI_LIB_PATH=<that file> makes nvfi_b200._lib load it: A/B
# timing of variants on one GPU box with tools/probe_ab.py)
set -e
NAME=$1; shift
SRCS=${SRCS:-backward_h}
cd "$(dirname "$0")/.."
B=nvfi_b200/csrc/build
OUT=nvfi_b200/_variants
mkdir -p $OUT
OBJS=$(ls $B/*.o | grep -v debug_)
for SRC in $SRCS; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr \
       -Xptxas -v -I include -I nvfi_b200/csrc "$@" -c nvfi_b200/csrc/$SRC.cu -o $OUT/${SRC}_$NAME.o 2> $OUT/${SRC}_$NAME.ptxas &
  OBJS=$(echo "$OBJS" | grep -v "/$SRC.o"; echo "$OUT/${SRC}_$NAME.o")
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $OUT/lib$NAME.so $OBJS -cudart static -ldl
for SRC in $SRCS; do
  grep -A2 "k_advect_bwd_hE\|k_sample_advect_hE" $OUT/${SRC}_$NAME.ptxas | grep "spill\|registers" | head -4
done
echo $OUT/lib$NAME.so
