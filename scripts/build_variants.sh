#!/bin/bash
# Build kernel-variant copies of libecwam_b200.so for A/B timing on the GPU box (developer tool).
# usage: scripts/build_variants.sh name1 "-DST_U=4 -DST_MINB=2" name2 "..." ...   -> build_variants/lib_<name>.so
set -e
cd "$(dirname "$0")/../ecwam_b200/csrc"
make -s
mkdir -p ../../build_variants
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
ARCH="-gencode arch=compute_100a,code=sm_100a"
FL="$ARCH -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I../../include"
while [ $# -gt 1 ]; do
  name=$1; flags=$2; shift 2
  ( $NVCC $FL $flags -Xptxas -v -c implsch.cu -o ../../build_variants/implsch_$name.o 2> ../../build_variants/implsch_$name.log
    $NVCC $FL $flags -Xptxas -v -c propag_fast.cu -o ../../build_variants/propag_fast_$name.o 2> ../../build_variants/propag_fast_$name.log
    $NVCC $ARCH -shared -o ../../build_variants/lib_$name.so api.o peaks.o propag.o outparam.o ../../build_variants/propag_fast_$name.o ../../build_variants/implsch_$name.o host_tables.o host_grid.o host_io.o -lnccl -lcudart -lgomp
    echo "$name: propags2_fast $(grep -A2 'propags2_fast_kernelILb0' ../../build_variants/propag_fast_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')"
    echo "$name: $(grep -A3 'k_sweepILi36ELi7ELb0' ../../build_variants/implsch_$name.log | grep -o 'Used [0-9]* registers\|[0-9]* bytes spill stores' | tr '\n' ' ')" ) &
done
wait
