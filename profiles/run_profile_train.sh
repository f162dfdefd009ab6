#!/bin/bash
# ncu captures of the training-side HBM kernels (relabel, radix scatter, histogram).
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python profiles/kernel_bench.py --only wl --cells 500000"
$CMD > $OUT/plain_train_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"wl_signature_kernel|hist_dense_kernel|radix_scatter_kernel|radix_count_kernel" -c 12 \
    -o $OUT/prof_train_$TAG $CMD > $OUT/ncu_train_$TAG.log 2>&1
echo train profile done
