#!/bin/bash
# Run on the GPU box (via gpurun): launch list + one full capture of the dominant kernel.
# Usage: bash profiles/run_profile.sh <tag>
set -u
TAG=${1:-r01}
OUT=gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --skip-cpu"
$CMD > $OUT/plain_$TAG.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv \
    --log-file $OUT/launches_$TAG.csv $CMD > $OUT/ncu_launches_$TAG.log 2>&1
$CMD > $OUT/plain2_$TAG.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:score_pool_kernel -s 2 -c 2 \
    -o $OUT/prof_score_$TAG $CMD > $OUT/ncu_full_$TAG.log 2>&1
echo profile done
