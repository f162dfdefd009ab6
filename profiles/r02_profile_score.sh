#!/bin/bash
# ncu capture of the packed scoring kernel on one config (run under gpurun, one GPU):
#   bash profiles/r02_profile_score.sh cfg2 r02a
CFG=${1:-cfg2}; TAG=${2:-r02a}
python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu > gpurun_out/plain_${TAG}_${CFG}.json 2> gpurun_out/plain_${TAG}_${CFG}.err || exit 1
ncu --set full --clock-control none --import-source on -k regex:score_cell_kernel -s 4 -c 2 \
    -o gpurun_out/prof_score_${TAG}_${CFG} -f python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu \
    > gpurun_out/ncu_full_${TAG}_${CFG}.log 2>&1
echo ncu_rc=$?
