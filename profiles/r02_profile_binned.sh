#!/bin/bash
# ncu captures of the thread-per-cell scoring kernel with the size-binned walk (cfg2, cfg3) + the launch list of a
# cfg2 bench; every profiler pass follows a plain run of the same command that exited 0.
set -u
OUT=gpurun_out
for CFG in cfg2 cfg3; do
  python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/plain_r02d_$CFG.json 2>/dev/null || exit 1
  ncu --set full --clock-control none --import-source on -k regex:score_cell_kernel -s 4 -c 1 -f -o $OUT/prof_score_r02d_$CFG \
      python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/ncu_full_r02d_$CFG.log 2>&1
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/r02_launches_cfg2.csv \
    python bench.py --configs "" --steps 20 --warmup 5 --skip-cpu > $OUT/r02_ncu_launches_cfg2.log 2>&1
echo binned_profile_done
