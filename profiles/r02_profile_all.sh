#!/bin/bash
# Round-2 ncu evidence (run under gpurun on ONE GPU; every ncu pass follows a plain run of the same command that
# exited 0): launch list of the full bench, then --set full captures of the dominant kernels.
set -u
OUT=gpurun_out
python bench.py --steps 3 --warmup 3 --skip-cpu > $OUT/plain_r02.json 2> $OUT/plain_r02.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_r02.csv \
    python bench.py --steps 3 --warmup 3 --skip-cpu > $OUT/ncu_launches_r02.log 2>&1
for CFG in cfg2 cfg3; do
  python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/plain_r02_$CFG.json 2>/dev/null || exit 1
  ncu --set full --clock-control none --import-source on -k regex:score_cell_kernel -s 4 -c 1 -f -o $OUT/prof_score_r02_$CFG \
      python bench.py --workload $CFG --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/ncu_full_r02_$CFG.log 2>&1
done
python bench.py --workload cfg4 --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/plain_r02_cfg4.json 2>/dev/null || exit 1
ncu --set full --clock-control none --import-source on -k regex:score_packed_kernel -s 4 -c 1 -f -o $OUT/prof_score_r02_cfg4 \
    python bench.py --workload cfg4 --configs "" --steps 3 --warmup 2 --skip-cpu > $OUT/ncu_full_r02_cfg4.log 2>&1
# Gram: symmetric schedule + u8 epilogue on a 32k-cell stress (same kernel as the 200k run, smaller capture)
python -c "
import sys; sys.path.insert(0,'.')
from nasbowl_b200.engine import Engine
eng=Engine.get(); c=eng.generate_cells(0,2024,0,32768); f=eng.wl_fit(c,8,3,False); r=eng.gram_stress(f,f.k_end(3),16384); print(r['ms'], r['checks'])" > $OUT/plain_r02_gram.log 2>&1 || exit 1
ncu --set full --clock-control none -k regex:gram_i8_pair_kernel -s 1 -c 1 -f -o $OUT/prof_gram_r02 python -c "
import sys; sys.path.insert(0,'.')
from nasbowl_b200.engine import Engine
eng=Engine.get(); c=eng.generate_cells(0,2024,0,32768); f=eng.wl_fit(c,8,3,False); r=eng.gram_stress(f,f.k_end(3),16384)" > $OUT/ncu_full_r02_gram.log 2>&1
echo profile_done
