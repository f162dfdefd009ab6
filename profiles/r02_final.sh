#!/bin/bash
# Round-2 evidence on the final build, one B200: every number under profiles/r02_* comes from this script (plain runs;
# the only profiler pass is the launch list at the end, and nothing printed under it is used as a bench value).
set -u
OUT=gpurun_out
python -m pytest tests -x -q -m gpu > $OUT/r02_pytest_gpu.log 2>&1; tail -2 $OUT/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke.log 2>&1; tail -2 $OUT/r02_smoke.log
python bench.py --steps 20 --warmup 5 > $OUT/r02_bench_1gpu.json 2> $OUT/r02_bench_1gpu.err || tail -5 $OUT/r02_bench_1gpu.err
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm.json 2> $OUT/r02_bench_reference_arm.err
python profiles/fit_breakdown.py > $OUT/r02_fit_breakdown.json 2> /dev/null
python profiles/factor_bench.py > $OUT/r02_factor_bench.json 2> /dev/null
python profiles/kernel_bench.py > $OUT/r02_kernel_bench.json 2> /dev/null
python profiles/oa_bench.py > $OUT/r02_oa_bench.json 2> /dev/null
python profiles/incremental_bench.py > $OUT/r02_incremental_bench.json 2> /dev/null
python profiles/linalg_bench.py > $OUT/r02_linalg_bench.json 2> /dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/r02_launches_cfg2.csv \
    python bench.py --configs "" --steps 20 --warmup 5 --skip-cpu > $OUT/r02_ncu_launches_cfg2.log 2>&1
echo final_done
