#!/bin/bash
# Final build, one B200: the whole GPU test-suite, smoke(), the bench line and the reference arm (plain runs).
set -u
OUT=gpurun_out
python -m pytest tests -x -q -m gpu > $OUT/r02_pytest_gpu.log 2>&1; tail -2 $OUT/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/r02_smoke.log 2>&1; tail -1 $OUT/r02_smoke.log
bash profiles/r02_run_bench.sh 1 | grep -v parity | cut -c1-250
python bench.py --impl reference --steps 3 --warmup 1 > $OUT/r02_bench_reference_arm.json 2> $OUT/r02_bench_reference_arm.err
echo short_done
