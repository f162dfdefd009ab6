#!/bin/bash
python -m pytest tests/test_gpu_packed.py -m gpu -q -x -k "not cfg4 and not gram" 2>&1 | tail -3
for CFG in cfg2 cfg3 cfg4; do
python bench.py --workload $CFG --configs "" --steps 10 --warmup 3 --skip-cpu 2> gpurun_out/iter_$CFG.err | python -c "
import json,sys
r=json.loads(sys.stdin.read()); print('$CFG', 'value %.3e'%r['value'], 'kernel_ms %.4f'%r['roofline']['kernel_ms'], 'ms/step %.4f'%r['ms_per_step'], 'e2e %.3e'%r['e2e']['value'])"
tail -3 gpurun_out/iter_$CFG.err
done
for CFG in cfg2 cfg3 cfg4; do python profiles/e2e_probe.py $CFG > gpurun_out/r02_e2e_probe_$CFG.txt 2>&1; cat gpurun_out/r02_e2e_probe_$CFG.txt; done
