#!/bin/bash
ncu --set full --import-source on --clock-control none -k regex:gp_factor_kernel -s 20 -c 1 -o gpurun_out/prof_factor_r02a -f python profiles/factor_bench.py > gpurun_out/r02_factor_prof.log 2>&1
echo rc=$?
