#!/bin/bash
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:gp_factor_kernel -c 40 --csv --log-file gpurun_out/r02_factor_launches.csv python profiles/factor_bench.py > gpurun_out/r02_factor_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r02_factor_launches.csv")) if len(r)>5]
hdr=rows[0]
ig=hdr.index("Grid Size"); iv=hdr.index("Metric Value"); 
for r in rows[1:]:
    print(r[ig], r[iv])
PY
