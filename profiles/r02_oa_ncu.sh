#!/bin/bash
# ncu capture of the histogram-intersection scoring kernel (item form, thread-per-cell): NB101-shaped cells, n = 500, h = 2
set -u
OUT=gpurun_out
cat > /tmp/oa_one.py <<'PY'
import sys; sys.path.insert(0, '.')
import torch
from nasbowl_b200 import synth
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
from nasbowl_b200.engine import Engine
from nasbowl_b200.kernels import WeisfilerLehman
eng = Engine.get()
fam, n, M, hh = int(sys.argv[1]), int(sys.argv[2]), 1_000_000, int(sys.argv[3])
train = synth.generate_cells_host(fam, 0, 0, n)
y = torch.randn(n, generator=torch.Generator().manual_seed(0))
gp = GraphGP(train, y, [WeisfilerLehman(h=hh, oa=True)])
gp.fit(wl_subtree_candidates=(hh,), optimize_lik=True)
acq = GraphExpectedImprovement(gp)
pool = eng.generate_cells(fam, 5, 10_000, M)
model = acq._model()
buf = eng.empty((M,), torch.float32)
for _ in range(4):
    eng.score_pool(model, pool, M, scores=buf)
torch.cuda.synchronize()
PY
python /tmp/oa_one.py 0 500 2 || exit 1
ncu --set full --clock-control none --import-source on -k regex:score_cell_kernel -s 2 -c 1 -f -o $OUT/prof_oa_r02_nb101_h2 \
    python /tmp/oa_one.py 0 500 2 > $OUT/ncu_oa_r02.log 2>&1
python /tmp/oa_one.py 1 150 5 || exit 1
ncu --set full --clock-control none --import-source on -k regex:score_cell_kernel -s 2 -c 1 -f -o $OUT/prof_oa_r02_nb201_h5 \
    python /tmp/oa_one.py 1 150 5 >> $OUT/ncu_oa_r02.log 2>&1
echo oa_profile_done
