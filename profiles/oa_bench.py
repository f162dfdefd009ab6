#!/usr/bin/env python
"""Pool-scoring throughput of the histogram-intersection (oa=True) kernel, the default base kernel of
the reference's NB101 / DARTS drivers: general G path (thermometer features)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
from nasbowl_b200 import synth  # noqa: E402
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP  # noqa: E402
from nasbowl_b200.cells import CellBatch  # noqa: E402
from nasbowl_b200.engine import Engine  # noqa: E402
from nasbowl_b200.kernels import WeisfilerLehman  # noqa: E402

eng = Engine.get()
out = {}
for fam, n, M in ((0, 500, 1_000_000), (2, 1000, 500_000)):
    for oa in (False, True):
        train = synth.generate_cells_host(fam, 0, 0, n)
        y = torch.randn(n, generator=torch.Generator().manual_seed(0))
        gp = GraphGP(train, y, [WeisfilerLehman(h=2, oa=oa)])
        gp.fit(wl_subtree_candidates=(2,), optimize_lik=True)
        acq = GraphExpectedImprovement(gp)
        pool = eng.generate_cells(fam, 5, 10_000, M)
        model = acq._model()
        buf = eng.empty((M,), torch.float32)
        for _ in range(3):
            eng.score_pool(model, pool, M, scores=buf)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            eng.score_pool(model, pool, M, scores=buf)
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 10
        out["family%d_n%d_oa%d" % (fam, n, int(oa))] = {"ms": ms, "cand_per_s": M / ms * 1e3, "n_features": model.n_features}
print(json.dumps(out))
