#!/usr/bin/env python
"""Pool-scoring throughput of the histogram-intersection (oa=True) kernel, the default base kernel of
the reference's NB101 / DARTS drivers: item form on the packed path (rank-1 chains + exception items) against the
linear kernel on the same cells and against the general G path of the first-generation kernel (NBW_SCORE_LEGACY=1)."""
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
for fam, n, M, hh in ((0, 500, 1_000_000, 2), (0, 500, 1_000_000, 4), (1, 150, 1_000_000, 2), (1, 150, 1_000_000, 5),
                      (2, 1000, 500_000, 2), (2, 1000, 500_000, 3)):
    for oa in (False, True):
        train = synth.generate_cells_host(fam, 0, 0, n)
        y = torch.randn(n, generator=torch.Generator().manual_seed(0))
        gp = GraphGP(train, y, [WeisfilerLehman(h=hh, oa=oa)])
        gp.fit(wl_subtree_candidates=(hh,), optimize_lik=True)
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
        rec = {"ms": ms, "cand_per_s": M / ms * 1e3, "n_features": model.n_features, "n_items": int(model.n_labels)}
        if oa:
            os.environ["NBW_SCORE_LEGACY"] = "1"
            model_l = acq._model()
            for _ in range(2):
                eng.score_pool(model_l, pool, M, scores=buf)
            torch.cuda.synchronize()
            a.record()
            for _ in range(5):
                eng.score_pool(model_l, pool, M, scores=buf)
            b.record()
            torch.cuda.synchronize()
            del os.environ["NBW_SCORE_LEGACY"]
            rec["general_path_ms"] = a.elapsed_time(b) / 5
            rec["vs_linear"] = ms / out["family%d_n%d_h%d_oa0" % (fam, n, hh)]["ms"]
        out["family%d_n%d_h%d_oa%d" % (fam, n, hh, int(oa))] = rec
print(json.dumps(out))
