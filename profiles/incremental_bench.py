#!/usr/bin/env python
"""Refit after appending 5 cells: incremental path (append-only dictionary + rank-k factor / operator updates) against
a from-scratch fit, single WL kernel (h = 2, NB101-shaped cells), wall clock of reset_XY + fit + pool operators."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from nasbowl_b200 import synth  # noqa: E402
from nasbowl_b200.bayesopt import GraphGP  # noqa: E402
from nasbowl_b200.kernels import WeisfilerLehman  # noqa: E402

out = {}
for n0 in (150, 500, 1000, 2000, 4000):
    big = synth.generate_cells_host(0, 3, 0, n0 + 40)
    y = torch.randn(n0 + 40, generator=torch.Generator().manual_seed(1)).double()
    rec = {}
    for mode in ("incremental", "scratch"):
        gp = GraphGP(big[:n0], y[:n0], [WeisfilerLehman(h=2)])
        gp.incremental = mode == "incremental"
        gp.incremental_min_n = 0          # force the append path at every size (the default policy starts at 768 cells)
        gp.fit(wl_subtree_candidates=(2,), optimize_lik=True)
        gp._score_state()
        ts = []
        for it in range(1, 8):
            nn = n0 + 5 * it
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            gp.reset_XY(big[:nn], y[:nn])
            gp.fit(wl_subtree_candidates=(2,), optimize_lik=True)
            gp._score_state()
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        rec[mode + "_ms"] = min(ts[1:])
        if mode == "incremental":
            rec["appended"] = bool(gp.last_fit_incremental)
    out[str(n0)] = rec
print(json.dumps(out))
