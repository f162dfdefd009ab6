#!/usr/bin/env python
"""One-launch cluster factorisation (gp_factor.cu) vs the per-panel launches (linalg_f64.cu), and the batched
h grid (nbw_gp_grid) vs one factorisation per candidate.  CUDA-event medians.
    python profiles/factor_bench.py"""
import json, os, statistics, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nasbowl_b200 import synth
from nasbowl_b200.engine import Engine

def timed(fn, reps=15):
    for _ in range(3): fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)

eng = Engine.get()
out = {}
for fam, n, hmax in ((1, 150, 5), (0, 300, 3), (0, 500, 3), (2, 1000, 3)):
    train = synth.generate_cells_host(fam, 0, 0, n)
    cells = eng.upload_cells(train)
    fit = eng.wl_fit(cells, train.n_max, hmax, False)
    K = torch.zeros((n, n), dtype=torch.float64, device="cuda")
    eng.gram_accumulate(eng.gram(fit.phi, fit.phi, 0, fit.k_end(2)), 1.0, 0.0, K)
    eng.gram_normalize(K)
    y = torch.randn(n, dtype=torch.float64, device="cuda")
    r = {}
    for tag, mx in (("fused", "99999"), ("legacy", "0")):
        os.environ["NBW_FACTOR_FUSED_MAX"] = mx
        r["full_ms_" + tag] = timed(lambda: eng.gp_factor(K, 0.01, y, True))
        r["likelihood_only_ms_" + tag] = timed(lambda: eng.gp_factor(K, 0.01, y, False))
        a = eng.gp_factor(K, 0.01, y, True)
        r["stats_" + tag] = [a[3], a[4], a[5], a[6]]
    os.environ.pop("NBW_FACTOR_FUSED_MAX")
    G = eng.level_grams_i32(fit, hmax)
    cands = list(range(hmax + 1))
    r["grid_batched_ms"] = timed(lambda: eng.gp_grid(G, n, [c + 1 for c in cands], 1e-3, y))
    def serial():
        for c in cands:
            Kc = torch.empty((n, n), dtype=torch.float64, device="cuda")
            eng.gram_accumulate(eng.gram(fit.phi, fit.phi, 0, fit.k_end(c)), 1.0, 0.0, Kc)
            eng.gp_factor(Kc, 1e-3, y, False)
    os.environ["NBW_FACTOR_FUSED_MAX"] = "0"
    r["grid_serial_legacy_ms"] = timed(serial)
    os.environ.pop("NBW_FACTOR_FUSED_MAX")
    r["level_grams_ms"] = timed(lambda: eng.level_grams_i32(fit, hmax))
    out[str(n)] = r
print(json.dumps(out, indent=1))
