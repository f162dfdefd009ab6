#!/usr/bin/env python
"""Where GraphGP.fit spends its time (wall clock with device syncs), per workload.  Each workload
is warmed once (fit + _score_state + one scored pool) so that one-off costs — lazy loading of each
kernel's module, torch's first fill kernel, allocator growth — are not charged to a stage."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from nasbowl_b200 import synth
from nasbowl_b200.bayesopt import GraphGP
from nasbowl_b200.engine import Engine
from nasbowl_b200.kernels import WeisfilerLehman

eng = Engine.get()
out = {}
for name, fam, n, cands in (("cfg2", 1, 150, tuple(range(6))), ("cfg3", 0, 500, (1, 2, 3)), ("cfg4", 2, 1000, (0, 1, 2, 3))):
    train = synth.generate_cells_host(fam, 0, 0, n)
    y = torch.randn(n, generator=torch.Generator().manual_seed(0))
    gp = GraphGP(train, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=cands)
    gp._score_state()
    gp.predict_diag(train)
    def sync(): torch.cuda.synchronize(); return time.perf_counter()
    t0 = sync(); cells = eng.upload_cells(train); t1 = sync()
    fit = eng.wl_fit(cells, train.n_max, max(cands), False); t2 = sync()
    k = gp.kernels[0]
    K = k.gram_device(fit, 2); t3 = sync()
    yd = gp._y_device(eng)
    r = eng.pd_inverse(K, 1e-3, yd); t4 = sync()
    Kn, s = gp.sum_kernels.combine_device([1.0], [K], normalize=True); t5 = sync()
    r = eng.gp_factor(Kn, 0.01, yd); t6 = sync()
    gp.reset_XY(train, y); k.change_kernel_params({'h': 2}); gp.likelihood = 1e-3
    t7 = sync(); gp.fit(wl_subtree_candidates=cands); t8 = sync(); gp._score_state(); t9 = sync()
    out[name] = {"n": n, "upload_ms": 1e3*(t1-t0), "wl_fit_ms": 1e3*(t2-t1), "gram_ms": 1e3*(t3-t2),
                 "one_pd_inverse_raw_ms": 1e3*(t4-t3), "normalise_ms": 1e3*(t5-t4), "one_gp_factor_ms": 1e3*(t6-t5),
                 "whole_fit_ms": 1e3*(t8-t7), "score_state_ms": 1e3*(t9-t8), "factorizations_per_fit": len(cands) + 21,
                 "dims": fit.dims}
print(json.dumps(out, indent=1))
