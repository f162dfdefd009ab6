#!/usr/bin/env python
"""Where GraphGP.fit spends its time: wall clock per stage with a device sync after each (so stages do not
overlap; the whole-fit figure is measured separately without the extra syncs).  Each workload is warmed once."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import bench
from nasbowl_b200.bayesopt import GraphGP
from nasbowl_b200.engine import Engine

eng = Engine.get()
out = {}

def sync():
    torch.cuda.synchronize()
    return time.perf_counter()

for name in ("cfg2", "cfg3", "cfg4"):
    w = bench.WORKLOADS[name]
    train, y = bench.make_training_set(w)
    kernels, weights = bench.make_kernels(w)
    gp = GraphGP(train, y, kernels, weights=weights)
    bench.fit_gp(w, gp, train, y); gp._score_state()
    n = len(train)
    r = {"n": n}
    # stage timings through monkey-patched engine entry points
    acc = {}
    def wrap(obj, attr, label):
        fn = getattr(obj, attr)
        def inner(*a, **k):
            t0 = sync(); res = fn(*a, **k); t1 = sync()
            acc[label] = acc.get(label, 0.0) + 1e3 * (t1 - t0)
            acc[label + "_calls"] = acc.get(label + "_calls", 0) + 1
            return res
        setattr(obj, attr, inner)
        return fn
    saved = [(eng, a, wrap(eng, a, a)) for a in ("upload_cells", "wl_fit", "level_grams_i32", "gp_grid", "gram", "gram_accumulate",
                                                 "axpby", "gram_normalize", "pd_inverse", "build_tables", "build_packed_tables",
                                                 "scale_features", "prefix_features", "gemm")]
    bench.reset_gp(w, gp, train, y)
    t0 = sync(); bench.fit_gp(w, gp, train, y); t1 = sync(); gp._score_state(); t2 = sync()
    r["fit_with_stage_syncs_ms"], r["score_state_with_stage_syncs_ms"] = 1e3 * (t1 - t0), 1e3 * (t2 - t1)
    r["stages"] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in acc.items()}
    for obj, a, fn in saved:
        setattr(obj, a, fn)
    ts = []
    for _ in range(5):
        bench.reset_gp(w, gp, train, y)
        t0 = sync(); bench.fit_gp(w, gp, train, y); t1 = sync(); gp._score_state(); t2 = sync()
        ts.append((1e3 * (t1 - t0), 1e3 * (t2 - t1)))
    r["fit_ms"], r["score_state_ms"] = min(t[0] for t in ts), min(t[1] for t in ts)
    out[name] = r
print(json.dumps(out, indent=1))
