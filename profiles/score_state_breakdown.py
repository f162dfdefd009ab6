#!/usr/bin/env python
"""Where the time between fit() and the first scored pool goes: the pieces of GraphGP._score_state
(feature scaling, prefix features, the two fp64 GEMMs per operator, hash tables), each bracketed by
a device sync.  Second pass per workload, so module loading and allocator growth are excluded.

    python profiles/score_state_breakdown.py
"""
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402
from nasbowl_b200.bayesopt import GraphGP  # noqa: E402
from nasbowl_b200.engine import Engine  # noqa: E402
from nasbowl_b200.kernels import WeisfilerLehman  # noqa: E402
from nasbowl_b200.pool_model import PoolModel  # noqa: E402


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


def main():
    eng = Engine.get()
    out = {}
    for name in ("cfg2", "cfg3", "cfg4"):
        wl = bench.WORKLOADS[name]
        from nasbowl_b200 import synth
        train = synth.generate_cells_host(wl["family"], seed=0, first_index=0, n_cells=wl["n_train"])
        y = torch.randn(wl["n_train"], generator=torch.Generator().manual_seed(0))
        for rep in range(2):
            gp = GraphGP(train, y, [WeisfilerLehman(h=2)], n_max=train.n_max)
            gp.fit(wl_subtree_candidates=wl["h_grid"])
            d = gp._dev
            n = gp.n
            fit, h, w = d['fits'][0], d['hs'][0], d['weights'][0]
            D = fit.k_end(h)
            T = fit.label_off[h + 1]
            t = [sync()]
            B = eng.empty((n, D), torch.float64)
            eng.scale_features(fit.phi, 0, D, d['s'], math.sqrt(w), B, D); t.append(sync())
            Linv, alpha = d['Linv'], d['alpha']; t.append(sync())
            mu_train = eng.empty((n,), torch.float64)
            eng.gemm(False, False, n, 1, n, 1.0, d['K'], n, alpha, 1, 0.0, mu_train, 1)
            mu_train.cpu(); t.append(sync())
            pm = PoolModel.from_fit(eng, fit, h); t.append(sync())
            Bp = eng.prefix_features(B, fit, h); t.append(sync())
            Wp = eng.empty((n, T + 1), torch.float64)
            eng.gemm(False, False, n, T + 1, n, 1.0, Linv, n, Bp, T + 1, 0.0, Wp, T + 1); t.append(sync())
            H = eng.empty((T + 1, T + 1), torch.float64)
            eng.gemm(True, False, T + 1, T + 1, n, 1.0, Wp, T + 1, Wp, T + 1, 0.0, H, T + 1); t.append(sync())
            w_mu_p = eng.empty((T + 1,), torch.float64)
            eng.gemm(True, False, T + 1, 1, n, 1.0, Bp, T + 1, alpha, 1, 0.0, w_mu_p, 1); t.append(sync())
            t0 = sync(); gp._score_state(); t1 = sync()
        names = ["scale_features", "Linv_access", "mu_train_gemv+argmax", "PoolModel.from_fit (hash tables, parents)",
                 "prefix_features", "gemm Linv*Bp [n x T x n]", "gemm Wp^T*Wp [T x T x n]", "gemv Bp^T*alpha"]
        out[name] = {"n": n, "h": h, "D": D, "T": T, "whole_score_state_ms": 1e3 * (t1 - t0)}
        for i, nm in enumerate(names):
            out[name][nm + "_ms"] = 1e3 * (t[i + 1] - t[i])
        out[name]["gemm_gflops"] = [2.0 * n * (T + 1) * n / (t[6] - t[5]) / 1e9, 2.0 * (T + 1) ** 2 * n / (t[7] - t[6]) / 1e9]
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
