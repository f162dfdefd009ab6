#!/usr/bin/env python
"""Latency of the fp64 GP factorisation pieces (CUDA events, median of 20): blocked Cholesky, the
triangular inverse, and the two flavours of nbw_gp_factor.

    python profiles/linalg_bench.py
"""
import ctypes as C
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def main():
    from nasbowl_b200.engine import Engine
    eng = Engine.get()
    lib, ctx, st = eng.lib, eng.ctx, eng.stream
    rng = np.random.RandomState(0)
    out = {}
    for n in (150, 500, 1000, 2000):
        A = rng.randn(n, n + 8)
        K = torch.tensor(A @ A.T / (n + 8) + 0.1 * np.eye(n), device="cuda")
        y = torch.tensor(rng.randn(n), device="cuda")
        L = eng.empty((n, n), torch.float64)
        Li = eng.empty((n, n), torch.float64)
        ld, info = C.c_double(0), C.c_int(0)

        def chol():
            eng._check(lib.nbw_chol_f64(ctx, eng._p(K), n, n, C.c_double(0.01), eng._p(L), n, C.byref(ld), C.byref(info), st))
        r = {"chol_ms": timed(chol)}
        r["trtri_ms"] = timed(lambda: eng._check(lib.nbw_trtri_lower_f64(ctx, eng._p(L), n, n, eng._p(Li), n, st)))
        r["gp_factor_full_ms"] = timed(lambda: eng.gp_factor(K, 0.01, y))
        r["gp_factor_likelihood_only_ms"] = timed(lambda: eng.gp_factor(K, 0.01, y, want_inverse=False))
        r["chol_gflops"] = n ** 3 / 3.0 / r["chol_ms"] / 1e6
        out[str(n)] = r
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
