#!/usr/bin/env python
"""A NAS-BOWL style Bayesian-optimisation loop on the GPU path, shaped like the reference's
nasbench_optimisation.py:102-216 (random pool -> propose_location -> evaluate -> reset_XY -> fit)
with a synthetic objective in place of the NAS-Bench look-up (which needs the multi-GB dataset).

    python examples/bo_loop.py [--iters 10] [--pool 200000] [--batch 5]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP  # noqa: E402
from nasbowl_b200.cells import NO_NODE, CellBatch  # noqa: E402
from nasbowl_b200.engine import Engine  # noqa: E402
from nasbowl_b200.kernels import WeisfilerLehman  # noqa: E402


def synthetic_objective(batch: CellBatch) -> np.ndarray:
    """Larger is better (the drivers feed -log(error)): rewards 3x3 convolutions on long paths."""
    lab = batch.labels
    conv3 = (lab == 3).sum(1)
    conv1 = (lab == 2).sum(1)
    pool = (lab == 4).sum(1)
    edges = batch.n_edges
    return (0.30 * conv3 + 0.12 * conv1 - 0.05 * pool + 0.04 * edges - 0.01 * (edges - 6) ** 2).astype(np.float64)


def run(iters=10, pool_size=200_000, batch_size=5, n_init=30, family=0, seed=0, verbose=True, check=None,
        strategy="random", n_best=10):
    """strategy "random": a fresh random pool per iteration (nasbench_optimisation.py --pool_strategy random);
    "mutate": half of the pool are BANANAS-style edits of the n_best cells observed so far, the rest
    random (generate_test_graphs.py:393-539), all generated on the device."""
    eng = Engine.get()
    x = CellBatch.from_records(eng.generate_cells(family, seed, 0, n_init).cpu().numpy(), 8)
    y = torch.tensor(synthetic_objective(x))
    gp = GraphGP(x, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(1, 2, 3), optimize_lik=True, max_lik=0.01)
    best = []
    for it in range(iters):
        t0 = time.perf_counter()
        first = 1_000_000 + it * pool_size
        if strategy == "mutate" and family == 0:
            top = np.argsort(-y.numpy())[:n_best]
            parents = eng.upload_cells(x[[int(i) for i in top]])
            n_mut = pool_size // 2
            pool = torch.cat([eng.mutate_cells(0, parents, seed + 2 + it, first, n_mut),
                              eng.generate_cells(family, seed + 1, first, pool_size - n_mut)])
        else:
            pool = eng.generate_cells(family, seed + 1, first, pool_size)     # stays on the device
        acq = GraphExpectedImprovement(gp, in_fill='best', augmented_ei=False)
        _, eis, idx = acq.propose_location(pool, top_n=batch_size)
        if check is not None:
            check(gp, x, y, pool, eis, idx)
        nxt = CellBatch.from_records(pool[torch.as_tensor(idx, device=pool.device)].cpu().numpy(), 8)
        x = CellBatch.concat([x, nxt])
        y = torch.cat([y, torch.tensor(synthetic_objective(nxt))])
        gp.reset_XY(x, y)
        gp.fit(wl_subtree_candidates=(1, 2, 3), optimize_lik=True, max_lik=0.01)
        best.append(float(y.max()))
        if verbose:
            print("iter %2d  n=%3d  h*=%d  best=%.4f  pool=%d  %.1f ms" %
                  (it, len(x), gp.kernels[0].h, best[-1], pool_size, 1e3 * (time.perf_counter() - t0)))
    return best, x, y


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--pool", type=int, default=200_000)
    ap.add_argument("--strategy", default="random", choices=["random", "mutate"])
    ap.add_argument("--batch", type=int, default=5)
    a = ap.parse_args()
    run(a.iters, a.pool, a.batch, strategy=a.strategy)
