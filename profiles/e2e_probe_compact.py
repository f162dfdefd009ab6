#!/usr/bin/env python
"""nbw_propose_host on 8-byte compact wire records: staged pipeline vs mapped reads, piece sweep.
    python profiles/e2e_probe_compact.py cfg2|cfg3 [pool]"""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
from nasbowl_b200.cells import CompactRecords
from nasbowl_b200.engine import Engine
eng = Engine.get()
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
w = bench.WORKLOADS[name]
M = int(sys.argv[2]) if len(sys.argv) > 2 else min(w["pool"], 2_000_000)
train, y = bench.make_training_set(w)
kernels, weights = bench.make_kernels(w)
gp = GraphGP(train, y, kernels, weights=weights)
bench.fit_gp(w, gp, train, y)
acq = GraphExpectedImprovement(gp)
dev = bench.generate_pool(eng, w, 0, M, 0, M)
parts = [b.to_compact() for b in bench.records_to_batches(w, dev.cpu().numpy(), 8)]
host = torch.from_numpy(np.concatenate([r for r, _ in parts], axis=1)).pin_memory()
model = acq._model(palettes=[p for _, p in parts])
print("workload", name, "pool", M, "wire bytes", host.shape[1])
ref = None
for mapped in (0, 1, 2):
    for pieces in (1, 2, 3, 4, 8):
        for _ in range(3):
            c, h = eng.propose_host(model, host, 5, True, chunks=pieces, mapped=mapped); eng.cands_read(c, 5)
        t = time.perf_counter()
        for _ in range(20):
            c, h = eng.propose_host(model, host, 5, True, chunks=pieces, mapped=mapped)
            v, i = eng.cands_read(c, 5)
        dt = (time.perf_counter() - t) / 20
        if ref is None: ref = (h.clone(), i.copy())
        assert torch.equal(ref[0], h) and np.array_equal(ref[1], i), "paths disagree"
        print("%s %2d: %.3f ms  (%.3g cand/s)" % (("staged chunks", "mapped reads, pieces", "mapped reads + writes, pieces")[mapped], pieces, dt * 1e3, M / dt))
