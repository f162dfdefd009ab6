#!/usr/bin/env python
"""Host-side cost of GraphExpectedImprovement.propose_location on pinned host records (cProfile, 200 calls)."""
import cProfile, pstats, sys, time, io
sys.path.insert(0, '.')
import torch, numpy as np
import bench
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
from nasbowl_b200.engine import Engine
eng = Engine.get()
w = bench.WORKLOADS["cfg2"]
train, y = bench.make_training_set(w)
kernels, weights = bench.make_kernels(w)
gp = GraphGP(train, y, kernels, weights=weights); bench.fit_gp(w, gp, train, y)
acq = GraphExpectedImprovement(gp)
M = 1_000_000
dev = bench.generate_pool(eng, w, 0, M, 0, M)
host = torch.empty((M, 16), dtype=torch.uint8).pin_memory(); host.copy_(dev.cpu())
for _ in range(20): acq.propose_location(host, top_n=5)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(100): acq.propose_location(host, top_n=5)
torch.cuda.synchronize(); print("propose_location: %.3f ms" % ((time.perf_counter() - t) * 10))
model = acq._model()
for pieces in (2, 3, 4):
    for _ in range(5): c, h = eng.propose_host(model, host, 5, True, chunks=pieces); eng.cands_read(c, 5)
    t = time.perf_counter()
    for _ in range(100): c, h = eng.propose_host(model, host, 5, True, chunks=pieces); eng.cands_read(c, 5)
    print("propose_host staged %d: %.3f ms" % (pieces, (time.perf_counter() - t) * 10))
from nasbowl_b200.cells import CellBatch, CompactRecords
rec, pal = CellBatch.from_records(dev.cpu().numpy(), 8).to_compact()
wire = CompactRecords(torch.from_numpy(rec).pin_memory(), [pal])
for _ in range(20): acq.propose_location(wire, top_n=5)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(100): acq.propose_location(wire, top_n=5)
torch.cuda.synchronize(); print("propose_location, 8-byte wire records: %.3f ms" % ((time.perf_counter() - t) * 10))
pr = cProfile.Profile(); pr.enable()
for _ in range(200): acq.propose_location(wire, top_n=5)
pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(30); print(s.getvalue()[:5000])
