#!/usr/bin/env python
"""Host-pool entry points (nbw_propose_host): staged H2D / kernel / D2H pipeline vs mapped reads, piece sweep.
    python profiles/e2e_probe.py cfg2|cfg3|cfg4 [pool]"""
import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
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
stride = dev.shape[1]
print("workload", name, "pool", M, "record bytes", stride, "h*", gp.kernels[0].h)
host = torch.empty((M, stride), dtype=torch.uint8).pin_memory(); host.copy_(dev.cpu())
d2 = torch.empty_like(dev)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): d2.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 10
print("H2D %.0f MB pinned: %.3f ms = %.1f GB/s" % (M * stride / 1e6, dt * 1e3, M * stride / dt / 1e9))
model = acq._model()
sc = eng.empty((M,), torch.float32)
for _ in range(3): eng.score_pool_topk(model, dev, M, 5, scores=sc)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): c = eng.score_pool_topk(model, dev, M, 5, scores=sc)
torch.cuda.synchronize(); print("resident fused score+top5: %.3f ms" % ((time.perf_counter() - t) / 10 * 1e3))
ref = None
for mapped in (False, True):
    for pieces in (1, 2, 4, 8, 16, 32):
        for _ in range(2):
            c, h = eng.propose_host(model, host, 5, True, chunks=pieces, mapped=mapped); eng.cands_read(c, 5)
        t = time.perf_counter()
        for _ in range(10):
            c, h = eng.propose_host(model, host, 5, True, chunks=pieces, mapped=mapped)
            v, i = eng.cands_read(c, 5)
        dt = (time.perf_counter() - t) / 10
        if ref is None: ref = (h.clone(), i.copy())
        assert torch.equal(ref[0], h) and np.array_equal(ref[1], i), "paths disagree"
        print("%s %2d: %.3f ms  (%.3g cand/s)" % ("mapped pieces" if mapped else "staged chunks", pieces, dt * 1e3, M / dt))
for _ in range(3): acq.propose_location(host, top_n=5)
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(10): xs, eis, idx = acq.propose_location(host, top_n=5)
torch.cuda.synchronize()
print("propose_location(pinned host): %.3f ms" % ((time.perf_counter() - t) / 10 * 1e3))
