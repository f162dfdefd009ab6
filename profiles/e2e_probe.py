import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
from nasbowl_b200.engine import Engine
from nasbowl_b200.kernels import WeisfilerLehman
eng = Engine.get()
wl_name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
wl = bench.WORKLOADS[wl_name]
bench.FAMILY, bench.N_TRAIN, bench.H_GRID = wl["family"], wl["n_train"], wl["h_grid"]
train, y = bench.make_training_set()
gp = GraphGP(train, y, [WeisfilerLehman(h=2)]); gp.fit(wl_subtree_candidates=wl["h_grid"])
acq = GraphExpectedImprovement(gp)
M = wl["pool"]
dev = eng.generate_cells(wl["family"], 1234, 10000, M)
stride = dev.shape[1]
print("workload", wl_name, "pool", M, "record bytes", stride, "h*", gp.kernels[0].h)
host = torch.empty((M, stride), dtype=torch.uint8).pin_memory(); host.copy_(dev.cpu())
d2 = torch.empty_like(dev)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(20): d2.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("H2D %.0f MB pinned: %.3f ms = %.1f GB/s" % (M*stride/1e6, dt*1e3, M*stride/dt/1e9))
hs = torch.empty((M,), dtype=torch.float32).pin_memory(); ds = torch.empty((M,), dtype=torch.float32, device='cuda')
t=time.perf_counter()
for _ in range(20): hs.copy_(ds, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("D2H 4MB pinned: %.3f ms = %.1f GB/s" % (dt*1e3, 4e6/dt/1e9))
t=time.perf_counter()
for _ in range(20): x = hs.numpy().reshape(-1,1).copy()
dt=(time.perf_counter()-t)/20
print("numpy copy 4MB: %.3f ms" % (dt*1e3))
model = acq._model()
ref = None
for mapped in (False, True):
    for chunks in (1, 2, 4, 8, 16, 32):
        for _ in range(3):
            eng.score_host_pool(model, host, chunks=chunks, mapped=mapped); torch.cuda.synchronize()
        t=time.perf_counter()
        for _ in range(20):
            s, h = eng.score_host_pool(model, host, chunks=chunks, mapped=mapped)
            v, i = eng.topk(s, 5)
        dt=(time.perf_counter()-t)/20
        if ref is None: ref = h.clone()
        assert torch.equal(ref, h), "host scores differ between paths"
        print("%s %2d: %.3f ms  (%.3g cand/s)" % ("mapped pieces" if mapped else "staged chunks", chunks, dt*1e3, M/dt))
# the parts of one call, staged 8 chunks
for name, fn in (("score only (no topk, sync)", lambda: (eng.score_host_pool(model, host, chunks=8), torch.cuda.synchronize())),
                 ("mapped 1 piece only (sync)", lambda: (eng.score_host_pool(model, host, chunks=1, mapped=True), torch.cuda.synchronize())),
                 ("topk only", lambda: eng.topk(s, 5)),
                 ("pinned alloc 4MB", lambda: torch.empty((M,), dtype=torch.float32, pin_memory=True))):
    fn()
    t=time.perf_counter()
    for _ in range(20): fn()
    print("%s: %.3f ms" % (name, (time.perf_counter()-t)/20*1e3))

# the public call bench.py times as e2e
for _ in range(5): acq.propose_location(host, top_n=5)
torch.cuda.synchronize()
for reps in (20, 100, 200):
    t=time.perf_counter()
    for _ in range(reps): xs, eis, idx = acq.propose_location(host, top_n=5)
    torch.cuda.synchronize()
    print("propose_location(pinned host) x%d: %.3f ms" % (reps, (time.perf_counter()-t)/reps*1e3))
t=time.perf_counter()
for _ in range(100): m = acq._model()
print("acq._model(): %.4f ms" % ((time.perf_counter()-t)/100*1e3))
t=time.perf_counter()
for _ in range(100): host.is_pinned()
print("is_pinned(): %.4f ms" % ((time.perf_counter()-t)/100*1e3))
