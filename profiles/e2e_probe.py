import sys, time, torch, numpy as np
sys.path.insert(0, '.')
import bench
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
from nasbowl_b200.engine import Engine
from nasbowl_b200.kernels import WeisfilerLehman
eng = Engine.get()
train, y = bench.make_training_set()
gp = GraphGP(train, y, [WeisfilerLehman(h=2)]); gp.fit(wl_subtree_candidates=tuple(range(6)))
acq = GraphExpectedImprovement(gp)
M = 1_000_000
dev = eng.generate_cells(1, 1234, 10000, M)
host = torch.empty((M, 16), dtype=torch.uint8).pin_memory(); host.copy_(dev.cpu())
d2 = torch.empty_like(dev)
torch.cuda.synchronize()
t=time.perf_counter()
for _ in range(20): d2.copy_(host, non_blocking=True)
torch.cuda.synchronize(); dt=(time.perf_counter()-t)/20
print("H2D 16MB pinned: %.3f ms = %.1f GB/s" % (dt*1e3, 16e6/dt/1e9))
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
for chunks in (1, 2, 4, 8, 16, 32):
    for _ in range(3):
        eng.score_host_pool(model, host, chunks=chunks); torch.cuda.synchronize()
    t=time.perf_counter()
    for _ in range(20):
        s, h = eng.score_host_pool(model, host, chunks=chunks)
        v, i = eng.topk(s, 5)
    dt=(time.perf_counter()-t)/20
    print("chunks %2d: %.3f ms  (%.3g cand/s)" % (chunks, dt*1e3, M/dt))
