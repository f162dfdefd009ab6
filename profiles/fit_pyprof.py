"""cProfile of GraphGP.fit on cfg2 (host-side overhead between the launches)."""
import sys; sys.path.insert(0, '.')
import cProfile, pstats, time, torch
import bench
from nasbowl_b200.bayesopt import GraphGP
w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
train, y = bench.make_training_set(w)
kernels, weights = bench.make_kernels(w)
gp = GraphGP(train, y, kernels, weights=weights)
for i in range(3):
    bench.reset_gp(w, gp, train, y)
    bench.fit_gp(w, gp, train, y)
torch.cuda.synchronize()
pr = cProfile.Profile()
N = 20
t0 = time.perf_counter()
pr.enable()
for i in range(N):
    bench.reset_gp(w, gp, train, y)
    bench.fit_gp(w, gp, train, y)
pr.disable()
torch.cuda.synchronize()
print("mean fit ms (with profiler)", 1e3 * (time.perf_counter() - t0) / N)
st = pstats.Stats(pr)
st.sort_stats("cumulative").print_stats(45)
