"""Worker of tests/test_gpu_multi.py (launched under torchrun, one rank per GPU): shard invariance of
the scoring path — every rank scores its shard with the broadcast pool model; the merged top-n and the
concatenated scores must equal what one GPU computes on the whole pool."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from nasbowl_b200 import parallel, synth  # noqa: E402
from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP  # noqa: E402
from nasbowl_b200.engine import Engine  # noqa: E402
from nasbowl_b200.kernels import WeisfilerLehman  # noqa: E402


def main():
    rank, world, _ = parallel.init_distributed()
    eng = Engine.get()
    for family, h, oa in ((0, 2, False), (1, 3, False), (2, 1, True)):
        train = synth.generate_cells_host(family, 3, 0, 40)
        y = torch.randn(40, generator=torch.Generator().manual_seed(family))
        gp = GraphGP(train, y, [WeisfilerLehman(h=h, oa=oa)], n_max=train.n_max)
        pm = None
        if rank == 0:
            gp.fit(wl_subtree_candidates=(h,), optimize_lik=True)
            pm = gp._score_state()
            if oa:
                pm.descriptor("EI")            # materialise the general operators before the broadcast
        pm = parallel.broadcast_pool_model(pm, src=0)
        if rank != 0:
            gp.set_pool_model(pm)
        acq = GraphExpectedImprovement(gp)
        M = 100_003
        lo, hi = parallel.shard_range(M, rank, world)
        shard = eng.generate_cells(family, 99, 1000 + lo, hi - lo)
        vals, idx, scores = parallel.propose_location_sharded(acq, shard, index_base=lo, top_n=5)
        # single-GPU answer, computed redundantly by every rank on the whole pool
        whole = eng.generate_cells(family, 99, 1000, M)
        s_all = acq.score_pool(whole)[0]
        v1, i1 = eng.topk(s_all, 5, distinct=True)
        assert np.array_equal(idx, i1) and np.array_equal(vals, v1), (rank, idx, i1)
        assert torch.equal(scores, s_all[lo:hi]), "shard scores differ bitwise from the single-GPU scores"
        gathered = [None] * world
        dist.all_gather_object(gathered, [int(i) for i in idx])
        assert all(g == gathered[0] for g in gathered), "ranks disagree on the proposal"
    # a sum of kernels over two record slots (normal | reduce), pool steps in flight (PoolPipeline): the exchange of
    # pool i runs under the scoring of pool i + 1 and every selection still equals the single-GPU one
    from nasbowl_b200.cells import CellSlots
    n = 50
    tr = CellSlots([synth.generate_cells_host(2, 5, 0, n), synth.generate_cells_host(2, 6, 0, n)])
    y = torch.randn(n, generator=torch.Generator().manual_seed(9))
    gp = GraphGP(tr, y, [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)], weights=[0.5, 0.5])
    pm = None
    if rank == 0:
        gp.fit(wl_subtree_candidates=(), optimize_lik=False)
        pm = gp._score_state()
    pm = parallel.broadcast_pool_model(pm, src=0)
    if rank != 0:
        gp.set_pool_model(pm)
    assert len(pm.more) == 1 and pm.n_slots == 2
    acq = GraphExpectedImprovement(gp)
    model = acq._model()
    M = 60_001
    lo, hi = parallel.shard_range(M, rank, world)
    pipe = parallel.PoolPipeline(eng, 5, True, depth=2)
    pools, tickets, results = [], [], []
    for p in range(4):
        shard = torch.cat([eng.generate_cells(2, 100 + p, 1000 + lo, hi - lo), eng.generate_cells(2, 200 + p, 1000 + lo, hi - lo)], dim=1).contiguous()
        pools.append(shard)
        tickets.append(pipe.submit(model, shard, hi - lo, lo))
        if p >= 1:
            results.append(pipe.collect(tickets[p - 1]))
    results.append(pipe.collect(tickets[-1]))
    for p in range(4):
        whole = torch.cat([eng.generate_cells(2, 100 + p, 1000, M), eng.generate_cells(2, 200 + p, 1000, M)], dim=1).contiguous()
        s_all = acq.score_pool(whole)[0]
        v1, i1 = eng.topk(s_all, 5, distinct=True)
        assert np.array_equal(results[p][1], i1) and np.array_equal(results[p][0], v1), (rank, p, results[p], i1)
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK world=%d" % world)


if __name__ == "__main__":
    main()
