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
    dist.barrier()
    dist.destroy_process_group()
    if rank == 0:
        print("MGPU_OK world=%d" % world)


if __name__ == "__main__":
    main()
