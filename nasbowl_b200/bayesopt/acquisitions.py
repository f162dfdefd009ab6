"""EI / augmented EI / UCB over a candidate pool, fused with the GP posterior on the GPU
(reference bayesopt/acquisitions.py:9-138).

The reference scores one candidate per call (`eval`), re-running WL on train ∪ {candidate}
and on train ∪ train (for the incumbent) every time; here `propose_location` packs the whole
pool, scores it with ONE launch of the fused kernel (csrc/score_pool.cu) and selects the
top-n distinct values on the device (csrc/topk.cu).  Values agree with the reference's
per-candidate path because batched and per-candidate posteriors are mathematically identical
(SURVEY Q11).
"""
from abc import ABC

import numpy as np
import torch

from ..cells import CellBatch, CellSlots, CompactRecords


class BaseAcquisition(ABC):
    def __init__(self, gp):
        self.gp = gp
        self.iters = 0
        self.next_location = None
        self.next_acq_value = None

    def propose_location(self, *args):
        raise NotImplementedError

    def optimize(self):
        raise NotImplementedError

    def eval(self, x):
        raise NotImplementedError

    def __call__(self, *args, **kwargs):
        return self.eval(*args, **kwargs)


class GraphExpectedImprovement(BaseAcquisition):
    def __init__(self, surrogate_model, augmented_ei=False, xi: float = 0.0, in_fill: str = 'best',
                 strict: bool = True):
        """strict=False restores the reference's habit of returning -1 on any error inside eval
        (acquisitions.py:64-67); by default errors propagate."""
        super().__init__(surrogate_model)
        assert in_fill in ['best', 'posterior']
        self.in_fill = in_fill
        self.augmented_ei = augmented_ei
        self.xi = xi
        self.strict = strict

    # -- model configuration ------------------------------------------------------------
    def _model(self, palettes=None):
        if self.in_fill == 'max':                      # never true (assert above): SURVEY Q13
            inc = float(torch.max(self.gp.y_))
        else:
            inc = None                                 # posterior-mean incumbent (acquisitions.py:89-92)
        return self.gp.pool_model("AEI" if self.augmented_ei else "EI", xi=self.xi, incumbent=inc, palettes=palettes)

    def _get_incumbent(self):
        if self.in_fill == 'max':
            return torch.max(self.gp.y_)
        return self.gp.y_[self.gp._score_state().incumbent_idx]

    def score_pool(self, candidates, want64=False):
        """Acquisition values of a whole pool: (scores f32 device tensor [M], mu, var, score64)."""
        cells, M = self.gp.upload_pool(candidates)      # (may re-derive the device state for wider cells)
        eng = self.gp._dev['eng']
        return eng.score_pool(self._model(), cells, M, want64=want64)

    # -- reference API --------------------------------------------------------------------
    def eval(self, x, asscalar=False):
        single = not isinstance(x, (list, tuple, CellBatch, CellSlots)) or \
            (isinstance(x, tuple) and len(x) > 0 and hasattr(x[0], 'nodes'))     # one multi-slot example
        try:
            _, _, _, sc = self.score_pool([x] if single else x, want64=True)
        except Exception:
            if self.strict:
                raise
            return -1.
        ei = sc.cpu()
        if asscalar:
            return ei.numpy().item()
        return ei

    def propose_location(self, candidates, top_n=5, return_distinct=True):
        """top_n: return the top n candidates wrt the acquisition function (acquisitions.py:94-113).
        One fused launch scores the pool and keeps the local top-n (nbw_score_pool_topk / nbw_propose_host);
        sums of kernels, oa kernels on the general path or top_n > 8 fall back to score + tournament."""
        self.iters += 1
        self.gp._require_fit()
        if isinstance(candidates, CompactRecords):
            # 8-byte wire records (cells.CompactRecords): half the PCIe bytes of a pool that lives in host memory
            model = self._model(palettes=candidates.palettes)
            candidates = candidates.data
        else:
            model = self._model()
        eng = self.gp._dev['eng']
        eng.lib.nbw_range_push(b"propose_location")
        try:
            return self._propose(eng, model, candidates, top_n, return_distinct)
        finally:
            eng.lib.nbw_range_pop()

    def _host_policy(self, model, candidates):
        """(mapped, pieces) for a pool of packed records in host memory (profiles/r02_e2e_probe*.txt, per 1M
        candidates).  8-node records are scored one thread per cell with one coalesced vector load per record,
        which PCIe serves at full rate: a PINNED pool is read by the kernel itself, and the scores go back through
        the mapped alias of the pinned result block — no staging copy, no D2H copy (8-byte wire records:
        0.41 ms against 0.56 staged; 2M NB101 cells: 0.49 against 0.68).  Everything else (16-node records, whose
        lane-per-node kernel reads bytes; pageable pools) takes the staged H2D / kernel / D2H pipeline in pieces
        of ~12 MB (192 MB of DARTS pairs: 4.6 ms staged, 16-22 ms mapped)."""
        if model.n_max == 8 and not model.oa and candidates.is_pinned():
            return 2, 1
        if model.oa and model.n_max == 8 and model.h >= 5 and candidates.is_pinned():
            return 1, 1
        nbytes = candidates.shape[0] * candidates.shape[1]
        return 0, int(min(16, max(2, -(-nbytes // (12 << 20)))))

    def _propose(self, eng, model, candidates, top_n, return_distinct):
        host_scores = scores = None
        distinct = bool(return_distinct)
        fused = eng.fused_topk_ok(model, top_n)
        if isinstance(candidates, torch.Tensor) and not candidates.is_cuda:
            mapped, pieces = self._host_policy(model, candidates)
            if fused:
                cands, host_scores = eng.propose_host(model, candidates, top_n, distinct, chunks=pieces, mapped=mapped)
                scores = eng._host_state["scores"]
                vals, indices = eng.cands_read(cands, top_n)                             # syncs
            else:
                scores, host_scores = eng.score_host_pool(model, candidates, chunks=pieces, mapped=bool(mapped))
                vals, indices = eng.topk(scores, top_n, distinct=distinct)              # syncs
        else:
            cells, M = self.gp.upload_pool(candidates)
            if fused:
                scores = eng.empty((M,), torch.float32)
                cands = eng.score_pool_topk(model, cells, M, top_n, distinct, scores=scores)
                vals, indices = eng.cands_read(cands, top_n)                             # syncs
            else:
                scores, _, _, _ = eng.score_pool(model, cells, M)
                vals, indices = eng.topk(scores, top_n, distinct=distinct)              # syncs
        if return_distinct and len(indices) < top_n:
            # fewer than top_n distinct values: the reference falls back to plain top-k (:106-108)
            if host_scores is not None:
                eng.wait_host_scores()
                scores = host_scores.to(eng.device)       # (write-through pools leave no device copy)
            vals, indices = eng.topk(scores, top_n, distinct=False)
        if host_scores is not None:
            eis = host_scores.numpy().reshape(-1, 1)       # view of a pinned block owned by this result
        else:
            eis = scores.cpu().numpy().reshape(-1, 1)
        if isinstance(candidates, (CellBatch, CellSlots, torch.Tensor)):
            xs = tuple(int(i) for i in indices)
        else:
            xs = tuple([candidates[int(i)] for i in indices])
        return xs, eis, indices

    def optimize(self):
        raise ValueError("The kernel invoked does not have hyperparameters to optimse over!")


class GraphUpperConfidentBound(GraphExpectedImprovement):
    def __init__(self, surrogate_model, beta=None, strict: bool = True):
        super().__init__(surrogate_model, strict=strict)
        self.beta = beta

    def _model(self, palettes=None):
        if self.beta is None:                          # acquisitions.py:133-135
            self.beta = 3. * torch.sqrt(0.5 * torch.log(torch.tensor(2. * self.iters + 1.)))
        return self.gp.pool_model("UCB", beta=float(self.beta), palettes=palettes)
