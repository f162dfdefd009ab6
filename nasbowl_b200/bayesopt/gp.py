"""GraphGP — the reference's GP surrogate (bayesopt/gp.py:27-294) with every numerical step
on the GPU through libnbw's C-ABI.  Host logic (what is looped over, in which order, with which
clamp) mirrors the reference line by line so that h*, lambda, the incumbent and the selected
architectures agree; see SURVEY.md Appendix A for the quirks that are reproduced on purpose.

Differences a caller can observe:
  * results are computed in fp64 (the reference mixes fp64 and fp32, Q8) and returned as
    CPU float64 tensors;
  * WL features of the training set are computed ONCE per fit instead of once per grid
    candidate and once per Adam iteration (the Gram does not depend on lambda);
  * `predict_diag` / the acquisition classes score a whole pool with one fused kernel instead
    of re-fitting WL on train ∪ {candidate} per candidate.
"""
from __future__ import annotations

import logging
import math
import os
from copy import deepcopy
from typing import List, Optional, Sequence

import networkx as nx
import numpy as np
import torch

from .. import _lib
from ..cells import CellBatch, CellSlots, pack_networkx
from ..kernels import GraphKernels, SumKernel, WeisfilerLehman
from ..pool_model import PoolModel

_LOG_2PI = math.log(2.0 * math.pi)


def normalize_y(y: torch.Tensor):
    """bayesopt/utils.py:72-76 (torch.std is the unbiased estimator)."""
    y_mean = torch.mean(y) if isinstance(y, torch.Tensor) else np.mean(y)
    y_std = torch.std(y) if isinstance(y, torch.Tensor) else np.std(y)
    y = (y - y_mean) / y_std
    return y, y_mean, y_std


def _as_f64(y):
    """The GPU path evaluates the GP in fp64: normalise the targets in fp64 as well."""
    return y.detach().double() if isinstance(y, torch.Tensor) else torch.as_tensor(np.asarray(y, dtype=np.float64))


def unnormalize_y(y, y_mean, y_std, scale_std=False):
    """bayesopt/utils.py:79-85."""
    if not scale_std:
        return y * y_std + y_mean
    return y * y_std


def _nlml_from_stats(logdet: float, yKy: float, n: int) -> float:
    """-compute_log_marginal_likelihood (bayesopt/utils.py:102-117) with the reference's
    logDetK = -log|K| (utils.py:138): (yKy/2 - log|K|/2 + n/2 log 2pi) / n."""
    return (0.5 * yKy - 0.5 * logdet + 0.5 * n * _LOG_2PI) / n


class _Optim32:
    """Adam / SGD on a handful of fp32 scalars with torch.optim's default arithmetic (Adam: betas (0.9, 0.999),
    eps 1e-8, bias-corrected step lr / (1 - b1^t), denominator sqrt(v) / sqrt(1 - b2^t) + eps — the update of
    gp.py:182-207).  The reference optimises <= 3 + #levels numbers; a torch optimiser object costs ~0.1 ms per
    step on the host, twenty steps more than the whole device side of a fit."""

    def __init__(self, kind: str, lr: float):
        self.kind, self.lr, self.t = kind, np.float32(lr), 0
        self.m, self.v = {}, {}

    def step(self, name: str, p: torch.Tensor, grad: torch.Tensor):
        x = p.detach().numpy()
        g = grad.detach().numpy().astype(np.float32)
        if self.kind == 'sgd':
            x -= self.lr * g
            return
        b1, b2, eps = np.float32(0.9), np.float32(0.999), np.float32(1e-8)
        m = self.m.get(name, np.zeros_like(x))
        v = self.v.get(name, np.zeros_like(x))
        m = b1 * m + (np.float32(1) - b1) * g
        v = b2 * v + (np.float32(1) - b2) * g * g
        self.m[name], self.v[name] = m, v
        bc1, bc2 = np.float32(1.0 - 0.9 ** self.t), np.float32(1.0 - 0.999 ** self.t)
        x -= (self.lr / bc1) * (m / (np.sqrt(v) / np.sqrt(bc2) + eps))


class GraphGP:
    def __init__(self, train_x, train_y, kernels: list, vectorial_features: list = None, likelihood=1e-3,
                 weights=None, vector_theta_bounds: tuple = (1e-5, 0.1), graph_theta_bounds: tuple = (1e-1, 1.e1),
                 verbose=False, n_max=None, out_dtype=None):
        assert len(train_x) == train_y.shape[0], 'mismatch of length between train and test GP'
        if not isinstance(train_x, (CellBatch, CellSlots)):
            # (extension: an example may be a TUPLE of graphs, one per record slot — "normal | reduce" —
            # for sums whose kernels read different cells of the architecture, WeisfilerLehman(slot=i))
            assert all([isinstance(x, nx.Graph) or (isinstance(x, tuple) and all(isinstance(g, nx.Graph) for g in x))
                        for x in train_x]), \
                'each of the training example in train_x needs to be a networkX graph!'
        self.likelihood = likelihood
        self.kernels = kernels
        self.n_kernels = len(kernels)
        self.n_graph_kernels = len([i for i in kernels if isinstance(i, GraphKernels)])
        self.n_vector_kernels = self.n_kernels - self.n_graph_kernels
        if self.n_vector_kernels > 0 or vectorial_features:
            raise NotImplementedError("vectorial kernels / hand-crafted features are outside the GPU hot path")
        self.x = train_x[:]
        self.feature_d = None
        self.vectorial_feactures = vectorial_features
        self.x_features, self.x_features_min, self.x_features_max = [None] * 3
        self.n = len(self.x)
        self.y_ = deepcopy(train_y)
        self.y, self.y_mean, self.y_std = normalize_y(_as_f64(train_y))
        if weights is not None:
            self.fixed_weights = True
            assert len(weights) == len(kernels), \
                "the weights vector, if supplied, needs to have the same length as the number of kernels!"
            self.weights = weights if isinstance(weights, torch.Tensor) else torch.tensor(weights).flatten()
        else:
            self.fixed_weights = False
            self.weights = torch.tensor([1. / len(kernels)] * len(kernels), )
        self.sum_kernels = SumKernel(*kernels, weights=self.weights)
        self.vector_theta_bounds = vector_theta_bounds
        self.graph_theta_bounds = graph_theta_bounds
        self.verbose = verbose
        self.n_max = n_max         # force 16-node records when pools may hold larger cells than train
        # The GP is evaluated in fp64 and returns fp64 tensors; the reference returns fp32 (K, K_i, mu, cov:
        # combine_kernels.py:36, utils.py:140).  out_dtype=torch.float32 casts what predict() / K / K_i hand out,
        # for callers that concatenate them with fp32 tensors.
        self.out_dtype = out_dtype
        self.theta_vector = None
        self.layer_weights = None
        self.nlml = None
        self.nlml_grid = {}
        self.logDetK = None
        self._fitted = False
        self._dev = None           # device state of the last fit (not pickled)
        self._prev = None          # ... of the fit before reset_XY: the next fit may append to it
        self.pack_cache = True     # False: every fit() walks all of train_x again (see _pack_train)
        self._pack_memo = None
        self.incremental = True    # False: every fit() starts from scratch
        # measured (profiles/r02_incremental_bench.json): appending 5 cells beats a from-scratch fit from ~750 cells on
        # (4.8 vs 5.3 ms at 1000, 18 vs 31 ms at 4000); below that the whole fit is 1.5-3 ms and the bookkeeping loses
        self.incremental_min_n = 768
        self.last_fit_incremental = False

    # ------------------------------------------------------------------ device helpers
    def _engine(self):
        from ..engine import Engine
        return Engine.get()

    def _pack(self, gr) -> CellSlots:
        """list[nx.DiGraph] | list[tuple of nx.DiGraph] | CellBatch | CellSlots -> CellSlots (directed; every
        kernel symmetrises its own view if it is `undirected`, like the per-kernel conversion of
        kernels/weisfilerlehman.py:127-136)."""
        if isinstance(gr, CellSlots):
            return gr
        if isinstance(gr, CellBatch):
            return CellSlots([gr])
        wl = [k for k in self.kernels if isinstance(k, WeisfilerLehman)]
        labels = set(k.node_label for k in wl)
        if len(labels) > 1:
            raise NotImplementedError("kernels of one GP must read the same node attribute (got %s)" % sorted(labels))
        k0 = wl[0]
        gr = list(gr)
        if len(gr) and isinstance(gr[0], tuple):
            n_slots = len(gr[0])
            return CellSlots([pack_networkx([g[i] for g in gr], k0.vocab, node_label=k0.node_label)
                              for i in range(n_slots)])
        return CellSlots([pack_networkx(gr, k0.vocab, node_label=k0.node_label)])

    @staticmethod
    def _graph_sig(x):
        """(nodes, edges) of an example — or of every cell of a multi-cell example — read off the networkx
        dictionaries directly (number_of_edges() alone costs 3 us per graph, half a pack)."""
        if isinstance(x, tuple):
            return tuple((len(g._node), sum(map(len, g._adj.values()))) for g in x)
        return len(x._node), sum(map(len, x._adj.values()))

    def _pack_train(self) -> CellSlots:
        """self.x as records, re-using the records packed for the previous fit for the examples both lists share as
        a prefix.  The reference's BO loops append a batch to x per iteration and refit
        (nasbench_optimisation.py:209-213, run_darts.py:168-170); walking networkx objects costs ~7 us per cell
        (~11 for DARTS cells), which from a few hundred examples on is more than the device side of the fit.  An
        example counts as unchanged if it is the SAME object with the same node and edge counts (graphs edited in
        place between fits beyond that need `gp.pack_cache = False`)."""
        x = self.x
        if isinstance(x, (CellBatch, CellSlots)) or not getattr(self, 'pack_cache', True):
            return self._pack(x)
        x = list(x)
        wl = [k for k in self.kernels if isinstance(k, WeisfilerLehman)]
        if not wl or not len(x):
            return self._pack(x)
        memo = getattr(self, '_pack_memo', None)
        sig = self._graph_sig
        keep, sigs = 0, []
        if memo is not None and memo['vocab'] is wl[0].vocab and memo['label'] == wl[0].node_label:
            for a, b, sb in zip(x, memo['objs'], memo['sigs']):
                if a is not b:
                    break
                sa = sig(a)
                if sa != sb:
                    break
                sigs.append(sa)
            keep = len(sigs)
        if keep == 0:
            slots = self._pack(x)
        elif keep == len(x):
            # (a shorter list: like a fresh pack, 16-node records only if a cell needs them)
            slots = memo['slots'] if keep == len(memo['objs']) else memo['slots'][:keep].narrow()
        else:
            slots = CellSlots.concat([memo['slots'][:keep], self._pack(x[keep:])]).narrow()
        sigs.extend(sig(g) for g in x[keep:])
        self._pack_memo = dict(objs=x, sigs=sigs, slots=slots, vocab=wl[0].vocab, label=wl[0].node_label)
        return slots

    @staticmethod
    def _kernel_view(slots: CellSlots, k) -> CellBatch:
        """The cells kernel k reads: its record slot, symmetrised if the kernel is undirected."""
        slot = int(getattr(k, 'slot', 0))
        if slot >= slots.n_slots:
            raise ValueError("kernel reads record slot %d but the inputs have %d slot(s)" % (slot, slots.n_slots))
        b = slots.slots[slot]
        return b.to_undirected() if getattr(k, 'undirected', False) else b

    def _like_train(self, slots: CellSlots) -> CellSlots:
        """Bring a pool batch to the record width the GP was fitted with.  A pool with wider cells than the
        training set (the reference accepts any size) makes the GP re-derive its device state on 16-node
        records — same h, lambda and weights, no grid search, no optimisation."""
        n_max = self._dev['n_max']
        if slots.n_slots != self._dev['n_slots']:
            raise ValueError("the GP was fitted on %d-slot examples, the pool has %d" % (self._dev['n_slots'], slots.n_slots))
        if slots.n_max > n_max:
            self._rewiden(slots.n_max)
            n_max = self._dev['n_max']
        return slots.widen(n_max)

    def _rewiden(self, n_max: int):
        saved = (self.nlml, dict(self.nlml_grid), self.logDetK, self.weights.clone(), self.layer_weights)
        self.n_max = n_max
        fixed = self.fixed_weights
        self.fixed_weights = True                    # keep the fitted weights: nothing is trained here
        try:
            self.fit(iters=0, wl_subtree_candidates=(), optimize_lik=False)
        finally:
            self.fixed_weights = fixed
        self.nlml, self.nlml_grid, self.logDetK, self.weights, self.layer_weights = saved

    def _y_device(self, eng):
        return torch.as_tensor(np.asarray(self.y.detach().cpu().numpy(), dtype=np.float64)).to(eng.device)

    # ------------------------------------------------------------------ incremental refit
    def _incremental_base(self, slots: CellSlots, wl_subtree_candidates, optimize_wl_layer_weights):
        """The previous fit's device state if this fit can extend it: one plain linear WL kernel, the same record
        width and WL depth, and the old training cells as a prefix of the new ones."""
        prev = self._prev
        if prev is None or not self.incremental or os.environ.get("NBW_NO_INCREMENTAL"):
            return None
        ks = list(self.sum_kernels.kernels)
        if len(ks) != 1 or not isinstance(ks[0], WeisfilerLehman) or ks[0].oa or ks[0].undirected or optimize_wl_layer_weights:
            return None
        k = ks[0]
        if prev.get('records') is None or slots.n_slots != 1 or prev['n_max'] != slots.n_max or len(prev['fits']) != 1:
            return None
        old = prev['records']
        n_old = old.shape[0]
        if n_old >= len(slots) or prev['fits'][0].oa or n_old < self.incremental_min_n:
            return None
        h_fit = max([int(c) for c in wl_subtree_candidates] + [k.h])
        if prev['fits'][0].h != h_fit or k._level_weights(k.h) is not None:
            return None
        if not np.array_equal(slots.slots[0].labels[:n_old], old[:, :slots.n_max]) or \
                not np.array_equal(slots.slots[0].to_records()[:n_old], old):
            return None
        return prev

    def _append_factor(self, eng, prev, K, lik, y_dev):
        """Rank-k append of the Cholesky factor and its inverse: with K' = [[K, k12], [k12^T, K22]] (the old block
        is unchanged: raw Gram entries and the cosine normalisation of old cells do not depend on the new ones),
            L' = [[L, 0], [A^T, L22]],  A = L^-1 k12,  L22 L22^T = K22 + lik I - A^T A,
            L'^-1 = [[L^-1, 0], [-L22^-1 A^T L^-1, L22^-1]],
        O(n^2 k) instead of the O(n^3) of compute_pd_inverse (utils.py:120-140).  Returns None when the Schur
        complement is not positive definite (the caller factors from scratch with the reference's jitter retry)."""
        n, m = prev['fits'][0].n_cells, self.n
        kk = m - n
        Lo, Li = prev['L'], prev['Linv']
        A = eng.empty((n, kk), torch.float64)
        eng.gemm(False, False, n, kk, n, 1.0, Li, n, K[:n, n:], m, 0.0, A, kk)
        S = K[n:, n:].clone()
        eng.gemm(True, False, kk, kk, n, -1.0, A, kk, A, kk, 1.0, S, kk)
        try:
            L22, L22i, _, logdet_s, _, _, _ = eng.gp_factor(S, lik, y_dev[:kk].contiguous(), True)
        except _lib.NotPositiveDefinite:
            return None
        L = torch.zeros((m, m), dtype=torch.float64, device=eng.device)
        Linv = torch.zeros((m, m), dtype=torch.float64, device=eng.device)
        L[:n, :n].copy_(Lo)
        L[n:, :n].copy_(A.t())
        L[n:, n:].copy_(L22)
        Linv[:n, :n].copy_(Li)
        Linv[n:, n:].copy_(L22i)
        T1 = eng.empty((kk, n), torch.float64)
        eng.gemm(True, False, kk, n, n, 1.0, A, kk, Li, n, 0.0, T1, n)
        eng.gemm(False, False, kk, n, kk, -1.0, L22i, kk, T1, n, 0.0, Linv[n:, :n], m)
        z = eng.empty((m,), torch.float64)
        eng.gemm(False, False, m, 1, m, 1.0, Linv, m, y_dev, 1, 0.0, z, 1)
        alpha = eng.empty((m,), torch.float64)
        eng.gemm(True, False, m, 1, m, 1.0, Linv, m, z, 1, 0.0, alpha, 1)
        yKy = eng.empty((1,), torch.float64)
        eng.gemm(True, False, 1, 1, m, 1.0, z, 1, z, 1, 0.0, yKy, 1)
        return L, Linv, alpha, prev['logdet'] + logdet_s, float(yKy.item())

    # ------------------------------------------------------------------ grid search over h
    def _grid_search_wl_kernel(self, k: WeisfilerLehman, fit, subtree_candidates, y_dev, lik):
        """bayesopt/gp.py:493-538: for each candidate h, marginal likelihood of the
        UN-normalised Gram + lik*I; strict '<' keeps the smallest h on ties."""
        eng = self._engine()
        best_nlml, best_h = math.inf, None
        grid = {}
        # the reference builds the jitter as `lik * torch.eye(n)` in fp32 (utils.py:129): the value
        # that reaches the fp64 Gram is fp32(lik); near-singular K_raw(h=0) makes that visible
        lik = float(np.float32(lik))
        cands = [int(c) for c in subtree_candidates]
        batched = {}
        plain = all(k._level_weights(hc) is None for hc in cands)
        # (measured, profiles/r02_factor_bench.json: 6 candidates at n = 150 1.58 -> 0.24 ms, 4 at n = 500 2.9 -> 1.0 ms,
        # 4 at n = 1000 6.1 -> 3.8 ms; a single candidate beyond n = 256 is faster through the per-panel launches)
        if plain and 0 < len(cands) <= 8 and self.n <= int(os.environ.get("NBW_GRID_FUSED_MAX", "2048")) and \
                (len(cands) >= 2 or self.n <= 256):
            # all candidates in one launch, one cluster per depth, each summing its own level Grams (nbw_gp_grid)
            G = eng.level_grams_i32(fit, max(cands))
            for hc, (logdet, yKy, info) in zip(cands, eng.gp_grid(G, self.n, [hc + 1 for hc in cands], lik, y_dev)):
                if info == 0:
                    batched[hc] = (logdet, yKy)
        for hc in cands:
            if hc in batched:
                logdet, yKy = batched[hc]
            else:       # weighted levels, or a breakdown: compute_pd_inverse's retry with 10x / 100x the jitter
                K = k.gram_device(fit, hc)
                _, _, _, logdet, yKy, _, _ = eng.pd_inverse(K, float(lik), y_dev, want_inverse=False)
            nlml = _nlml_from_stats(logdet, yKy, self.n)
            grid[int(hc)] = nlml
            if nlml < best_nlml:
                best_nlml, best_h = nlml, hc
        k.change_kernel_params({'h': best_h})
        self.nlml_grid[id(k)] = grid
        return best_h

    # ------------------------------------------------------------------ fit
    def fit(self, iters=20, optimizer='adam', wl_subtree_candidates: tuple = tuple(range(5)),
            wl_lengthscales=tuple([np.e ** i for i in range(-2, 3)]), optimize_lik=True, max_lik=0.01,
            optimize_wl_layer_weights=False, optimizer_kwargs=None):
        if optimizer_kwargs is None:
            optimizer_kwargs = {'lr': 0.1}
        eng = self._engine()
        eng.lib.nbw_range_push(b"GraphGP.fit")
        try:
            return self._fit(eng, iters, optimizer, wl_subtree_candidates, optimize_lik, max_lik,
                             optimize_wl_layer_weights, optimizer_kwargs)
        finally:
            eng.lib.nbw_range_pop()

    def _fit(self, eng, iters, optimizer, wl_subtree_candidates, optimize_lik, max_lik, optimize_wl_layer_weights,
             optimizer_kwargs):
        slots = self._pack_train()
        if self.n_max is not None:
            slots = slots.widen(self.n_max)
        y_dev = self._y_device(eng)
        self.nlml_grid = {}
        prev = self._incremental_base(slots, wl_subtree_candidates, optimize_wl_layer_weights)
        self._prev = None
        self.last_fit_incremental = False

        # 1. WL features once per kernel, deep enough for every candidate; h by marginal likelihood.
        #    Every kernel reads ITS view of the examples (record slot, undirected or not): uploaded once per view.
        fits, views, view_cells = [], [], {}
        for k in self.sum_kernels.kernels:
            if not isinstance(k, WeisfilerLehman):
                logging.warning('Graph kernel optimisation for ' + str(k) + " not implemented yet.")
                raise NotImplementedError("only WeisfilerLehman kernels run on the GPU path")
            key = (int(getattr(k, 'slot', 0)), bool(k.undirected))
            if key not in view_cells:
                vb = self._kernel_view(slots, k)
                view_cells[key] = (vb, eng.upload_cells(vb))
            batch, cells = view_cells[key]
            views.append(key)
            cands = [int(c) for c in wl_subtree_candidates]
            h_fit = max(cands + [k.h])
            fit = None
            if prev is not None:
                fit = eng.wl_fit_append(prev['fits'][0], cells)
                if fit is None:
                    prev = None
                else:
                    k._fit, k._fit_batch, k._ref_cache, k._lin_cache = fit, batch, None, None
            if fit is None:
                fit = k.fit_device(cells, batch.n_max, h=h_fit, batch=batch)
            if len(cands):
                self._grid_search_wl_kernel(k, fit, cands, y_dev, self.likelihood)
            fits.append(fit)

        # 2. K = sum_i w_i K_raw,i, cosine-normalised (combine_kernels.py:36-56)
        weights = self.weights.clone()
        train_weights = (not self.fixed_weights) and len(self.kernels) > 1
        kernels = self.sum_kernels.kernels
        # trainable WL layer weights (gp.py:157-165): a vector of ones as long as the first kernel with
        # more than one level; kernels of another depth keep ones (weisfeiler_lehman.py:126-127)
        layer_weights = None
        if optimize_wl_layer_weights:
            for k in kernels:
                if k.h + 1 > 1:
                    layer_weights = torch.ones(k.h + 1)
                    break
        shares = [layer_weights is not None and k.h + 1 == layer_weights.shape[0] for k in kernels]
        levels = None
        if train_weights or layer_weights is not None:
            levels = [k.level_grams_device(f) for k, f in zip(kernels, fits)]

        own_lw = [k._level_weights(k.h) for k in kernels]      # fixed, user-given layer weights (or None = ones)

        def level_weight(ki, l):
            if shares[ki]:
                return float(layer_weights.detach()[l])
            return 1.0 if own_lw[ki] is None else float(own_lw[ki][l])

        def build_K():
            """Raw Gram per kernel from its level Grams and the current layer weights, then the
            weighted, cosine-normalised sum."""
            if layer_weights is None:
                gr = [k.gram_device(f) for k, f in zip(kernels, fits)]
            else:
                gr = []
                for i, k in enumerate(kernels):
                    R = eng.empty((self.n, self.n), torch.float64)
                    for l, Kl in enumerate(levels[i]):
                        eng.axpby(Kl, level_weight(i, l), 0.0 if l == 0 else 1.0, R)
                    gr.append(R)
            return self.sum_kernels.combine_device(weights, gr, normalize=True)
        K, inv_sqrt_diag = build_K()

        # 3. noise (and, for several kernels, their weights; optionally the layer weights): Adam / SGD
        #    with the reference's clamps (gp.py:141-210).  Gradients are closed-form device reductions
        #    instead of autograd.
        likelihood = torch.tensor(self.likelihood, )
        nlml = None
        nlml_pending = False
        factor_cache = {}

        def factor(lam: float):
            """K does not depend on lambda and, with the reference's sign quirk, lambda sits at
            max_lik from the second iteration on: identical lambdas are factored once."""
            if lam not in factor_cache:
                factor_cache[lam] = eng.pd_inverse(K, lam, y_dev)
            return factor_cache[lam]
        if optimize_lik or levels is not None:
            optim_vars = []
            if train_weights:
                weights.requires_grad_(True)
                optim_vars.append(weights)
            if optimize_lik:
                likelihood.requires_grad_(True)
                optim_vars.append(likelihood)
            if layer_weights is not None:
                layer_weights.requires_grad_(True)
                optim_vars.append(layer_weights)
            assert optimizer.lower() in ['adam', 'sgd']
            light = set(optimizer_kwargs) <= {'lr'}          # default arithmetic: the fp32 numpy twin of torch.optim
            if light:
                optim = _Optim32(optimizer.lower(), optimizer_kwargs.get('lr', 1e-3 if optimizer.lower() == 'adam' else None))
                for v in optim_vars:
                    v.requires_grad_(False)
            else:
                optim = (torch.optim.Adam if optimizer.lower() == 'adam' else torch.optim.SGD)(optim_vars,
                                                                                             **optimizer_kwargs)
            # With the reference's sign quirk the noise gradient is negative for every K (Q9:
            # -(|alpha|^2 + tr K_l^-1) / 2n, and tr K_l^-1 >= n / (n + lambda) for a unit-diagonal K), so the
            # first Adam step is +lr (1 - eps/|g|) = +lr in fp32 and the clamp lands on max_lik: when only the
            # noise is trained and lambda_0 + lr >= max_lik, the factorisation at lambda_0 cannot influence
            # anything the fit returns (nlml is the value of the LAST iteration) and is skipped.
            skip_first = (light and optimizer.lower() == 'adam' and optimize_lik and levels is None and iters >= 2
                          and not self.verbose and float(likelihood) + float(optimizer_kwargs.get('lr', 1e-3)) * 0.999 >= max_lik
                          and max_lik >= 1e-5)
            # (when the loop only ever visits max_lik and the previous fit can be extended, the factorisation is
            #  left to the append below: nlml of the last iteration is that of the final factor)
            defer = skip_first and prev is not None and prev['hs'] == [k.h for k in kernels] and \
                prev['lik'] == float(np.float32(max_lik))
            for i in range(iters):
                if defer:
                    likelihood.fill_(max_lik)
                    nlml_pending = True
                    break
                if not light:
                    optim.zero_grad()
                if skip_first and i == 0:
                    likelihood.fill_(max_lik)
                    optim.t = 1
                    optim.m['lik'], optim.v['lik'] = np.float32(-0.1) * np.ones((), np.float32), np.float32(1e-3) * np.ones((), np.float32)
                    continue
                if levels is not None:
                    # K moves with the weights: rebuild and refactor every iteration.  The K of the LAST
                    # iteration (weights before the final step) is the one the reference keeps (gp.py:186-215).
                    with torch.no_grad():
                        K, inv_sqrt_diag = build_K()
                    factor_cache.clear()
                lam_i = float(likelihood.item())
                _, Linv_i, alpha_i, logdet, yKy, trKinv, alpha_sq = factor(lam_i)
                nlml = _nlml_from_stats(logdet, yKy, self.n)
                if levels is not None:
                    # the gradient form is linear in the raw Gram: one value per (kernel, level) Gram
                    flat = [Kl for lev in levels for Kl in lev]
                    g = eng.kernel_weight_grad(Linv_i, alpha_i, K, inv_sqrt_diag, flat)
                    gw = torch.zeros(len(kernels))
                    glw = torch.zeros(layer_weights.shape[0]) if layer_weights is not None else None
                    pos = 0
                    for ki, lev in enumerate(levels):
                        for l in range(len(lev)):
                            gw[ki] += level_weight(ki, l) * g[pos]
                            if shares[ki]:
                                glw[l] += float(weights.detach()[ki]) * g[pos]
                            pos += 1
                    if train_weights:
                        weights.grad = gw.to(weights.dtype)
                    if layer_weights is not None:
                        layer_weights.grad = glw.to(layer_weights.dtype)
                if optimize_lik:
                    # d nlml / d lambda = -(|K_l^-1 y|^2 + tr K_l^-1) / (2n): always negative (Q9)
                    likelihood.grad = torch.tensor(-(0.5 * alpha_sq + 0.5 * trKinv) / self.n, dtype=likelihood.dtype)
                if self.verbose and i % 10 == 0:
                    print('Iteration:', i, "/", iters, 'Negative log-marginal likelihood:', nlml, weights, likelihood)
                if light:
                    optim.t += 1
                    if train_weights:
                        optim.step('w', weights, weights.grad)
                    if optimize_lik:
                        optim.step('lik', likelihood, likelihood.grad)
                    if layer_weights is not None:
                        optim.step('lw', layer_weights, layer_weights.grad)
                else:
                    optim.step()
                with torch.no_grad():
                    if train_weights:
                        weights.clamp_(0., 1.)
                    if optimize_lik:
                        likelihood.clamp_(1e-5, max_lik)
                    if layer_weights is not None:
                        layer_weights.clamp_(0., 1.)
                # Only the noise is trained and it sits on the clamp: its gradient is negative for every K (Q9), so
                # Adam's first moment never changes sign, every further step pushes against the same clamp and every
                # further iteration re-evaluates the same (memoised) factorisation — the loop's result is final.
                if (light and optimizer.lower() == 'adam' and optimize_lik and levels is None and not self.verbose
                        and lam_i == float(np.float32(max_lik)) and float(likelihood.item()) == lam_i
                        and float(likelihood.grad) < 0.0):
                    break
            weights = weights.detach()
        if layer_weights is not None:
            layer_weights = layer_weights.detach()
            for ki, k in enumerate(kernels):       # predict() runs with the final layer weights (gp.py:218,269)
                if shares[ki]:
                    k.change_kernel_params({'layer_weights': layer_weights.clone()})
        self.layer_weights = layer_weights
        lik = float(likelihood.item())
        appended = None
        if prev is not None and lik not in factor_cache and prev['hs'] == [k.h for k in kernels] and prev['lik'] == lik:
            appended = self._append_factor(eng, prev, K, lik, y_dev)
        if appended is not None:
            L, Linv, alpha, logdet, yKy = appended
            nlml = _nlml_from_stats(logdet, yKy, self.n) if nlml_pending else nlml
            self.last_fit_incremental = True
        else:
            L, Linv, alpha, logdet, yKy, _, _ = factor(lik)
            if nlml_pending:
                nlml = _nlml_from_stats(logdet, yKy, self.n)
        # predict() normalises the joint Gram with the STORED weights, which are one optimiser step ahead of
        # the K behind K_i (gp.py:186-215 keeps the K of the last loop iteration): the pool path needs the
        # training diagonal under those final weights
        s_pool = inv_sqrt_diag
        if levels is not None and iters > 0:
            _, s_pool = self.sum_kernels.combine_device(weights, [k.gram_device(f) for k, f in zip(kernels, fits)],
                                                        normalize=True)

        # 4. state (gp.py:212-219)
        self.weights = weights.clone() / torch.sum(weights)
        self.likelihood = lik
        self.logDetK = torch.tensor(-logdet, dtype=torch.float64)
        self.nlml = torch.tensor(nlml, dtype=torch.float64) if nlml is not None else None
        self.sum_kernels.weights = weights.clone()
        self._fitted = True
        self._dev = dict(eng=eng, fits=fits, hs=[k.h for k in self.sum_kernels.kernels], K=K, L=L, Linv=Linv, logdet=logdet,
                         lik=lik, records=slots.slots[0].to_records() if slots.n_slots == 1 else None,
                         prev_score=(prev.get('score') if prev is not None and self.last_fit_incremental else None),
                         prev_n=(prev['fits'][0].n_cells if prev is not None else 0),
                         alpha=alpha, s=inv_sqrt_diag, s_pool=s_pool, y=y_dev, n_max=slots.n_max, n_slots=slots.n_slots,
                         views=views, view_cells=view_cells, slots=slots,
                         weights=[float(w) for w in weights], score=None, K_host=None, Ki_host=None)
        if self.verbose:
            print('Optimisation summary: ')
            print('Optimal NLML: ', nlml)
            print('Optimal h: ', [k.h for k in self.kernels])
            print('Weights: ', self.weights)
            print('Lik:', self.likelihood)

    # ------------------------------------------------------------------ reference attributes
    @property
    def K(self):
        if not self._fitted or self._dev is None:
            return None
        if self._dev['K_host'] is None:
            self._dev['K_host'] = self._out(self._dev['K'].cpu())
        return self._dev['K_host']

    @property
    def K_i(self):
        """(K + lambda I)^-1 = Linv^T Linv (torch.cholesky_inverse in utils.py:139)."""
        if not self._fitted or self._dev is None:
            return None
        if self._dev['Ki_host'] is None:
            eng, Linv, n = self._dev['eng'], self._dev['Linv'], self.n
            Ki = eng.empty((n, n), torch.float64)
            eng.gemm(True, False, n, n, n, 1.0, Linv, n, Linv, n, 0.0, Ki, n)
            self._dev['Ki_host'] = self._out(Ki.cpu())
        return self._dev['Ki_host']

    def _out(self, t):
        dt = getattr(self, 'out_dtype', None)
        return t if dt is None else t.to(dt)

    def _require_fit(self):
        if not self._fitted or self._dev is None:
            raise ValueError("Inverse of Gram matrix is not instantiated. Please call the optimize function to "
                             "fit on the training data first!")

    # ------------------------------------------------------------------ predict (reference path)
    def predict(self, X_s, preserve_comp_graph=False):
        """Kriging predictions (mu [M], cov [M, M]) exactly as bayesopt/gp.py:240-283: joint WL
        on train ∪ test, full M x M covariance with the element-wise clamp."""
        if not isinstance(X_s, (list, CellBatch, CellSlots)):
            X_s = [X_s]
        self._require_fit()
        n, M = self.n, len(X_s)
        test = self._like_train(self._pack(X_s))
        d = self._dev                            # (a wider pool re-derives the device state)
        eng = d['eng']
        grams, joint = [], {}
        for k, h, key in zip(self.sum_kernels.kernels, d['hs'], d['views']):
            if key not in joint:
                tb = self._kernel_view(test, k)
                joint[key] = torch.cat([d['view_cells'][key][1], eng.upload_cells(tb)], dim=0).contiguous()
            # The reference leaves the kernel object fitted on train ∪ test (Q19), which silently corrupts its
            # own interpretability outputs afterwards; here the kernel keeps its TRAINING fit (k._fit).
            fit = eng.wl_fit(joint[key], d['n_max'], h, k.oa)
            grams.append(k.gram_device(fit, h))
        K_full, _ = self.sum_kernels.combine_device(d['weights'], grams, normalize=True)
        N = n + M
        K_s = K_full[:n, n:]                     # views: explicit leading dimension N below
        K_ss = K_full[n:, n:]
        W = eng.empty((n, M), torch.float64)     # W = L^-1 K_s
        eng.gemm(False, False, n, M, n, 1.0, d['Linv'], n, K_s, N, 0.0, W, M)
        Q = eng.empty((M, M), torch.float64)     # K_s^T K_i K_s = W^T W
        eng.gemm(True, False, M, M, n, 1.0, W, M, W, M, 0.0, Q, M)
        mu = eng.empty((M,), torch.float64)      # K_s^T K_i y
        eng.gemm(True, False, M, 1, n, 1.0, K_s, N, d['alpha'], 1, 0.0, mu, 1)
        cov = eng.posterior_cov_finish(K_ss, N, Q, M, self.likelihood, float(self.y_std))
        mu_s = unnormalize_y(mu.cpu(), float(self.y_mean), float(self.y_std))
        return self._out(mu_s), self._out(cov.cpu())

    # ------------------------------------------------------------------ fused pool path
    def _score_state(self):
        """Feature-space posterior operators of the fitted surrogate (csrc/score_packed.cu, score_pool.cu).

        With omega_a = kernel weight x WL layer weight of feature column a (combine_kernels.py:36-40,
        weisfeiler_lehman.py:316-327) and s_i = 1/sqrt(K_comb,ii) (combine_kernels.py:54-56):
            B = diag(s) [Phi_1 | Phi_2 | ...] diag(omega),   k_s = B phi_c / sqrt(k_cc),
            k_cc = sum_a omega_a phi_c,a^2 over ALL labels of the candidate,
        so mu ~ phi_c . (B^T alpha) and k_s^T K_lam^-1 k_s = phi_c^T (B^T K_lam^-1 B) phi_c / k_cc.  A sum
        of kernels is the same computation over the concatenated feature space; its kernels may read
        different records of the candidate (WeisfilerLehman(slot=i))."""
        self._require_fit()
        d = self._dev
        if d.get('score') is not None:
            return d['score']
        eng, n = d['eng'], self.n
        with eng.range("GraphGP.score_state"):
            pm = self._build_score_state(eng, d, n)
        d['score'] = pm
        return pm

    def _build_score_state(self, eng, d, n):
        kernels = list(self.sum_kernels.kernels)
        fits, hs, weights = d['fits'], d['hs'], d['weights']
        Linv, alpha = d['Linv'], d['alpha']
        lws = [k._level_weights(h) for k, h in zip(kernels, hs)]           # None = ones
        plain = len(kernels) == 1 and lws[0] is None
        any_oa = any(f.oa for f in fits)
        if not plain and any_oa:
            raise NotImplementedError("the fused pool path scores sums / layer-weighted kernels in prefix form, which "
                                      "needs linear WL kernels (oa=False); use predict() for oa sums")
        if len(kernels) > _lib.NBW_MAX_PARTS:
            raise NotImplementedError("the fused pool path takes sums of up to %d kernels; use predict()" % _lib.NBW_MAX_PARTS)
        parts = []
        for i, (k, fit, h) in enumerate(zip(kernels, fits, hs)):
            pm = PoolModel.from_fit(eng, fit, h)
            pm.slot, pm.undirected = d['views'][i]
            if not plain:
                lw = np.ones(h + 1) if lws[i] is None else lws[i]
                pm.level_weight = [float(weights[i]) * float(x) for x in lw]
            parts.append(pm)
        first = parts[0]
        first.more = parts[1:]
        first.n_slots = d['n_slots']
        base = 0
        for pm in parts:
            pm.set_label_base(base)
            base += pm.n_labels
        T = first.total_labels = base

        def scaled_features(i):
            """B_i = diag(s) Phi_i diag(omega) [n, D_i]; the single plain kernel folds sqrt(w) instead (its
            weight cancels against the integer k_cc the kernels accumulate)."""
            fit, h = fits[i], hs[i]
            D = fit.k_end(h)
            B = eng.empty((n, D), torch.float64)
            if plain:
                eng.scale_features(fit.phi, 0, D, d['s_pool'], math.sqrt(weights[i]), B, D)
            else:
                for l in range(h + 1):
                    c0, c1 = fit.col_off[l], fit.col_off[l + 1]
                    eng.scale_features(fit.phi, c0, c1, d['s_pool'], parts[i].level_weight[l], B[:, c0:], D)
            return B, D

        def build_general():
            """G and w_mu over feature columns: needed by the oa kernel and by the general path of
            the linear kernel; the prefix form below does not use them, so they are built lazily."""
            B, D = scaled_features(0)
            Wt = eng.empty((n, D), torch.float64)
            eng.gemm(False, False, n, D, n, 1.0, Linv, n, B, D, 0.0, Wt, D)
            # one extra all-zero row/column: the landing slot of labels outside the dictionary
            G = torch.zeros((D + 1, D + 1), dtype=torch.float64, device=eng.device)
            eng.gemm(True, False, D, D, n, 1.0, Wt, D, Wt, D, 0.0, G, D + 1)
            w_mu = torch.zeros((D + 1,), dtype=torch.float64, device=eng.device)
            eng.gemm(True, False, D, 1, n, 1.0, B, D, alpha, 1, 0.0, w_mu, 1)
            return G, w_mu
        # incumbent: y_raw[argmax of the posterior mean at the training points] (acquisitions.py:82-92)
        mu_train = eng.empty((n,), torch.float64)
        eng.gemm(False, False, n, 1, n, 1.0, d['K'], n, alpha, 1, 0.0, mu_train, 1)
        if plain:
            first.general_builder = build_general
        # Histogram intersection has two evaluation forms: items (below) and the general G path ((node, level)^2
        # gathers over thermometer columns).  Measured (profiles/r02_oa_bench.json): items win on 8-node cells at
        # every depth (1.3-2.3x) and on 16-node cells from h = 3; 16-node cells at h <= 2 stay on the general path
        # (the item operator of a DARTS dictionary outgrows L2).  NBW_OA_FORM=items|general overrides.
        oa_form = os.environ.get("NBW_OA_FORM") or ("general" if fits[0].n_max == 16 and hs[0] <= 2 else "items")
        if any_oa and oa_form == "general":
            first.G, first.w_mu = build_general()
        elif any_oa:
            # Histogram intersection in ITEM form (nbw_model.oa_exc): every (node, level) of a candidate activates the
            # thermometer column (label, rank inside the label class).  Item s < T = the rank-1 columns along the
            # ancestor chain of label s; item T + x = exception e(label, r) = column (label, r) [r <= max count]
            # - column (label, 1).  Bs = B [A1^T | Dx^T], then H and w_mu exactly like the prefix form.
            fit, h = fits[0], hs[0]
            B, D = scaled_features(0)
            lo, co = fit.label_off, fit.col_off
            Tl = lo[h + 1]
            tb = fit.thermo_base[:Tl].cpu().numpy().astype(np.int64)
            mc = fit.max_count[:Tl].cpu().numpy().astype(np.int64)
            par = fit.parents[:Tl].cpu().numpy().astype(np.int64)
            level_of = np.repeat(np.arange(h + 1), [lo[l + 1] - lo[l] for l in range(h + 1)])
            lo_a, co_a = np.asarray(lo, dtype=np.int64), np.asarray(co, dtype=np.int64)
            col1 = co_a[level_of] + tb
            X = int(mc.sum())
            # rank-2 chain items (nbw_model.oa_d2_base): one per label, as long as the operator stays L2-sized
            twins = (2 * Tl + X + 1) ** 2 * 8 <= 96 << 20 and not os.environ.get("NBW_OA_NO_TWINS")
            S = Tl + X + (Tl if twins else 0)
            Kc = 2 * (h + 1)
            SKIP = 0x7FFFFFFF
            idx = np.full((S, Kc), SKIP, dtype=np.int64)
            cur, lev = np.arange(Tl), level_of.copy()
            for k in range(h + 1):
                alive = lev >= 0
                idx[:Tl][alive, k] = col1[cur[alive]]
                if twins:       # sum over the chain of e(a, 2) = column (a, 2) [if seen twice in a training cell] - column (a, 1)
                    a = cur[alive]
                    idx[Tl + X:][alive, 2 * k] = np.where(mc[a] >= 2, col1[a] + 1, SKIP)
                    idx[Tl + X:][alive, 2 * k + 1] = ~col1[a]
                up = alive & (lev > 0)
                cur[up] = lo_a[lev[up] - 1] + par[cur[up]]
                lev = np.where(up, lev - 1, -1)
            first_exc = Tl + np.cumsum(mc) - mc                      # item of e(label, 2)
            lab = np.repeat(np.arange(Tl), mc)                       # label of every exception item
            r = np.arange(X) - (first_exc[lab] - Tl) + 2             # its rank (2 .. max + 1)
            idx[Tl:Tl + X, 0] = np.where(r <= mc[lab], col1[lab] + r - 1, SKIP)
            idx[Tl:Tl + X, 1] = ~col1[lab]
            Bs = torch.zeros((n, S + 1), dtype=torch.float64, device=eng.device)
            eng.combine_columns(B, torch.from_numpy(idx.astype(np.int32)).to(eng.device), Bs, S + 1)
            Wp = eng.empty((n, S + 1), torch.float64)
            eng.gemm(False, False, n, S + 1, n, 1.0, Linv, n, Bs, S + 1, 0.0, Wp, S + 1)
            H = eng.empty((S + 1, S + 1), torch.float64)
            eng.gemm(True, False, S + 1, S + 1, n, 1.0, Wp, S + 1, Wp, S + 1, 0.0, H, S + 1)
            w_mu_p = eng.empty((S + 1,), torch.float64)
            eng.gemm(True, False, S + 1, 1, n, 1.0, Bs, S + 1, alpha, 1, 0.0, w_mu_p, 1)
            first.H, first.w_mu_prefix, first.total_labels = H, w_mu_p, S
            exc = (first_exc << 5) | np.minimum(mc, 31)
            first.oa_exc = torch.from_numpy(exc.astype(np.uint32).view(np.int32)).to(eng.device)
            first.oa_d2_base = Tl + X if twins else 0
        else:
            # prefix form (nbw_model.H): the operators summed over ancestor chains of labels, all parts
            # side by side: Bp [n, T + 1], last column zero (landing slot of labels outside the dictionary)
            Bp = torch.zeros((n, T + 1), dtype=torch.float64, device=eng.device)
            for i, pm in enumerate(parts):
                B, _ = scaled_features(i)
                eng.prefix_features(B, fits[i], hs[i], out=Bp[:, pm.label_base:], ldp=T + 1)
            ps, n_old = d.get('prev_score'), d.get('prev_n', 0)
            if ps is not None and plain and ps.H is not None and ps.h == first.h and not ps.more and 0 < n_old < n:
                # Rank-k update (row f2): the factor was APPENDED, so L^-1 of the old cells is unchanged, and — the
                # dictionary being append-only — so are their rows of Bp in the columns of old labels; in the
                # column of a NEW label an old cell carries the prefix sum of that label's nearest old ancestor
                # (its own count there is zero).  Hence H' = E H E^T + Wn^T Wn: E copies row / column p(a) of the
                # previous H to every label a (p = nearest old ancestor, a itself if old, the zero slot if none),
                # Wn = the k new rows of L'^-1 Bp'.  O(k n T + k T^2 + T^2) instead of O(n^2 T + n T^2).
                kk = n - n_old
                Wn = eng.empty((kk, T + 1), torch.float64)
                eng.gemm(False, False, kk, T + 1, n, 1.0, Linv[n_old:, :], n, Bp, T + 1, 0.0, Wn, T + 1)
                lo_o, lo_n = ps.label_off, first.label_off
                par = fits[0].parents.cpu().numpy().astype(np.int64)
                anc = np.full(T + 1, ps.total_labels, dtype=np.int64)          # default: the zero slot of the old H
                for l in range(first.h + 1):
                    d_old, d_new = lo_o[l + 1] - lo_o[l], lo_n[l + 1] - lo_n[l]
                    anc[lo_n[l]:lo_n[l] + d_old] = lo_o[l] + np.arange(d_old)
                    if l > 0 and d_new > d_old:
                        pa = par[lo_n[l] + d_old:lo_n[l + 1]]                    # level-(l-1) ids of the new labels' parents
                        anc[lo_n[l] + d_old:lo_n[l + 1]] = anc[lo_n[l - 1] + pa]
                idx = torch.from_numpy(anc).to(eng.device)
                H = ps.H.index_select(0, idx).index_select(1, idx).contiguous()
                eng.gemm(True, False, T + 1, T + 1, kk, 1.0, Wn, T + 1, Wn, T + 1, 1.0, H, T + 1)
            else:
                Wp = eng.empty((n, T + 1), torch.float64)
                eng.gemm(False, False, n, T + 1, n, 1.0, Linv, n, Bp, T + 1, 0.0, Wp, T + 1)
                H = eng.empty((T + 1, T + 1), torch.float64)
                eng.gemm(True, False, T + 1, T + 1, n, 1.0, Wp, T + 1, Wp, T + 1, 0.0, H, T + 1)
            w_mu_p = eng.empty((T + 1,), torch.float64)
            eng.gemm(True, False, T + 1, 1, n, 1.0, Bp, T + 1, alpha, 1, 0.0, w_mu_p, 1)
            first.H, first.w_mu_prefix = H, w_mu_p
        inc_idx = int(np.argmax(mu_train.cpu().numpy()))
        first.lik, first.y_mean, first.y_std = float(self.likelihood), float(self.y_mean), float(self.y_std)
        first.incumbent_idx, first.incumbent = inc_idx, float(self.y_[inc_idx])
        return first

    def set_pool_model(self, pm: "PoolModel"):
        """Install a pool model received from another rank (parallel.broadcast_pool_model)."""
        from ..engine import Engine
        self._fitted = True
        self._dev = dict(eng=Engine.get(), n_max=pm.n_max, n_slots=pm.n_slots, score=pm)
        self.likelihood = pm.lik

    def pool_model(self, acq="MEAN", xi=0.0, beta=0.0, incumbent=None, palettes=None):
        """nbw_model descriptor of the fitted surrogate configured for one acquisition."""
        return self._score_state().descriptor(acq, xi=xi, beta=beta, incumbent=incumbent, palettes=palettes)

    def upload_pool(self, X):
        """list[nx.DiGraph] | CellBatch | device records -> (device records, n_cells)."""
        if isinstance(X, torch.Tensor):
            return X, X.shape[0]
        batch = self._like_train(self._pack(X))
        return self._dev['eng'].upload_cells(batch), len(batch)

    def predict_diag(self, X_s):
        """(mu [M], var [M]) — the diagonal of predict() — through the fused pool kernel."""
        self._require_fit()
        cells, M = self.upload_pool(X_s)
        eng = self._dev['eng']
        _, mu, var, _ = eng.score_pool(self.pool_model("MEAN"), cells, M, want64=True)
        return mu.cpu(), var.cpu()

    def reset_XY(self, train_x, train_y):
        """bayesopt/gp.py:285-294."""
        self.x = train_x
        self.n = len(self.x)
        self.y_ = train_y
        self.y, self.y_mean, self.y_std = normalize_y(_as_f64(train_y))
        self.logDetK = None
        # the device state of the last fit stays around: when the new examples EXTEND the old ones (the BO loop
        # appends a batch per iteration, nasbench_optimisation.py:209-213) the next fit() appends to it instead of
        # starting over (row f2 of SURVEY §8: rank-k Cholesky append, append-only dictionary)
        if self._fitted and self._dev is not None and 'fits' in self._dev:
            self._prev = self._dev
        self._fitted = False
        self._dev = None

    def dmu_dphi(self, X_s=None, average_across_features=True, average_across_occurrences=False):
        """Derivative of the posterior mean with respect to the WL embedding of each cell in X_s
        (default: the training cells) — reference bayesopt/gp.py:296-424.

        The reference differentiates k(phi*, Phi) = (phi* . x_i) / (|phi*| |x_i|) by autograd, one cell
        at a time.  In closed form, with B = rows x_i / |x_i| and w = B^T alpha (alpha = K_i y):
            d mu / d phi* = ( w - (phi^ . w) phi^ ) / |phi*|,   phi^ = phi* / |phi*|
        w is one device GEMV over the training features; the embeddings of X_s come from the pool
        feature kernel; the per-cell vector algebra (N x D) is host-side.  Like the reference, the
        tensor form of the kernel is linear even for oa kernels.  Labels outside the training
        vocabulary are dropped (the reference mis-aligns them, see WeisfilerLehman.forward_t); every
        consumer in the reference passes the training cells, where the two agree."""
        self._require_fit()
        if len(self.kernels) != 1:
            raise NotImplementedError("dmu_dphi handles a single WL kernel")
        d = self._dev
        eng, n, k = d['eng'], self.n, self.kernels[0]
        X = k.embedding(None, pad_levels=True)                      # [n, W] training embedding, reference layout
        Y = X if X_s is None else k.embedding(X_s, pad_levels=True)
        D = X.shape[1]
        # w = B^T alpha on the device (fp64): B = diag(1/|x_i|) X
        Xd = torch.as_tensor(X / np.sqrt((X * X).sum(axis=1))[:, None], device=eng.device).contiguous()
        w = eng.empty((D,), torch.float64)
        eng.gemm(True, False, D, 1, n, float(self.weights[0]), Xd, D, d['alpha'], 1, 0.0, w, 1)
        w = w.cpu().numpy()
        ny = np.sqrt((Y * Y).sum(axis=1))
        Yh = Y / ny[:, None]
        grads = (w[None, :] - (Yh @ w)[:, None] * Yh) / ny[:, None]
        feature_matrix = torch.tensor(Y)
        if average_across_features:
            return get_grad(torch.tensor(grads), feature_matrix, average_across_occurrences)
        dmu = [torch.tensor(g).reshape(1, -1) for g in grads]
        return dmu, None, feature_matrix.sum(dim=0) if average_across_occurrences else feature_matrix

    def __getstate__(self):
        st = dict(self.__dict__)
        st['_dev'] = None
        st['_prev'] = None
        st['_pack_memo'] = None    # (re-derived from x at the next fit)
        st['_fitted'] = False      # device state does not travel; call fit() after unpickling
        return st


def get_grad(grad_matrix, feature_matrix, average_occurrences=False):
    """Monte-Carlo average of the gradients over the sample, per feature (and per occurrence count) —
    reference bayesopt/gp.py:427-476.  Host-side bookkeeping over an [N, D] matrix."""
    assert grad_matrix.shape == feature_matrix.shape
    keep = torch.nonzero((feature_matrix != 0).any(dim=0)).reshape(-1)      # drop all-zero columns
    feature_matrix, grad_matrix = feature_matrix[:, keep], grad_matrix[:, keep]
    N, D = feature_matrix.shape
    if average_occurrences:
        avg, var = torch.zeros(D), torch.zeros(D)
        for c in range(D):
            col = feature_matrix[:, c]
            _, inverse, counts = torch.unique(col, return_inverse=True, return_counts=True)
            wv = counts[inverse].to(torch.float32)
            wv = wv / wv.sum()
            mean = torch.sum(wv * grad_matrix[:, c])
            avg[c] = mean
            var[c] = torch.sum(wv * grad_matrix[:, c] ** 2) - mean ** 2
        return avg, var, feature_matrix.sum(dim=0)
    max_occur = 7                                                           # gp.py:459-461
    avg, var, inc = torch.zeros(D, max_occur), torch.zeros(D, max_occur), torch.zeros(D, max_occur)
    for c in range(D):
        col = feature_matrix[:, c]
        values, counts = torch.unique(col, return_counts=True)
        for i, val in enumerate(values):
            at = grad_matrix[col == val]        # all columns of those rows: the reference's own indexing (:471)
            avg[c, int(val)] = torch.mean(at)
            var[c, int(val)] = torch.var(at)
            inc[c, int(val)] = counts[i]
    return avg, var, inc
