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
from copy import deepcopy
from typing import List, Optional, Sequence

import networkx as nx
import numpy as np
import torch

from .. import _lib
from ..cells import CellBatch
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


class GraphGP:
    def __init__(self, train_x, train_y, kernels: list, vectorial_features: list = None, likelihood=1e-3,
                 weights=None, vector_theta_bounds: tuple = (1e-5, 0.1), graph_theta_bounds: tuple = (1e-1, 1.e1),
                 verbose=False, n_max=None):
        assert len(train_x) == train_y.shape[0], 'mismatch of length between train and test GP'
        if not isinstance(train_x, CellBatch):
            assert all([isinstance(x, nx.Graph) for x in train_x]), \
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
        self.theta_vector = None
        self.layer_weights = None
        self.nlml = None
        self.nlml_grid = {}
        self.logDetK = None
        self._fitted = False
        self._dev = None           # device state of the last fit (not pickled)

    # ------------------------------------------------------------------ device helpers
    def _engine(self):
        from ..engine import Engine
        return Engine.get()

    def _pack(self, gr) -> CellBatch:
        return self.kernels[0].pack(gr)

    def _like_train(self, batch: CellBatch) -> CellBatch:
        """Bring a pool batch to the record width the GP was fitted with."""
        n_max = self._dev['n_max']
        if batch.n_max > n_max:
            raise ValueError("pool cells have up to %d nodes but the GP was fitted on %d-node records; "
                             "construct the GP with n_max=16" % (batch.n_max, n_max))
        return batch.widen(n_max)

    def _y_device(self, eng):
        return torch.as_tensor(np.asarray(self.y.detach().cpu().numpy(), dtype=np.float64)).to(eng.device)

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
        for hc in subtree_candidates:
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
        batch = self._pack(self.x)
        if self.n_max is not None:
            batch = batch.widen(self.n_max)
        cells = eng.upload_cells(batch)
        y_dev = self._y_device(eng)
        self.nlml_grid = {}

        # 1. WL features once per kernel, deep enough for every candidate; h by marginal likelihood
        fits = []
        for k in self.sum_kernels.kernels:
            if not isinstance(k, WeisfilerLehman):
                logging.warning('Graph kernel optimisation for ' + str(k) + " not implemented yet.")
                raise NotImplementedError("only WeisfilerLehman kernels run on the GPU path")
            cands = [int(c) for c in wl_subtree_candidates]
            h_fit = max(cands + [k.h])
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
            optim = (torch.optim.Adam if optimizer.lower() == 'adam' else torch.optim.SGD)(optim_vars,
                                                                                         **optimizer_kwargs)
            for i in range(iters):
                optim.zero_grad()
                if levels is not None:
                    # K moves with the weights: rebuild and refactor every iteration.  The K of the LAST
                    # iteration (weights before the final step) is the one the reference keeps (gp.py:186-215).
                    with torch.no_grad():
                        K, inv_sqrt_diag = build_K()
                    factor_cache.clear()
                _, Linv_i, alpha_i, logdet, yKy, trKinv, alpha_sq = factor(float(likelihood.item()))
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
                optim.step()
                with torch.no_grad():
                    if train_weights:
                        weights.clamp_(0., 1.)
                    if optimize_lik:
                        likelihood.clamp_(1e-5, max_lik)
                    if layer_weights is not None:
                        layer_weights.clamp_(0., 1.)
            weights = weights.detach()
        if layer_weights is not None:
            layer_weights = layer_weights.detach()
            for ki, k in enumerate(kernels):       # predict() runs with the final layer weights (gp.py:218,269)
                if shares[ki]:
                    k.change_kernel_params({'layer_weights': layer_weights.clone()})
        self.layer_weights = layer_weights
        lik = float(likelihood.item())
        L, Linv, alpha, logdet, yKy, _, _ = factor(lik)

        # 4. state (gp.py:212-219)
        self.weights = weights.clone() / torch.sum(weights)
        self.likelihood = lik
        self.logDetK = torch.tensor(-logdet, dtype=torch.float64)
        self.nlml = torch.tensor(nlml, dtype=torch.float64) if nlml is not None else None
        self.sum_kernels.weights = weights.clone()
        self._fitted = True
        self._dev = dict(eng=eng, fits=fits, hs=[k.h for k in self.sum_kernels.kernels], K=K, L=L, Linv=Linv,
                         alpha=alpha, s=inv_sqrt_diag, y=y_dev, n_max=batch.n_max, cells=cells, batch=batch,
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
            self._dev['K_host'] = self._dev['K'].cpu()
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
            self._dev['Ki_host'] = Ki.cpu()
        return self._dev['Ki_host']

    def _require_fit(self):
        if not self._fitted or self._dev is None:
            raise ValueError("Inverse of Gram matrix is not instantiated. Please call the optimize function to "
                             "fit on the training data first!")

    # ------------------------------------------------------------------ predict (reference path)
    def predict(self, X_s, preserve_comp_graph=False):
        """Kriging predictions (mu [M], cov [M, M]) exactly as bayesopt/gp.py:240-283: joint WL
        on train ∪ test, full M x M covariance with the element-wise clamp."""
        if not isinstance(X_s, (list, CellBatch)):
            X_s = [X_s]
        self._require_fit()
        d = self._dev
        eng = d['eng']
        n, M = self.n, len(X_s)
        test = self._like_train(self._pack(X_s))
        cells_all = torch.cat([d['cells'], eng.upload_cells(test)], dim=0).contiguous()
        grams = []
        for k, h in zip(self.sum_kernels.kernels, d['hs']):
            fit = eng.wl_fit(cells_all, d['n_max'], h, k.oa)
            k._fit = fit          # the reference leaves the kernel fitted on train ∪ test (Q19)
            k._fit_batch, k._ref_cache = CellBatch.concat([d['batch'], test]), None
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
        return mu_s, cov.cpu()

    # ------------------------------------------------------------------ fused pool path
    def _score_state(self):
        """Feature-space posterior operators (see csrc/score_pool.cu): w_mu = B^T alpha,
        G = (L^-1 B)^T (L^-1 B) with B = diag(1/sqrt(K_raw,ii)) sqrt(w) Phi_train."""
        self._require_fit()
        d = self._dev
        if d.get('score') is not None:
            return d['score']
        if len(self.sum_kernels.kernels) != 1:
            raise NotImplementedError("the fused pool path handles a single WL kernel; use predict()")
        if self.sum_kernels.kernels[0]._level_weights(d['hs'][0]) is not None:
            raise NotImplementedError("the fused pool path assumes unit WL layer weights; use predict()")
        eng, n = d['eng'], self.n
        fit, h, w = d['fits'][0], d['hs'][0], d['weights'][0]
        D = fit.k_end(h)
        # the normalisation of K uses diag(w * K_raw): s already contains 1/sqrt(w K_raw,ii)
        B = eng.empty((n, D), torch.float64)
        eng.scale_features(fit.phi, 0, D, d['s'], math.sqrt(w), B, D)
        Linv, alpha = d['Linv'], d['alpha']

        def build_general():
            """G and w_mu over feature columns: needed by the oa kernel and by the general path of
            the linear kernel; the prefix form below does not use them, so they are built lazily."""
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
        eng.gemm(False, False, n, 1, n, 1.0, d['K'], n, d['alpha'], 1, 0.0, mu_train, 1)
        inc_idx = int(np.argmax(mu_train.cpu().numpy()))
        pm = PoolModel.from_fit(eng, fit, h)
        pm.general_builder = build_general
        if fit.oa:
            pm.G, pm.w_mu = build_general()
        if not fit.oa:
            # prefix form (nbw_model.H): the same operators summed over ancestor chains of labels
            T = fit.label_off[h + 1]
            Bp = eng.prefix_features(B, fit, h)                       # [n, T+1], last column zero
            Wp = eng.empty((n, T + 1), torch.float64)
            eng.gemm(False, False, n, T + 1, n, 1.0, d['Linv'], n, Bp, T + 1, 0.0, Wp, T + 1)
            H = eng.empty((T + 1, T + 1), torch.float64)
            eng.gemm(True, False, T + 1, T + 1, n, 1.0, Wp, T + 1, Wp, T + 1, 0.0, H, T + 1)
            w_mu_p = eng.empty((T + 1,), torch.float64)
            eng.gemm(True, False, T + 1, 1, n, 1.0, Bp, T + 1, d['alpha'], 1, 0.0, w_mu_p, 1)
            pm.H, pm.w_mu_prefix = H, w_mu_p
        pm.lik, pm.y_mean, pm.y_std = float(self.likelihood), float(self.y_mean), float(self.y_std)
        pm.incumbent_idx, pm.incumbent = inc_idx, float(self.y_[inc_idx])
        d['score'] = pm
        return pm

    def set_pool_model(self, pm: "PoolModel"):
        """Install a pool model received from another rank (parallel.broadcast_pool_model)."""
        from ..engine import Engine
        self._fitted = True
        self._dev = dict(eng=Engine.get(), n_max=pm.n_max, score=pm)
        self.likelihood = pm.lik

    def pool_model(self, acq="MEAN", xi=0.0, beta=0.0, incumbent=None):
        """nbw_model descriptor of the fitted surrogate configured for one acquisition."""
        return self._score_state().descriptor(acq, xi=xi, beta=beta, incumbent=incumbent)

    def upload_pool(self, X):
        """list[nx.DiGraph] | CellBatch | device records -> (device records, n_cells)."""
        d = self._dev
        if isinstance(X, torch.Tensor):
            return X, X.shape[0]
        batch = self._like_train(self._pack(X))
        return d['eng'].upload_cells(batch), len(batch)

    def predict_diag(self, X_s):
        """(mu [M], var [M]) — the diagonal of predict() — through the fused pool kernel."""
        self._require_fit()
        eng = self._dev['eng']
        cells, M = self.upload_pool(X_s)
        _, mu, var, _ = eng.score_pool(self.pool_model("MEAN"), cells, M, want64=True)
        return mu.cpu(), var.cpu()

    def reset_XY(self, train_x, train_y):
        """bayesopt/gp.py:285-294."""
        self.x = train_x
        self.n = len(self.x)
        self.y_ = train_y
        self.y, self.y_mean, self.y_std = normalize_y(_as_f64(train_y))
        self.logDetK = None
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
