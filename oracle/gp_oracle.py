"""TEST INFRASTRUCTURE — fp64 CPU restatement of the reference's GP surrogate and acquisitions.

Not part of the product (see oracle/wl_oracle.py header).  Follows, formula by formula:

* y normalisation              bayesopt/utils.py:72-85
* Cholesky inverse + "logDetK" bayesopt/utils.py:120-140   (sign quirk; jitter retry Q7)
* marginal likelihood          bayesopt/utils.py:102-117
* WL-depth grid search         bayesopt/gp.py:493-538       (un-normalised Gram, strict <, Q5)
* cosine normalisation         kernels/combine_kernels.py:36-56
* noise optimisation (Adam)    bayesopt/gp.py:179-219       (Q9)
* trainable kernel weights     bayesopt/gp.py:139-212 + kernels/combine_kernels.py:36-56
* posterior                    bayesopt/gp.py:240-283       (Q11)
* EI / AEI / UCB, incumbent    bayesopt/acquisitions.py:59-92,130-138 (Q12-Q14)
* top-n distinct by value      bayesopt/acquisitions.py:94-113 (Q15)

The reference evaluates the GP in fp32 (Q8); this restatement evaluates the same formulas in
fp64 and is validated against the reference at fp32 tolerance (tests/test_oracle_golden.py).
Quantities that the reference *stores* in fp32 and feeds back (the noise lambda) are rounded
to fp32 here too so that h*, lambda and the clamp threshold agree exactly.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Optional, Sequence

import numpy as np

from . import wl_oracle

_LOG_2PI = math.log(2.0 * math.pi)


def f32(x) -> float:
    return float(np.float32(x))


def normalize_y(y):
    """(y - mean) / std with the unbiased std of torch.std (utils.py:72-76)."""
    y = np.asarray(y, dtype=np.float64)
    mean = y.mean()
    std = y.std(ddof=1)
    return (y - mean) / std, mean, std


def pd_inverse(K: np.ndarray, jitter: float):
    """compute_pd_inverse (utils.py:120-140): Cholesky of K + jitter*10^k*I, k = 0,1,2;
    logDetK := -2*sum(log(diag(L))) i.e. MINUS log|K| (the reference's sign quirk)."""
    n = K.shape[0]
    L = None
    for k in range(3):
        try:
            L = np.linalg.cholesky(K + jitter * (10 ** k) * np.eye(n))
            break
        except np.linalg.LinAlgError:
            continue
    if L is None:
        raise RuntimeError("Gram matrix not positive definite despite of jitter")
    logdet_quirk = -2.0 * np.log(np.diag(L)).sum()
    Linv = np.linalg.solve(L, np.eye(n))
    K_i = Linv.T @ Linv
    return K_i, logdet_quirk, L


def nlml_prime(K_i, logdet_quirk, y) -> float:
    """-compute_log_marginal_likelihood(...) (utils.py:102-117), the quantity the reference
    minimises: (1/2 y^T K^-1 y + 1/2 logDetK_quirk + n/2 log 2pi) / n."""
    n = y.shape[0]
    lml = -0.5 * float(y @ K_i @ y) - 0.5 * logdet_quirk - n / 2.0 * _LOG_2PI
    return -lml / n


def grid_search_h(K_levels: np.ndarray, y_norm, lik: float, candidates: Sequence[int]):
    """_grid_search_wl_kernel (gp.py:493-538).  K_levels[l] = integer Gram of WL level l, so
    K_raw(h) = sum_{l<=h} K_levels[l].  Returns (h*, [nlml'(h) for h in candidates])."""
    best, best_h, vals = math.inf, None, []
    csum = np.cumsum(K_levels.astype(np.float64), axis=0)
    for hc in candidates:
        K_i, ld, _ = pd_inverse(csum[hc], lik)
        v = nlml_prime(K_i, ld, y_norm)
        vals.append(v)
        if v < best:
            best, best_h = v, hc
    return best_h, vals


def normalize_gram(K_raw: np.ndarray) -> np.ndarray:
    """K / sqrt(diag ⊗ diag) (combine_kernels.py:54-56)."""
    d = np.sqrt(np.diag(K_raw).astype(np.float64))
    return K_raw.astype(np.float64) / np.outer(d, d)


def optimise_noise(K: np.ndarray, y_norm, lik0: float, iters: int = 20, lr: float = 0.1,
                   max_lik: float = 0.01):
    """The Adam loop of GraphGP.fit on the scalar noise lambda (gp.py:179-210), fp32 state.

    d nlml'/d lambda = -(1/2 |K_l^-1 y|^2 + 1/2 tr K_l^-1) / n  (always negative because of
    the Q6 sign quirk).  Returns (lambda_final, nlml' of the last loop iteration)."""
    lam = np.float32(lik0)
    m = np.float32(0.0)
    v = np.float32(0.0)
    b1, b2, eps = np.float32(0.9), np.float32(0.999), np.float32(1e-8)
    n = y_norm.shape[0]
    last = None
    for t in range(1, iters + 1):
        K_i, ld, _ = pd_inverse(K, float(lam))
        last = nlml_prime(K_i, ld, y_norm)
        alpha = K_i @ y_norm
        g = np.float32(-(0.5 * float(alpha @ alpha) + 0.5 * float(np.trace(K_i))) / n)
        m = b1 * m + (np.float32(1) - b1) * g
        v = b2 * v + (np.float32(1) - b2) * g * g
        bc1 = np.float32(1.0 - 0.9 ** t)
        bc2 = np.float32(1.0 - 0.999 ** t)
        denom = np.float32(np.sqrt(v) / np.sqrt(bc2)) + eps
        lam = np.float32(lam - (np.float32(lr) / bc1) * (m / denom))
        lam = np.float32(min(max(lam, np.float32(1e-5)), np.float32(max_lik)))
    return float(lam), last


@dataclass
class GPState:
    """What GraphGP holds after fit() (gp.py:212-219) plus what predict() needs."""
    h: int
    oa: bool
    lik: float
    y_mean: float
    y_std: float
    y_norm: np.ndarray
    y_raw: np.ndarray
    K: np.ndarray            # normalised training Gram
    K_i: np.ndarray          # (K + lik I)^-1
    logDetK: float           # quirk sign
    L: np.ndarray
    diag_raw: np.ndarray     # K_raw(x_i, x_i)
    nlml: Optional[float] = None
    nlml_grid: list = field(default_factory=list)


def gp_fit(train, y, h_candidates: Sequence[int] = tuple(range(5)), oa: bool = False,
           likelihood: float = 1e-3, optimize_lik: bool = True, max_lik: float = 0.01,
           iters: int = 20, h_fixed: Optional[int] = None) -> GPState:
    """GraphGP(...).fit(...) for a single WL kernel (gp.py:107-238)."""
    y = np.asarray(y, dtype=np.float64)
    y_norm, y_mean, y_std = normalize_y(y)
    lik = f32(likelihood)
    grid = []
    if len(h_candidates):
        hmax = max(h_candidates)
        K_levels = wl_oracle.gram_raw_levels(train, hmax, oa)
        h_star, grid = grid_search_h(K_levels, y_norm, lik, h_candidates)
        K_raw = K_levels[:h_star + 1].sum(axis=0)
    else:
        h_star = h_fixed
        K_raw = wl_oracle.gram_raw(train, h_star, oa)
    K = normalize_gram(K_raw)
    nlml = None
    if optimize_lik:
        lik, nlml = optimise_noise(K, y_norm, lik, iters=iters, max_lik=max_lik)
    K_i, ld, L = pd_inverse(K, lik)
    return GPState(h=h_star, oa=oa, lik=lik, y_mean=float(y_mean), y_std=float(y_std),
                   y_norm=y_norm, y_raw=y, K=K, K_i=K_i, logDetK=ld, L=L,
                   diag_raw=np.diag(K_raw).astype(np.float64), nlml=nlml, nlml_grid=grid)


@dataclass
class SumGPState:
    """State of a GraphGP over a SUM of WL kernels with trained weights (gp.py:139-219)."""
    hs: list
    oas: list
    weights: np.ndarray      # as stored: post-step weights / their sum (gp.py:212)
    lik: float
    y_mean: float
    y_std: float
    y_norm: np.ndarray
    y_raw: np.ndarray
    K: np.ndarray            # the Gram of the LAST loop iteration (weights before the final step)
    K_i: np.ndarray
    nlml: Optional[float] = None


def weight_gradient(K_i, alpha, S, raws, n):
    """d nlml'/d w_i of nlml' = (0.5 y^T K^-1 y - 0.5 log|K| + c)/n with K = normalise(S) + lik I,
    S = sum_i w_i R_i — the closed form of what the reference gets by autograd through
    combine_kernels.py:36-56 and utils.py:102-140."""
    Mm = np.outer(alpha, alpha) + K_i
    s = 1.0 / np.sqrt(np.diag(S))
    N = S * np.outer(s, s)
    rowdot = (Mm * N).sum(axis=1)
    out = []
    for R in raws:
        out.append(-((Mm * R * np.outer(s, s)).sum() - (np.diag(R) * s * s * rowdot).sum()) / (2.0 * n))
    return np.array(out)


def gp_fit_sum(train, y, kernels, h_candidates: Sequence[int] = tuple(range(5)), likelihood: float = 1e-3,
               optimize_lik: bool = True, max_lik: float = 0.01, iters: int = 20, lr: float = 0.1,
               optimizer: str = "adam", optimize_layer_weights: bool = False) -> SumGPState:
    """GraphGP([k_1, ..., k_m]).fit(...) with trainable kernel weights and, optionally, trainable WL
    layer weights (gp.py:107-238).  kernels: [(h_initial, oa), ...].  Every kernel gets its own depth
    by the grid search of gp.py:493-538; then Adam (fp32 state, lr 0.1) moves [weights (if m > 1),
    lambda, layer weights] together; weights and layer weights are clamped to [0, 1], lambda to
    [1e-5, max_lik] after each step (:199-208).  The layer-weight vector has the length of the first
    kernel with more than one level (:158-165) and applies to every kernel of that depth; the others
    fall back to ones (weisfeiler_lehman.py:126-127)."""
    y = np.asarray(y, dtype=np.float64)
    y_norm, y_mean, y_std = normalize_y(y)
    n = len(y)
    lik0 = f32(likelihood)
    hs, levels = [], []
    for (h0, oa) in kernels:
        if len(h_candidates):
            K_levels = wl_oracle.gram_raw_levels(train, max(h_candidates), oa)
            h_star, _ = grid_search_h(K_levels, y_norm, lik0, h_candidates)
        else:
            h_star = h0
            K_levels = wl_oracle.gram_raw_levels(train, h_star, oa)
        levels.append([K_levels[l].astype(np.float64) for l in range(h_star + 1)])
        hs.append(h_star)
    m = len(kernels)
    n_lw = 0
    if optimize_layer_weights:
        for h in hs:
            if h + 1 > 1:
                n_lw = h + 1
                break
    shares = [n_lw > 0 and h + 1 == n_lw for h in hs]              # kernels the layer weights apply to
    x = np.array([1.0 / m] * m + [lik0] + [1.0] * n_lw, dtype=np.float32)
    trainable = np.array([m > 1] * m + [optimize_lik] + [True] * n_lw)
    mom, vel = np.zeros_like(x), np.zeros_like(x)
    b1, b2, eps = np.float32(0.9), np.float32(0.999), np.float32(1e-8)
    K = None
    last = None

    def kernel_gram(i, lw):
        return sum((float(lw[l]) if shares[i] else 1.0) * levels[i][l] for l in range(hs[i] + 1))
    for t in range(1, iters + 1):
        lw = x[m + 1:]
        S = sum(float(x[i]) * kernel_gram(i, lw) for i in range(m))
        K = normalize_gram(S)
        K_i, ld, _ = pd_inverse(K, float(x[m]))
        last = nlml_prime(K_i, ld, y_norm)
        alpha = K_i @ y_norm
        g = np.zeros_like(x)
        for i in range(m):
            gl = weight_gradient(K_i, alpha, S, levels[i], n)      # the form is linear in R: one value per level Gram
            g[i] = np.float32(sum((float(lw[l]) if shares[i] else 1.0) * gl[l] for l in range(hs[i] + 1)))
            if shares[i]:
                g[m + 1:] += (float(x[i]) * gl).astype(np.float32)
        g[m] = np.float32(-(0.5 * float(alpha @ alpha) + 0.5 * float(np.trace(K_i))) / n)
        mom = b1 * mom + (np.float32(1) - b1) * g
        vel = b2 * vel + (np.float32(1) - b2) * g * g
        bc1, bc2 = np.float32(1.0 - 0.9 ** t), np.float32(1.0 - 0.999 ** t)
        step = (np.float32(lr) / bc1) * (mom / (np.sqrt(vel) / np.sqrt(bc2) + eps))
        if optimizer == "sgd":
            step = np.float32(lr) * g
        x = np.where(trainable, x - step.astype(np.float32), x).astype(np.float32)
        x[:m] = np.clip(x[:m], np.float32(0.0), np.float32(1.0))
        x[m] = np.float32(min(max(x[m], np.float32(1e-5)), np.float32(max_lik)))
        x[m + 1:] = np.clip(x[m + 1:], np.float32(0.0), np.float32(1.0))
    lik = float(x[m])
    K_i, _, _ = pd_inverse(K, lik)
    w = x[:m].astype(np.float64)
    st = SumGPState(hs=hs, oas=[oa for _, oa in kernels], weights=w / w.sum(), lik=lik, y_mean=float(y_mean),
                    y_std=float(y_std), y_norm=y_norm, y_raw=y, K=K, K_i=K_i, nlml=last)
    st.layer_weights = x[m + 1:].astype(np.float64) if n_lw else None
    st.shares = shares
    return st


def gp_predict_full_sum(st: SumGPState, train, pool):
    """predict() of the multi-kernel GP (gp.py:240-283): joint Grams per kernel at the fitted depths,
    weighted by the STORED weights (which are one optimiser step ahead of K_i, as in the reference)."""
    n, M = len(train), len(pool)
    from nasbowl_b200.cells import CellBatch
    joint = CellBatch.concat([train, pool])
    lw = getattr(st, "layer_weights", None)
    S = 0.0
    for i, (w, h, oa) in enumerate(zip(st.weights, st.hs, st.oas)):
        lev = wl_oracle.gram_raw_levels(joint, h, oa).astype(np.float64)
        for l in range(h + 1):
            S = S + w * (lw[l] if lw is not None and st.shares[i] else 1.0) * lev[l]
    K_full = normalize_gram(S)
    K_s = K_full[:n, n:]
    K_ss = K_full[n:, n:] + st.lik * np.eye(M)
    mu = (K_s.T @ st.K_i @ st.y_norm) * st.y_std + st.y_mean
    cov = np.maximum(K_ss - K_s.T @ st.K_i @ K_s, st.lik)
    cov = (np.sqrt(cov) * st.y_std) ** 2
    return mu, cov


def _per_kernel(x, m):
    """One CellBatch for every kernel of the sum (the reference's behaviour, combine_kernels.py:36-56), or a
    list with one CellBatch per kernel (kernels that read different record slots of an example)."""
    return list(x) if isinstance(x, (list, tuple)) else [x] * m


def gp_fit_sum_fixed(train, y, kernels, weights, likelihood: float = 1e-3, layer_weights=None):
    """GraphGP(train, y, [k_1..k_m], weights=w).fit(wl_subtree_candidates=(), optimize_lik=False): a sum of WL
    kernels with FIXED weights and depths (gp.py:139-219 with nothing to train).  kernels: [(h, oa), ...];
    train: see _per_kernel; layer_weights: optional list (per kernel) of per-level weights."""
    y = np.asarray(y, dtype=np.float64)
    y_norm, y_mean, y_std = normalize_y(y)
    m = len(kernels)
    trains = _per_kernel(train, m)
    w = np.array([f32(v) for v in weights], dtype=np.float64)      # the reference keeps the weights in fp32 (gp.py:66)
    lik = f32(likelihood)
    S = 0.0
    for i, (h, oa) in enumerate(kernels):
        lev = wl_oracle.gram_raw_levels(trains[i], h, oa).astype(np.float64)
        lw = np.ones(h + 1) if layer_weights is None or layer_weights[i] is None else np.asarray(layer_weights[i], dtype=np.float64)
        S = S + w[i] * sum(lw[l] * lev[l] for l in range(h + 1))
    K = normalize_gram(S)
    K_i, _, _ = pd_inverse(K, lik)
    st = SumGPState(hs=[h for h, _ in kernels], oas=[oa for _, oa in kernels], weights=w / w.sum(), lik=lik,
                    y_mean=float(y_mean), y_std=float(y_std), y_norm=y_norm, y_raw=y, K=K, K_i=K_i, nlml=None)
    st.raw_weights = w
    st.level_weights = [np.ones(h + 1) if layer_weights is None or layer_weights[i] is None
                        else np.asarray(layer_weights[i], dtype=np.float64) for i, (h, _) in enumerate(kernels)]
    return st


def gp_predict_diag_sum(st: SumGPState, train, pool, chunk: int = 2048):
    """Posterior mean and marginal variance of a sum-of-kernels GP over a pool (the diagonal of
    gp_predict_full_sum, gp.py:240-283), chunked.  Uses st.weights and, if present, st.level_weights."""
    m = len(st.hs)
    trains, pools = _per_kernel(train, m), _per_kernel(pool, m)
    M = len(pools[0])
    mu, var = np.empty(M), np.empty(M)
    a = st.K_i @ st.y_norm
    lws = getattr(st, "level_weights", None)
    for s in range(0, M, chunk):
        K_s, kss, ktt = 0.0, 0.0, 0.0
        for i, (w, h, oa) in enumerate(zip(st.weights, st.hs, st.oas)):
            lw = np.ones(h + 1) if lws is None else lws[i]
            Ks_i, kss_i, ktt_i = wl_oracle.pool_features_weighted(trains[i], pools[i][s:s + chunk], h, oa, lw)
            K_s, kss, ktt = K_s + w * Ks_i, kss + w * kss_i, ktt + w * ktt_i
        K_s = K_s / np.sqrt(np.outer(ktt, kss))
        mu[s:s + chunk] = (K_s.T @ a) * st.y_std + st.y_mean
        q = np.einsum("ij,ij->j", K_s, st.K_i @ K_s)
        cov = np.maximum(1.0 + st.lik - q, st.lik)
        var[s:s + chunk] = (np.sqrt(cov) * st.y_std) ** 2
    return mu, var


def incumbent_sum(st: SumGPState) -> float:
    """_get_incumbent (acquisitions.py:82-92) for a sum-of-kernels state: y_raw[argmax of the posterior mean at
    the training points]."""
    mu_train = st.K @ (st.K_i @ st.y_norm)
    return float(st.y_raw[int(np.argmax(mu_train))])


def gp_predict_diag(st: GPState, train, pool, chunk: int = 4096):
    """Posterior mean and marginal variance over a pool (gp.py:240-283, diagonal only):
    mu = (K_s^T K_i y)*y_std + y_mean ; var = clamp(1 + lik - k_s^T K_i k_s, lik) * y_std^2."""
    M = len(pool)
    mu = np.empty(M)
    var = np.empty(M)
    a = st.K_i @ st.y_norm
    for s in range(0, M, chunk):
        sub = pool[s:s + chunk]
        K_s_raw, kss, ktt = wl_oracle.pool_features(train, sub, st.h, st.oa)
        K_s = K_s_raw / np.sqrt(np.outer(ktt, kss))
        mu[s:s + chunk] = (K_s.T @ a) * st.y_std + st.y_mean
        q = np.einsum("ij,ij->j", K_s, st.K_i @ K_s)
        cov = np.maximum(1.0 + st.lik - q, st.lik)
        var[s:s + chunk] = (np.sqrt(cov) * st.y_std) ** 2
    return mu, var


def gp_predict_full(st: GPState, train, pool):
    """(mu [M], cov [M, M]) exactly as GraphGP.predict returns them, incl. the element-wise
    clamp of the full covariance (gp.py:276)."""
    n, M = len(train), len(pool)
    from nasbowl_b200.cells import CellBatch
    joint = CellBatch.concat([train, pool])
    K_full = normalize_gram(wl_oracle.gram_raw(joint, st.h, st.oa))
    K_s = K_full[:n, n:]
    K_ss = K_full[n:, n:] + st.lik * np.eye(M)
    mu = (K_s.T @ st.K_i @ st.y_norm) * st.y_std + st.y_mean
    cov = np.maximum(K_ss - K_s.T @ st.K_i @ K_s, st.lik)
    cov = (np.sqrt(cov) * st.y_std) ** 2
    return mu, cov


def dmu_dphi(st: GPState, X: np.ndarray, Y: np.ndarray, weight: float = 1.0) -> np.ndarray:
    """Gradient of the posterior mean with respect to the embedding of each row of Y
    (gp.py:296-376): the vector-Jacobian product of k(Y, X) = (Y X^T) / (|y| |x_i|)
    (grakel_replace/utils.py:60-71, the tensor form is linear whatever `oa` is) with
    alpha = K_i y.  X, Y: embeddings from wl_oracle.wl_embed_reference_order.  Returns [M, D]."""
    alpha = st.K_i @ st.y_norm
    nx = np.sqrt((X * X).sum(axis=1))
    out = np.empty_like(Y)
    for j in range(Y.shape[0]):
        y = Y[j]
        ny = np.sqrt(y @ y)
        # d/dy [ (y . x_i) / (ny nx_i) ] = x_i / (ny nx_i) - (y . x_i) y / (ny^3 nx_i)
        J = X / (ny * nx)[:, None] - np.outer((X @ y) / (ny ** 3 * nx), y)
        out[j] = weight * (alpha @ J)
    return out


def get_grad(grad_matrix: np.ndarray, feature_matrix: np.ndarray, average_occurrences: bool = False):
    """gp.py:427-476, incl. its indexing at :471 (the mean / variance run over ALL columns of the
    rows whose feature d equals the value) and torch.var's n-1 denominator (nan for one row)."""
    keep = [c for c in range(feature_matrix.shape[1]) if not np.all(feature_matrix[:, c] == 0)]
    F, G = feature_matrix[:, keep], grad_matrix[:, keep]
    N, D = F.shape
    if average_occurrences:
        avg, var = np.zeros(D), np.zeros(D)
        for d in range(D):
            vals, inverse, counts = np.unique(F[:, d], return_inverse=True, return_counts=True)
            w = counts[inverse].astype(np.float32)
            w = w / w.sum()
            g = G[:, d]
            mean = np.sum(w * g)
            avg[d] = mean
            var[d] = np.sum(w * g ** 2) - mean ** 2
        return avg, var, F.sum(axis=0)
    max_occur = 7
    avg, var, inc = np.zeros((D, max_occur)), np.zeros((D, max_occur)), np.zeros((D, max_occur))
    for d in range(D):
        vals, counts = np.unique(F[:, d], return_counts=True)
        for i, val in enumerate(vals):
            at = G[F[:, d] == val]
            avg[d, int(val)] = at.mean()
            with np.errstate(invalid="ignore", divide="ignore"):
                var[d, int(val)] = at.var(ddof=1) if at.size > 1 else np.nan
            inc[d, int(val)] = counts[i]
    return avg, var, inc


def incumbent(st: GPState) -> float:
    """_get_incumbent (acquisitions.py:82-92): `in_fill == 'max'` is never true (Q13), so the
    incumbent is y_raw[argmax of the posterior mean at the training points]."""
    mu_train = st.K @ (st.K_i @ st.y_norm)
    return float(st.y_raw[int(np.argmax(mu_train))])


def _phi(u):
    return np.exp(-0.5 * u * u) / math.sqrt(2.0 * math.pi)


def _Phi(u):
    from scipy.special import erfc
    return 0.5 * erfc(-u / math.sqrt(2.0))


def expected_improvement(mu, var, mu_star: float, xi: float = 0.0, augmented: bool = False,
                         lik: float = 0.0):
    """EI = std*pdf(u) + (mu - mu* - xi)*cdf(u), u = (mu - mu* - xi)/std (acquisitions.py:68-74);
    AEI factor 1 - sqrt(lik)/sqrt(lik + var) (:75-77)."""
    std = np.sqrt(var)
    d = mu - mu_star - xi
    u = d / std
    ei = std * _phi(u) + d * _Phi(u)
    if augmented:
        ei = ei * (1.0 - math.sqrt(lik) / np.sqrt(lik + var))
    return ei


def ucb_beta(iters: int) -> float:
    """3*sqrt(0.5*log(2*iters+1)) (acquisitions.py:133-135) — evaluated, like the reference,
    with fp32 torch scalars."""
    import torch
    return float(3. * torch.sqrt(0.5 * torch.log(torch.tensor(2. * iters + 1.))))


def upper_confidence_bound(mu, var, beta: float):
    return mu + beta * np.sqrt(var)


def top_n_distinct(scores: np.ndarray, top_n: int) -> np.ndarray:
    """propose_location's selection (acquisitions.py:101-105): first index of each distinct
    score value, then the top_n largest distinct values.  Order unspecified -> returns the
    indices sorted by descending score (compare as sets against the reference)."""
    vals, first = np.unique(np.asarray(scores).reshape(-1), return_index=True)
    take = min(top_n, vals.shape[0])
    sel = np.argsort(-vals, kind="stable")[:take]
    return first[sel]
