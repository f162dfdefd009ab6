"""TEST INFRASTRUCTURE — differential stress of the CPU oracle against the LIVE, unmodified reference.

    python oracle/stress_vs_reference.py [first_seed] [n_seeds]        # build container only (/root/reference)
    python oracle/stress_vs_reference.py sum [first_seed] [n_seeds]    # sums of two WL kernels with trained weights

tests/test_oracle_vs_reference.py walks two fixed seeds per family inside the suite; this script walks as many as one
cares to wait for (three families per seed, random training-set sizes 6..80, random depth grids, both base kernels,
optimised and fixed noise, directed and undirected cells): integer Gram bit for bit, chosen depth, fitted noise,
incumbent, and — at the tolerances the reference's fp32 posterior allows — mean, full covariance, EI and UCB.
The summary of the run behind the numbers in DESIGN.md is profiles/r02_oracle_stress.log.
"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nasbowl_b200.cells import OpVocab, pack_networkx  # noqa: E402
from oracle import gp_oracle, make_golden as mg, wl_oracle  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402


def one_case(R, seed, family):
    mg.seed_all(seed)
    n_train, n_pool = int(np.random.randint(6, 80)), 12
    graphs = mg.sample_cells(R, family, n_train + n_pool)
    batch = pack_networkx(graphs, OpVocab())
    train_g, pool_g = graphs[:n_train], graphs[n_train:]
    train, pool = batch[:n_train], batch[n_train:]
    # a target the cells explain, so that EI is not in the far tail everywhere
    y = torch.tensor([0.25 * g.number_of_edges() + 0.5 * sum('3x3' in d['op_name'] for _, d in g.nodes(data=True))
                      for g in train_g], dtype=torch.float32) + 0.3 * torch.randn(n_train)
    # ---- integer part
    for oa in (False, True):
        h = int(np.random.randint(0, 6))
        K_ref, _, _ = mg.ref_levels(R, train_g, h, oa)
        assert np.array_equal(np.rint(K_ref).astype(np.int64), wl_oracle.gram_raw(train, h, oa)), ("gram", h, oa)
    h = int(np.random.randint(0, 4))
    K_ref = R.WeisfilerLehman(h=h, undirected=True).fit_transform(train_g, rebuild_model=True).numpy()
    assert np.array_equal(np.rint(K_ref).astype(np.int64), wl_oracle.gram_raw(train.to_undirected(), h, False)), \
        ("undirected gram", h)
    # ---- GP
    oa, opt = bool(seed % 2), bool((seed // 2) % 2)
    cands = tuple(range(int(np.random.randint(0, 2)), int(np.random.randint(2, 5))))
    gp = R.GraphGP(train_g, y.clone(), [R.WeisfilerLehman(h=2, oa=oa)])
    gp.fit(wl_subtree_candidates=cands, optimize_lik=opt, max_lik=0.01)
    st = gp_oracle.gp_fit(train, y.numpy().astype(np.float64), h_candidates=cands, oa=oa, optimize_lik=opt, max_lik=0.01)
    assert st.h == gp.kernels[0].h, ("h", st.h, gp.kernels[0].h)
    assert abs(st.lik - float(gp.likelihood)) < 1e-7, ("lik", st.lik, float(gp.likelihood))
    with torch.no_grad():
        mu_r, cov_r = gp.predict(pool_g)
        ei = R.GraphExpectedImprovement(gp)
        ei_r = np.array([float(ei.eval(c)) for c in pool_g])
        inc_r = float(ei._get_incumbent())
        ucb = R.GraphUpperConfidentBound(gp)
        ucb.iters = 1
        ucb_r = np.array([float(ucb.eval(c)) for c in pool_g])
    mu_o, cov_o = gp_oracle.gp_predict_full(st, train, pool)
    np.testing.assert_allclose(mu_o, mu_r.numpy(), rtol=3e-3, atol=3e-3)
    np.testing.assert_allclose(cov_o, cov_r.numpy(), rtol=5e-2, atol=2e-3)        # fp32 K_ss - K_s^T K_i K_s
    var_o = np.diag(cov_o)
    assert abs(gp_oracle.incumbent(st) - inc_r) < 1e-5, ("incumbent", gp_oracle.incumbent(st), inc_r)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent(st))
    np.testing.assert_allclose(ei_o, ei_r, rtol=6e-2, atol=2e-3)
    ucb_o = gp_oracle.upper_confidence_bound(mu_o, var_o, gp_oracle.ucb_beta(1))
    # beta * (sqrt(var +- 2e-3) - sqrt(var)) at var ~ lambda = 0.01: a few hundredths
    np.testing.assert_allclose(ucb_o, ucb_r, rtol=3e-2, atol=5e-2)


def one_sum_case(R, seed, family):
    """GraphGP over a SUM of two WL kernels with trainable weights (gp.py:139-212).  Returns "degenerate" for the
    corner the reference itself cannot finish: when both kernels end up identical (same base kernel, same depth
    after the grid search) the marginal likelihood does not depend on the weights, the true gradient is zero, and
    the reference's Adam walks both weights down on fp32 rounding noise until the clamp at 0 turns K into 0/0
    (RuntimeError "Gram matrix not positive definite").  The oracle's gradient is exactly zero there and the
    weights stay at 1/m."""
    mg.seed_all(seed)
    n_train, n_pool = int(np.random.randint(10, 50)), 8
    graphs = mg.sample_cells(R, family, n_train + n_pool)
    batch = pack_networkx(graphs, OpVocab())
    train_g, pool_g = graphs[:n_train], graphs[n_train:]
    train, pool = batch[:n_train], batch[n_train:]
    y = torch.tensor([0.25 * g.number_of_edges() + 0.5 * sum('3x3' in d['op_name'] for _, d in g.nodes(data=True))
                      for g in train_g], dtype=torch.float32) + 0.3 * torch.randn(n_train)
    ks = [(int(np.random.randint(0, 3)), bool(np.random.randint(0, 2))) for _ in range(2)]
    cands = (0, 1, 2) if seed % 2 else ()
    yo = y.numpy().astype(np.float64)

    def ref_fit(**kw):
        gp = R.GraphGP(train_g, y.clone(), [R.WeisfilerLehman(h=h, oa=oa) for h, oa in ks])
        gp.fit(wl_subtree_candidates=cands, optimize_lik=True, max_lik=0.01, **kw)
        return gp
    # one Adam step and three SGD steps pin the gradient (sign and value); tight
    gp = ref_fit(iters=1)
    st = gp_oracle.gp_fit_sum(train, yo, ks, h_candidates=cands, optimize_lik=True, max_lik=0.01, iters=1)
    assert st.hs == [k.h for k in gp.kernels], ("hs", st.hs, [k.h for k in gp.kernels])
    if st.hs[0] == st.hs[1] and ks[0][1] == ks[1][1]:
        return "degenerate"
    np.testing.assert_allclose(st.weights, gp.weights.detach().numpy(), atol=2e-5)
    gp = ref_fit(iters=3, optimizer='sgd', optimizer_kwargs={'lr': 0.02})
    st = gp_oracle.gp_fit_sum(train, yo, ks, h_candidates=cands, optimize_lik=True, max_lik=0.01, iters=3, lr=0.02,
                              optimizer="sgd")
    np.testing.assert_allclose(st.weights, gp.weights.detach().numpy(), rtol=2e-4, atol=2e-6)
    np.testing.assert_allclose(st.nlml, float(gp.nlml), rtol=3e-3, atol=1e-4)
    # the default run: 20 Adam steps.  Adam's step m / sqrt(v) does not shrink with the gradient, so where the
    # likelihood is flat in the weights the reference's fp32 gradient noise moves them by ~1e-3 per step
    gp = ref_fit()
    st = gp_oracle.gp_fit_sum(train, yo, ks, h_candidates=cands, optimize_lik=True, max_lik=0.01)
    np.testing.assert_allclose(st.weights, gp.weights.detach().numpy(), atol=2e-2)
    assert abs(st.lik - float(gp.likelihood)) < 1e-6, ("lik", st.lik, float(gp.likelihood))
    np.testing.assert_allclose(st.nlml, float(gp.nlml), rtol=3e-3, atol=1e-4)
    with torch.no_grad():
        mu_r, cov_r = gp.predict(pool_g)
    mu, cov = gp_oracle.gp_predict_full_sum(st, train, pool)
    np.testing.assert_allclose(mu, mu_r.numpy(), rtol=1e-2, atol=1e-2)
    np.testing.assert_allclose(cov, cov_r.numpy(), rtol=5e-2, atol=3e-3)
    return "ok"


def main():
    args = [a for a in sys.argv[1:] if a != "sum"]
    sums = "sum" in sys.argv[1:]
    first = int(args[0]) if len(args) > 0 else 9100
    count = int(args[1]) if len(args) > 1 else 120
    R = load_reference()
    bad, n, degenerate, t0 = 0, 0, 0, time.time()
    for seed in range(first, first + count):
        for family in ("nasbench101", "nasbench201", "darts"):
            n += 1
            try:
                if sums:
                    degenerate += one_sum_case(R, seed, family) == "degenerate"
                else:
                    one_case(R, seed, family)
            except AssertionError as e:
                bad += 1
                print("MISMATCH seed=%d family=%s: %s" % (seed, family, str(e)[:400]), flush=True)
    print("oracle vs live reference%s: %d cases (seeds %d..%d x 3 families), %d mismatches%s, %.0f s"
          % (" (sums of two kernels, trained weights)" if sums else "", n, first, first + count - 1, bad,
             ", %d degenerate (identical kernels: depth checked only)" % degenerate if sums else "", time.time() - t0))
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
