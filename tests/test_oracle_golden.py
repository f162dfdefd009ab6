"""The CPU oracle (oracle/*.py) against golden vectors produced by the unmodified reference
(oracle/make_golden.py).  This is what pins parity: the GPU tests then compare the CUDA path
against the oracle."""
import os

import numpy as np
import pytest

from conftest import golden_kraw, load_golden
from nasbowl_b200.cells import CellBatch
from oracle import gp_oracle, wl_oracle


def test_appendix_b_known_answer():
    z, batch = load_golden("appendix_b")
    train, test = batch[:4], batch[4:]
    names = list(z["op_names"])
    ids, dims = wl_oracle.wl_labels_reference_order(train, 2, names)
    assert [d[1] for d in dims] == list(z["feature_dims_h2"][1:])
    for l in range(3):
        phi = wl_oracle.histogram(ids[l], *dims[l])
        ref = z["phi_h2_l%d" % l][:, :phi.shape[1]]         # drop the spurious column (Q17)
        assert np.array_equal(phi, ref)
    K = wl_oracle.gram_raw(train, 2)
    assert np.array_equal(K, z["Kraw_h2_oa0"])
    assert np.array_equal(K, [[9, 7, 7, 7], [7, 9, 7, 7], [7, 7, 12, 7], [7, 7, 7, 12]])
    st = gp_oracle.gp_fit(train, z["y"], h_candidates=(0, 1, 2), optimize_lik=False)
    assert st.h == int(z["gp_h"]) == 2
    assert st.lik == float(z["gp_lik"])
    np.testing.assert_allclose(st.K, z["gp_K"], rtol=2e-6)
    np.testing.assert_allclose(st.K_i, z["gp_Ki"], rtol=2e-4)
    np.testing.assert_allclose(st.logDetK, z["gp_logDetK"], rtol=1e-5)
    mu, cov = gp_oracle.gp_predict_full(st, train, test)
    np.testing.assert_allclose(mu, z["gp_mu"], rtol=1e-4)
    np.testing.assert_allclose(cov, z["gp_cov"], rtol=2e-3, atol=1e-6)
    mu_d, var_d = gp_oracle.gp_predict_diag(st, train, test)
    np.testing.assert_allclose(mu_d, mu, rtol=1e-12)
    np.testing.assert_allclose(var_d, np.diag(cov), rtol=1e-10)
    inc = gp_oracle.incumbent(st)
    assert inc == pytest.approx(float(z["incumbent"]))
    ei = gp_oracle.expected_improvement(mu_d, var_d, inc)
    np.testing.assert_allclose(ei, z["ei"], rtol=5e-3, atol=1e-7)
    ucb = gp_oracle.upper_confidence_bound(mu_d, var_d, gp_oracle.ucb_beta(1))
    np.testing.assert_allclose(ucb, z["ucb"], rtol=1e-4)


@pytest.mark.parametrize("oa", [False, True])
@pytest.mark.parametrize("h", [0, 1, 2, 3, 5])
def test_raw_gram_bit_exact(family, h, oa):
    _, z, train, _ = family
    ref = z["Kraw_h%d_oa%d" % (h, int(oa))]
    assert np.array_equal(wl_oracle.gram_raw(train, h, oa, fast=True), ref)
    if h <= 3:
        assert np.array_equal(wl_oracle.gram_raw(train, h, oa, fast=False), ref)


def test_histograms_identical_to_reference(family):
    """String-credential restatement reproduces the reference's label ids column for column."""
    _, z, train, _ = family
    ids, dims = wl_oracle.wl_labels_reference_order(train, 3, list(z["op_names"]))
    assert [0] + [d[1] for d in dims] == list(z["feature_dims_h3"])
    for l in range(4):
        phi = wl_oracle.histogram(ids[l], *dims[l])
        assert np.array_equal(phi, z["phi_h3_l%d" % l][:, :phi.shape[1]])
    # the fast restatement yields the same partition (column multiset)
    fast = wl_oracle.wl_features(train, 3, fast=True)
    for l in range(4):
        a = sorted(map(tuple, wl_oracle.histogram(ids[l], *dims[l]).T.tolist()))
        b = sorted(map(tuple, fast[l].T.tolist()))
        assert a == b


@pytest.mark.parametrize("oa", [False, True])
def test_joint_gram_and_pool_features(family, oa):
    _, z, train, pool = family
    joint = CellBatch.concat([train, pool])
    ref = z["Kraw_joint_h2_oa%d" % int(oa)]
    assert np.array_equal(wl_oracle.gram_raw(joint, 2, oa), ref)
    n = len(train)
    K_s, kss, ktt = wl_oracle.pool_features(train, pool, 2, oa)
    assert np.array_equal(K_s, ref[:n, n:])
    assert np.array_equal(kss, np.diag(ref)[n:])
    assert np.array_equal(ktt, np.diag(ref)[:n])


@pytest.mark.parametrize("opt", [False, True])
@pytest.mark.parametrize("oa", [False, True])
def test_gp_fit_predict_acquisitions(family, oa, opt):
    """fp64 restatement vs the fp32 reference: h*, lambda, incumbent and top-5 exact; the
    rest to fp32 tolerance (sigma^2 only to ~1e-2: SURVEY H2)."""
    _, z, train, pool = family
    tag = "oa%d_opt%d" % (int(oa), int(opt))
    cands = tuple(int(c) for c in z["gp_cands_" + tag])
    st = gp_oracle.gp_fit(train, z["y"].astype(np.float64), h_candidates=cands, oa=oa,
                          optimize_lik=opt, max_lik=0.01)
    np.testing.assert_allclose(st.nlml_grid, z["gp_grid_" + tag], rtol=2e-5)
    assert st.h == int(z["gp_h_" + tag])
    assert st.lik == float(z["gp_lik_" + tag])
    np.testing.assert_allclose(st.K, z["gp_K_" + tag], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(st.logDetK, z["gp_logDetK_" + tag], rtol=1e-4, atol=1e-4)
    if opt:
        np.testing.assert_allclose(st.nlml, z["gp_nlml_" + tag], rtol=1e-4)
    mu, var = gp_oracle.gp_predict_diag(st, train, pool)
    np.testing.assert_allclose(mu, z["gp_mu_" + tag], rtol=2e-3, atol=2e-3)
    # the reference forms 1+lik - k^T K_i k in fp32 with |K_i| ~ 1/lik: absolute noise ~1e-3*y_std^2
    np.testing.assert_allclose(var, np.diag(z["gp_cov_" + tag]), rtol=5e-2, atol=1e-3)
    mu_f, cov_f = gp_oracle.gp_predict_full(st, train, pool)
    np.testing.assert_allclose(mu_f, mu, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(np.diag(cov_f), var, rtol=1e-7, atol=1e-12)
    np.testing.assert_allclose(cov_f, z["gp_cov_" + tag], rtol=5e-2, atol=2e-3)
    inc = gp_oracle.incumbent(st)
    assert inc == pytest.approx(float(z["incumbent_" + tag]), rel=1e-6)
    ei = gp_oracle.expected_improvement(mu, var, inc)
    np.testing.assert_allclose(ei, z["ei_" + tag], rtol=5e-2, atol=2e-3)
    if opt:
        aei = gp_oracle.expected_improvement(mu, var, inc, augmented=True, lik=st.lik)
        np.testing.assert_allclose(aei, z["aei_" + tag], rtol=5e-2, atol=2e-3)
        ucb = gp_oracle.upper_confidence_bound(mu, var, gp_oracle.ucb_beta(1))
        np.testing.assert_allclose(ucb, z["ucb_" + tag], rtol=5e-3, atol=5e-3)
        top = gp_oracle.top_n_distinct(ei, 5)
        ref_top = set(int(i) for i in z["top5_" + tag])
        # compare as sets; allow a swap only between candidates whose reference EI differs
        # by less than the fp32 noise of the reference itself
        if set(int(i) for i in top) != ref_top:
            ref_ei = z["ei_" + tag]
            diff = set(int(i) for i in top) ^ ref_top
            vals = [ref_ei[i] for i in diff]
            assert max(vals) - min(vals) < 1e-4 * max(1.0, abs(max(vals)))


def test_cfg1_regression_case():
    """BASELINE.json configs[0] (nas_regression.py:78-83): n_train=50, 400 test cells, default h grid,
    optimize_lik=True — oracle vs the unmodified reference."""
    z, batch = load_golden("cfg1_nb101")
    n = int(z["n_train"])
    train, test = batch[:n], batch[n:]
    st = gp_oracle.gp_fit(train, z["y"].astype(np.float64), h_candidates=tuple(range(5)), optimize_lik=True)
    assert st.h == int(z["gp_h"]) and st.lik == float(z["gp_lik"])
    np.testing.assert_allclose(st.nlml, z["gp_nlml"], rtol=1e-4)
    mu, var = gp_oracle.gp_predict_diag(st, train, test)
    np.testing.assert_allclose(mu, z["gp_mu"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(var, z["gp_var"], rtol=5e-2, atol=1e-3)
    inc = gp_oracle.incumbent(st)
    assert inc == pytest.approx(float(z["incumbent"]), rel=1e-6)
    ei = gp_oracle.expected_improvement(mu, var, inc)
    np.testing.assert_allclose(ei, z["ei"], rtol=5e-2, atol=2e-3)
    top = set(int(i) for i in gp_oracle.top_n_distinct(ei, 5))
    ref_top = set(int(i) for i in z["top5"])
    if top != ref_top:
        vals = [z["ei"][i] for i in top ^ ref_top]
        assert max(vals) - min(vals) < 1e-4 * max(1.0, abs(max(vals)))


@pytest.mark.parametrize("name", ["cfg2_nb201_n150", "cfg3_nb101_n500", "cfg4_darts_n1000"])
def test_baseline_training_set_sizes(name):
    """BASELINE.json configs[1] / configs[2] at their TRAINING-SET sizes (NB201 n=150 with the h grid 0..5;
    NB101 n=500 at h=2) — oracle vs the unmodified reference (make_golden.py::baseline_size_case): integer Gram
    bit-exact, marginal-likelihood grid and chosen depth, fitted noise, batched posterior over 400 candidates, and
    the reference's own per-candidate EI / propose_location top-5 on a slice.  The reference's posterior is fp32
    (K_i, K_s: gp.py:271-276), which sets the tolerances on mu / var / EI."""
    z, batch = load_golden(name)
    n = int(z["n_train"])
    train, pool = batch[:n], batch[n:]
    cands = tuple(int(c) for c in z["cands"])
    st = gp_oracle.gp_fit(train, z["y"].astype(np.float64), h_candidates=cands, optimize_lik=True, max_lik=0.01)
    assert st.h == int(z["gp_h"]) and st.lik == float(z["gp_lik"])
    assert np.array_equal(wl_oracle.gram_raw(train, st.h), golden_kraw(z))
    np.testing.assert_allclose(np.array(st.nlml_grid), z["gp_grid"], rtol=1e-5)
    np.testing.assert_allclose(st.nlml, z["gp_nlml"], rtol=1e-5)
    np.testing.assert_allclose(st.logDetK, z["gp_logDetK"], rtol=1e-5)
    mu, var = gp_oracle.gp_predict_diag(st, train, pool)
    np.testing.assert_allclose(mu, z["gp_mu"], rtol=2e-3, atol=1e-2)
    np.testing.assert_allclose(var, z["gp_var"], rtol=5e-2, atol=5e-3)
    inc = gp_oracle.incumbent(st)
    assert inc == pytest.approx(float(z["incumbent"]), rel=1e-6)
    ne = int(z["n_eval"])
    ei = gp_oracle.expected_improvement(mu[:ne], var[:ne], inc)
    np.testing.assert_allclose(ei, z["ei"], rtol=5e-2, atol=2e-3)
    assert set(int(i) for i in gp_oracle.top_n_distinct(ei, 5)) == set(int(i) for i in z["top5"])


def _dmu_fixture(family):
    from nasbowl_b200.cells import CellBatch
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dmu_dphi.npz"))
    batch = CellBatch.from_records(z[family + "_records"], int(z[family + "_n_max"]))
    n = len(z[family + "_y"])
    return z, batch[:n], batch[n:], list(z[family + "_op_names"])


@pytest.mark.parametrize("oa", [False, True])
@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201", "darts"])
def test_dmu_dphi_matches_reference(fam, oa):
    """Interpretability gradients (gp.py:296-476): embeddings bit-exact, raw Jacobian-vector
    products and the three averaged forms at fp32 tolerance (the reference's K_i is fp32)."""
    z, train, pool, names = _dmu_fixture(fam)
    tag = "%s_oa%d_" % (fam, int(oa))
    h = int(z[tag + "h"])
    st = gp_oracle.gp_fit(train, z[fam + "_y"].astype(np.float64), h_candidates=(1, 2), oa=oa, optimize_lik=True,
                          max_lik=0.01)
    assert st.h == h and abs(st.lik - float(z[tag + "lik"])) < 1e-7
    X, Y = wl_oracle.wl_embed_reference_order(train, pool, h, names, layout="forward_t")
    assert np.array_equal(Y, z[tag + "feat"])
    # cells without labels outside the training vocabulary: the aligned layout is the same thing
    Xp, Yp = wl_oracle.wl_embed_reference_order(train, pool, h, names, layout="padded")
    _, Yt = wl_oracle.wl_embed_reference_order(train, pool, h, names, layout="trimmed")
    clean = Yt.sum(axis=1) == Y.sum(axis=1)
    assert np.array_equal(Xp, X) and np.array_equal(Yp[clean], Y[clean])
    g = gp_oracle.dmu_dphi(st, X, Y)
    np.testing.assert_allclose(g, z[tag + "raw"], rtol=2e-3, atol=2e-5)
    for nm, avg in (("occ", False), ("avg", True)):
        a, v, inc = gp_oracle.get_grad(g, Y, avg)
        np.testing.assert_allclose(a, z[tag + nm + "_mu"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(v, z[tag + nm + "_var"], rtol=5e-3, atol=2e-5, equal_nan=True)
        assert np.array_equal(inc, z[tag + nm + "_inc"])
    a, v, inc = gp_oracle.get_grad(gp_oracle.dmu_dphi(st, X, X), X, False)
    np.testing.assert_allclose(a, z[tag + "train_mu"], rtol=2e-3, atol=2e-5)
    assert np.array_equal(inc, z[tag + "train_inc"])


@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201"])
def test_multi_kernel_trainable_weights(fam):
    """Sum of two WL kernels with trained weights (gp.py:139-212): depth per kernel, weights, noise,
    K of the last iteration, K_i, predictions.  Three SGD steps pin the gradient values (Adam's
    sign-driven steps run into the [0, 1] clamp)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "multi_kernel.npz"))
    batch = CellBatch.from_records(z[fam + "_records"], int(z[fam + "_n_max"]))
    y = z[fam + "_y"].astype(np.float64)
    n = len(y)
    train, pool = batch[:n], batch[n:]
    kernels = [(1, False), (2, True)]
    st = gp_oracle.gp_fit_sum(train, y, kernels, h_candidates=(0, 1, 2), optimize_lik=True, max_lik=0.01)
    assert st.hs == list(z[fam + "_hs"])
    np.testing.assert_allclose(st.weights, z[fam + "_weights"], atol=1e-6)
    assert abs(st.lik - float(z[fam + "_lik"])) < 1e-7
    np.testing.assert_allclose(st.nlml, float(z[fam + "_nlml"]), rtol=2e-3)
    np.testing.assert_allclose(st.K, z[fam + "_K"], rtol=1e-4, atol=1e-5)
    np.testing.assert_allclose(st.K_i, z[fam + "_Ki"], rtol=5e-3, atol=5e-3 * np.abs(z[fam + "_Ki"]).max())
    mu, cov = gp_oracle.gp_predict_full_sum(st, train, pool)
    np.testing.assert_allclose(mu, z[fam + "_mu"], rtol=2e-3, atol=2e-3)
    np.testing.assert_allclose(cov, z[fam + "_cov"], rtol=5e-3, atol=1e-4)
    sg = gp_oracle.gp_fit_sum(train, y, kernels, h_candidates=(), optimize_lik=True, max_lik=0.01, iters=3, lr=0.02,
                              optimizer="sgd")
    np.testing.assert_allclose(sg.weights, z[fam + "_sgd_weights"], rtol=1e-4)
    np.testing.assert_allclose(sg.nlml, float(z[fam + "_sgd_nlml"]), rtol=2e-3)


@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201"])
def test_trainable_wl_layer_weights(fam):
    """fit(optimize_wl_layer_weights=True) on one WL kernel (gp.py:157-165,208,218)."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "multi_kernel.npz"))
    batch = CellBatch.from_records(z[fam + "_records"], int(z[fam + "_n_max"]))
    y = z[fam + "_y"].astype(np.float64)
    n = len(y)
    train, pool = batch[:n], batch[n:]
    for tag, kw in (("lw", {}), ("lwsgd", dict(iters=3, lr=0.02, optimizer="sgd"))):
        st = gp_oracle.gp_fit_sum(train, y, [(2, False)], h_candidates=(), optimize_lik=True, max_lik=0.01,
                                  optimize_layer_weights=True, **kw)
        np.testing.assert_allclose(st.layer_weights, z[fam + "_%s_layer_weights" % tag], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(st.nlml, float(z[fam + "_%s_nlml" % tag]), rtol=2e-3)
        mu, cov = gp_oracle.gp_predict_full_sum(st, train, pool)
        np.testing.assert_allclose(mu, z[fam + "_%s_mu" % tag], rtol=2e-3, atol=2e-3)
        np.testing.assert_allclose(np.diag(cov), z[fam + "_%s_var" % tag], rtol=5e-3, atol=1e-4)
