"""GPU tests of the reference-shaped API corners: undirected WL, per-level weights, several WL
kernels with fixed weights (dense predict path), embeddings, pickling, mixed record widths."""
import pickle

import numpy as np
import pytest
import torch

from conftest import load_golden
from nasbowl_b200.cells import CellBatch
from oracle import gp_oracle, wl_oracle

pytestmark = pytest.mark.gpu


def test_undirected_and_layer_weights(family):
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, _ = family
    und = train.to_undirected()
    K = WeisfilerLehman(h=2, undirected=True).fit_transform(train, rebuild_model=True)
    assert np.array_equal(K.numpy(), wl_oracle.gram_raw(und, 2))
    assert not np.array_equal(K.numpy(), z["Kraw_h2_oa0"])
    lw = np.array([1.0, 0.5, 2.0])
    Kw = WeisfilerLehman(h=2, layer_weights=lw).fit_transform(train, rebuild_model=True)
    levels = wl_oracle.gram_raw_levels(train, 2)
    np.testing.assert_allclose(Kw.numpy(), (levels * lw[:, None, None]).sum(0), rtol=1e-14)
    # weights of the wrong length are reset to ones (weisfeiler_lehman.py:126-127)
    K1 = WeisfilerLehman(h=1, layer_weights=lw).fit_transform(train, rebuild_model=True)
    assert np.array_equal(K1.numpy(), z["Kraw_h1_oa0"])
    # cached Gram is returned unless rebuild_model (weisfilerlehman.py:125-126)
    k = WeisfilerLehman(h=2)
    a = k.fit_transform(train, rebuild_model=True)
    assert k.fit_transform(train[:5]) is k._gram and np.array_equal(a.numpy(), k._gram.numpy())


def test_embeddings_restricted_to_fitted_vocabulary(family):
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    k = WeisfilerLehman(h=2)
    k.fit_transform(train, rebuild_model=True)
    emb = k.transform(pool)
    assert len(emb) == 3 and all(e.shape[0] == len(pool) for e in emb)
    phi_train = [e.numpy().astype(np.int64) for e in k.transform(train)]
    K_s = sum(pt @ e.numpy().astype(np.int64).T for pt, e in zip(phi_train, emb))
    assert np.array_equal(K_s, wl_oracle.pool_features(train, pool, 2)[0])


def test_two_kernels_fixed_weights_dense_predict(family):
    """SumKernel of two WL kernels with fixed weights (combine_kernels.py:36-56) through
    GraphGP.predict: K = normalise(w1 K1 + w2 K2) on train ∪ test."""
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    y = z["y"].astype(np.float64)
    w = [0.3, 0.7]
    gp = GraphGP(train, torch.tensor(y), [WeisfilerLehman(h=1), WeisfilerLehman(h=2, oa=True)], weights=w)
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    n, M = len(train), len(pool)
    joint = CellBatch.concat([train, pool])
    wf = [gp_oracle.f32(v) for v in w]      # the reference keeps kernel weights as an fp32 tensor (gp.py:66)
    Kj = gp_oracle.normalize_gram(wf[0] * wl_oracle.gram_raw(joint, 1) + wf[1] * wl_oracle.gram_raw(joint, 2, oa=True))
    yn, ym, ys = gp_oracle.normalize_y(y)
    lik = gp_oracle.f32(1e-3)
    K_i, _, _ = gp_oracle.pd_inverse(Kj[:n, :n], lik)
    K_s = Kj[:n, n:]
    mu_o = (K_s.T @ K_i @ yn) * ys + ym
    cov_o = (np.sqrt(np.maximum(Kj[n:, n:] + lik * np.eye(M) - K_s.T @ K_i @ K_s, lik)) * ys) ** 2
    mu, cov = gp.predict(pool)
    np.testing.assert_allclose(gp.K.numpy(), Kj[:n, :n], rtol=1e-12)
    np.testing.assert_allclose(mu.numpy(), mu_o, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(cov.numpy(), cov_o, rtol=1e-6, atol=1e-12)
    assert torch.allclose(gp.weights.double(), torch.tensor(w, dtype=torch.float64))
    with pytest.raises(NotImplementedError):
        gp.predict_diag(pool)                  # the fused pool path is single-kernel


def test_pickle_roundtrip_and_refit(family):
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    gp = GraphGP(train, torch.tensor(z["y"]), [WeisfilerLehman(h=2, oa=True)])
    gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True)
    mu0, var0 = gp.predict_diag(pool)
    gp2 = pickle.loads(pickle.dumps({"gp": gp, "x": train}))["gp"]        # run_darts.py:110-125
    assert gp2.kernels[0].h == gp.kernels[0].h and gp2.likelihood == gp.likelihood
    with pytest.raises(ValueError):
        gp2.predict_diag(pool)                                              # device state does not travel
    gp2.reset_XY(train, torch.tensor(z["y"]))
    gp2.fit(wl_subtree_candidates=(1, 2), optimize_lik=True)               # what the resume path does
    mu1, var1 = gp2.predict_diag(pool)
    np.testing.assert_allclose(mu1.numpy(), mu0.numpy(), rtol=1e-12)
    np.testing.assert_allclose(var1.numpy(), var0.numpy(), rtol=1e-12)
    assert GraphExpectedImprovement(gp2).propose_location(pool, top_n=3)[2].shape == (3,)


def test_small_cells_scored_by_a_gp_fitted_on_wide_records():
    """NB cells (8-node records) scored by a surrogate fitted on 16-node records, and the reverse."""
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    z, batch = load_golden("nasbench101")
    train, pool = batch[:40], batch[40:]
    y = torch.tensor(z["y"])
    g8 = GraphGP(train, y, [WeisfilerLehman(h=2)])
    g8.fit(wl_subtree_candidates=(2,), optimize_lik=False)
    g16 = GraphGP(train, y, [WeisfilerLehman(h=2)], n_max=16)
    g16.fit(wl_subtree_candidates=(2,), optimize_lik=False)
    a, b = g8.predict_diag(pool), g16.predict_diag(pool)
    np.testing.assert_allclose(a[0].numpy(), b[0].numpy(), rtol=1e-12)
    np.testing.assert_allclose(a[1].numpy(), b[1].numpy(), rtol=1e-12)
    # the reverse: a pool with wider cells than the training set re-derives the state on 16-node records
    # (the reference accepts cells of any size); both surrogates then agree
    zd, darts = load_golden("darts")
    c, d = g8.predict_diag(darts[:4]), g16.predict_diag(darts[:4])
    assert g8._dev['n_max'] == 16 and c[0].shape == (4,)
    np.testing.assert_allclose(c[0].numpy(), d[0].numpy(), rtol=1e-12)
    np.testing.assert_allclose(c[1].numpy(), d[1].numpy(), rtol=1e-12)


def test_feature_map_and_feature_value_match_reference(family):
    """Interpretability accessors (row f3): the motif strings and their numbering, and the
    embedding of unseen cells over them, exactly as the reference produces them
    (kernels/weisfilerlehman.py:219-266, weisfeiler_lehman.py:577-599)."""
    from nasbowl_b200 import OpVocab
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    for as_graphs in (False, True):
        k = WeisfilerLehman(h=3, requires_grad=True)
        if as_graphs:
            vocab = OpVocab(list(z["op_names"]))
            train_in, pool_in = train.to_networkx(vocab), pool.to_networkx(vocab)
        else:
            k.vocab = OpVocab(list(z["op_names"]))      # packed records carry the fixture's op codes
            train_in, pool_in = train, pool
        k.fit_transform(train_in, rebuild_model=True)
        emb, names = k.feature_value(pool_in)
        fm = k.feature_map(flatten=True)
        assert list(fm.keys()) == list(z["featmap_h3_ids"])
        assert list(fm.values()) == list(z["featmap_h3_names"]) == names
        layered = k.feature_map(flatten=False)
        assert [len(layered[l]) for l in range(4)] == list(np.diff(z["feature_dims_h3"]))
        assert np.array_equal(emb.numpy().astype(np.int64), z["featval_h3_pool"].astype(np.int64))
        # the training embedding reproduces the reference's ordered histograms column for column
        emb_t, _ = k.feature_value(train_in)
        ref = np.concatenate([z["phi_h3_l%d" % l][:, :w] for l, w in enumerate(np.diff(z["feature_dims_h3"]))], axis=1)
        assert np.array_equal(emb_t.numpy().astype(np.int64), ref.astype(np.int64))
    # oa: same counts (the base kernel changes the Gram, not the histogram)
    k = WeisfilerLehman(h=3, oa=True)
    k.vocab = OpVocab(list(z["op_names"]))
    k.fit_transform(train, rebuild_model=True)
    emb_oa, _ = k.feature_value(train)
    assert np.array_equal(emb_oa.numpy().astype(np.int64), ref.astype(np.int64))


def test_bo_loop_matches_oracle_every_iteration():
    """The reference's outer loop (nasbench_optimisation.py:153-216) on the GPU path: at every
    iteration the fitted depth, the noise and the proposed batch agree with the CPU oracle."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import bo_loop

    def check(gp, x, y, pool, eis, idx):
        st = gp_oracle.gp_fit(x, y.numpy(), h_candidates=(1, 2, 3), optimize_lik=True, max_lik=0.01)
        assert st.h == gp.kernels[0].h and np.float32(st.lik) == np.float32(gp.likelihood)
        sub = CellBatch.from_records(pool.cpu().numpy(), 8)
        mu, var = gp_oracle.gp_predict_diag(st, x, sub)
        ei = gp_oracle.expected_improvement(mu, var, gp_oracle.incumbent(st))
        np.testing.assert_allclose(eis.reshape(-1), ei, rtol=1e-4, atol=1e-9)
        assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(eis.reshape(-1), 5))

    best, x, y = bo_loop.run(iters=4, pool_size=3000, batch_size=5, n_init=20, verbose=False, check=check)
    assert len(x) == 40 and best[-1] >= best[0]
    # the reference's default pool strategy: half mutated from the best cells so far, half random
    best, x, y = bo_loop.run(iters=3, pool_size=3000, batch_size=5, n_init=20, verbose=False, check=check,
                             strategy="mutate")
    assert len(x) == 35 and best[-1] >= best[0]


@pytest.mark.gpu
@pytest.mark.parametrize("oa", [False, True])
@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201", "darts"])
def test_dmu_dphi_matches_reference_and_oracle(fam, oa):
    """GraphGP.dmu_dphi / forward_t / get_grad (gp.py:296-476).  Against the reference's golden
    vectors on every cell inside the training vocabulary (incl. the training-set call all of the
    reference's consumers make) and against the oracle's aligned form on every cell."""
    import os
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.cells import CellBatch, OpVocab
    from nasbowl_b200.kernels import WeisfilerLehman
    from oracle import gp_oracle, wl_oracle
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "dmu_dphi.npz"))
    batch = CellBatch.from_records(z[fam + "_records"], int(z[fam + "_n_max"]))
    names = list(z[fam + "_op_names"])
    y = z[fam + "_y"]
    n = len(y)
    train, pool = batch[:n], batch[n:]
    tag = "%s_oa%d_" % (fam, int(oa))
    k = WeisfilerLehman(h=2, oa=oa, requires_grad=True)
    k.vocab = OpVocab(names)
    gp = GraphGP(train, torch.tensor(y), [k], n_max=batch.n_max)
    gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True, max_lik=0.01)
    h = int(z[tag + "h"])
    assert gp.kernels[0].h == h
    raw, _, feat = gp.dmu_dphi(pool, average_across_features=False)
    raw = torch.cat(raw).numpy()
    # --- oracle, aligned layout, every pool cell
    st = gp_oracle.gp_fit(train, y.astype(np.float64), h_candidates=(1, 2), oa=oa, optimize_lik=True, max_lik=0.01)
    X, Y = wl_oracle.wl_embed_reference_order(train, pool, h, names, layout="padded")
    assert np.array_equal(feat.numpy(), Y)
    np.testing.assert_allclose(raw, gp_oracle.dmu_dphi(st, X, Y), rtol=1e-6, atol=1e-9)
    # --- reference golden vectors on the cells without unknown labels
    # (NB201 has few distinct cells, so its pool has many; the sparser families are covered by the
    # training-set call below, where no label can be unknown)
    clean = np.all(z[tag + "feat"] == Y, axis=1)
    assert clean.sum() >= 3 or fam != "nasbench201"
    np.testing.assert_allclose(raw[clean], z[tag + "raw"][clean], rtol=2e-3, atol=2e-5)
    # --- the consumers' call: training cells, averaged per occurrence count
    a, v, inc = gp.dmu_dphi()
    np.testing.assert_allclose(a.numpy(), z[tag + "train_mu"], rtol=2e-3, atol=2e-5)
    np.testing.assert_allclose(v.numpy(), z[tag + "train_var"], rtol=2e-2, atol=2e-5, equal_nan=True)
    assert np.array_equal(inc.numpy(), z[tag + "train_inc"])
    a2, v2, tot = gp.dmu_dphi(average_across_occurrences=True)
    ao, vo, to = gp_oracle.get_grad(gp_oracle.dmu_dphi(st, X, X), X, True)
    np.testing.assert_allclose(a2.numpy(), ao, rtol=1e-4, atol=1e-6)
    assert np.array_equal(tot.numpy(), to)
    # --- forward_t: autograd through the returned graph gives the same Jacobian-vector product
    K, y_, x_ = k.forward_t(pool[:2])
    alpha = torch.tensor(st.K_i @ st.y_norm).reshape(1, -1)
    jv = torch.autograd.grad(outputs=K[:1], inputs=y_, grad_outputs=alpha)[0][0].numpy()
    np.testing.assert_allclose(jv, raw[0], rtol=1e-5, atol=1e-8)


@pytest.mark.gpu
@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201"])
def test_multi_kernel_trainable_weights(fam):
    """GraphGP over a sum of two WL kernels: the kernel weights are trained with the noise
    (gp.py:139-212) from closed-form device gradients (nbw_kernel_weight_grad)."""
    import os
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.cells import CellBatch, OpVocab
    from nasbowl_b200.kernels import WeisfilerLehman
    from oracle import gp_oracle
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "multi_kernel.npz"))
    batch = CellBatch.from_records(z[fam + "_records"], int(z[fam + "_n_max"]))
    y = z[fam + "_y"]
    n = len(y)
    train, pool = batch[:n], batch[n:]

    def kernels():
        ks = [WeisfilerLehman(h=1, oa=False), WeisfilerLehman(h=2, oa=True)]
        for k in ks:
            k.vocab = OpVocab(list(z[fam + "_op_names"]))
        return ks
    gp = GraphGP(train, torch.tensor(y), kernels(), n_max=batch.n_max)
    gp.fit(wl_subtree_candidates=(0, 1, 2), optimize_lik=True, max_lik=0.01)
    st = gp_oracle.gp_fit_sum(train, y.astype(np.float64), [(1, False), (2, True)], h_candidates=(0, 1, 2),
                              optimize_lik=True, max_lik=0.01)
    assert [k.h for k in gp.kernels] == st.hs == list(z[fam + "_hs"])
    np.testing.assert_allclose(gp.weights.numpy(), st.weights, atol=1e-6)
    np.testing.assert_allclose(gp.weights.numpy(), z[fam + "_weights"], atol=1e-6)
    assert abs(gp.likelihood - st.lik) < 1e-9
    np.testing.assert_allclose(gp.K.numpy(), st.K, rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(gp.K_i.numpy(), st.K_i, rtol=1e-6, atol=1e-6 * np.abs(st.K_i).max())
    np.testing.assert_allclose(float(gp.nlml), st.nlml, rtol=1e-8)
    mu, cov = gp.predict(pool)
    mo, co = gp_oracle.gp_predict_full_sum(st, train, pool)
    np.testing.assert_allclose(mu.numpy(), mo, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(cov.numpy(), co, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(mu.numpy(), z[fam + "_mu"], rtol=2e-3, atol=2e-3)
    # gradient values: three small SGD steps
    gp2 = GraphGP(train, torch.tensor(y), kernels(), n_max=batch.n_max)
    gp2.fit(iters=3, optimizer='sgd', optimizer_kwargs={'lr': 0.02}, wl_subtree_candidates=(), optimize_lik=True,
            max_lik=0.01)
    np.testing.assert_allclose(gp2.weights.numpy(), z[fam + "_sgd_weights"], rtol=1e-4)
    # fixed weights are left alone
    gp3 = GraphGP(train, torch.tensor(y), kernels(), weights=[0.3, 0.7], n_max=batch.n_max)
    gp3.fit(wl_subtree_candidates=(), optimize_lik=False)
    np.testing.assert_allclose(gp3.weights.numpy(), [0.3, 0.7], rtol=1e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("fam", ["nasbench101", "nasbench201"])
def test_trainable_wl_layer_weights(fam):
    """fit(optimize_wl_layer_weights=True) (gp.py:157-165,208,218): per-level weights of the WL kernel
    trained with the noise; predict() uses the final layer weights with the K_i of the last iteration."""
    import os
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.cells import CellBatch, OpVocab
    from nasbowl_b200.kernels import WeisfilerLehman
    from oracle import gp_oracle
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "multi_kernel.npz"))
    batch = CellBatch.from_records(z[fam + "_records"], int(z[fam + "_n_max"]))
    y = z[fam + "_y"]
    n = len(y)
    train, pool = batch[:n], batch[n:]
    for tag, kw, okw in (("lw", {}, {}),
                         ("lwsgd", dict(iters=3, optimizer='sgd', optimizer_kwargs={'lr': 0.02}),
                          dict(iters=3, lr=0.02, optimizer="sgd"))):
        k = WeisfilerLehman(h=2, oa=False)
        k.vocab = OpVocab(list(z[fam + "_op_names"]))
        gp = GraphGP(train, torch.tensor(y), [k], n_max=batch.n_max)
        gp.fit(wl_subtree_candidates=(), optimize_lik=True, max_lik=0.01, optimize_wl_layer_weights=True, **kw)
        st = gp_oracle.gp_fit_sum(train, y.astype(np.float64), [(2, False)], h_candidates=(), optimize_lik=True,
                                  max_lik=0.01, optimize_layer_weights=True, **okw)
        np.testing.assert_allclose(gp.layer_weights.numpy(), st.layer_weights, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(gp.layer_weights.numpy(), z[fam + "_%s_layer_weights" % tag], rtol=1e-4, atol=1e-6)
        # the optimiser state is fp32 (torch here, numpy in the oracle): agreement to ~1e-5
        np.testing.assert_allclose(gp.K.numpy(), st.K, rtol=1e-4, atol=1e-7)
        mu, cov = gp.predict(pool)
        mo, co = gp_oracle.gp_predict_full_sum(st, train, pool)
        np.testing.assert_allclose(mu.numpy(), mo, rtol=1e-3, atol=1e-4)
        np.testing.assert_allclose(cov.numpy(), co, rtol=1e-3, atol=1e-6)
        np.testing.assert_allclose(mu.numpy(), z[fam + "_%s_mu" % tag], rtol=2e-3, atol=2e-3)
        # the fused pool path carries the trained layer weights (nbw_model.level_weight)
        mu_d, var_d = gp.predict_diag(pool)
        np.testing.assert_allclose(mu_d.numpy(), mu.numpy(), rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(var_d.numpy(), np.diag(cov.numpy()), rtol=1e-8, atol=1e-13)


@pytest.mark.gpu
def test_likelihood_only_factorisation_matches_full():
    """nbw_gp_factor without the inverse (what the h grid search uses per candidate) gives the same
    log-determinant and y^T K^-1 y as the full factorisation and as numpy."""
    from nasbowl_b200.engine import Engine
    eng = Engine.get()
    rng = np.random.RandomState(0)
    for n in (1, 5, 32, 33, 150, 700):
        A = rng.randn(n, n + 3)
        K = A @ A.T / (n + 3) + 0.05 * np.eye(n)
        y = rng.randn(n)
        Kd, yd = torch.tensor(K, device="cuda"), torch.tensor(y, device="cuda")
        full = eng.gp_factor(Kd, 0.01, yd)
        lite = eng.gp_factor(Kd, 0.01, yd, want_inverse=False)
        assert lite[1] is None and lite[2] is None
        Kl = K + 0.01 * np.eye(n)
        np.testing.assert_allclose(lite[3], np.linalg.slogdet(Kl)[1], rtol=1e-10, atol=1e-10)
        np.testing.assert_allclose(lite[4], y @ np.linalg.solve(Kl, y), rtol=1e-9)
        np.testing.assert_allclose([lite[3], lite[4]], [full[3], full[4]], rtol=1e-11, atol=1e-12)
        assert torch.equal(full[0], lite[0])


@pytest.mark.gpu
def test_blocked_cholesky_and_inverse_at_block_boundaries():
    """L L^T = K + lik I, Linv L = I and the NotPositiveDefinite path, at sizes around the 32-wide
    panels (the diagonal block is factored warp-synchronously inside the panel kernel)."""
    from nasbowl_b200 import _lib
    from nasbowl_b200.engine import Engine
    eng = Engine.get()
    rng = np.random.RandomState(1)
    for n in (1, 2, 31, 32, 33, 64, 65, 150, 257, 1000):
        A = rng.randn(n, n + 2)
        Kh = A @ A.T / (n + 2) + 0.1 * np.eye(n)
        K = torch.tensor(Kh, device="cuda")
        y = torch.tensor(rng.randn(n), device="cuda")
        L, Linv, alpha, logdet, yKy, trKinv, a2 = eng.gp_factor(K, 0.01, y)
        Lh, Lih = L.cpu().numpy(), Linv.cpu().numpy()
        Kl = Kh + 0.01 * np.eye(n)
        assert np.array_equal(np.triu(Lh, 1), np.zeros_like(Lh))
        np.testing.assert_allclose(Lh @ Lh.T, Kl, rtol=1e-12, atol=1e-12)
        np.testing.assert_allclose(Lh, np.linalg.cholesky(Kl), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(Lih @ Lh, np.eye(n), atol=1e-9)
        np.testing.assert_allclose(alpha.cpu().numpy(), np.linalg.solve(Kl, y.cpu().numpy()), rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(trKinv, np.trace(np.linalg.inv(Kl)), rtol=1e-9)
    bad = torch.tensor(np.array([[1.0, 2.0], [2.0, 1.0]]), device="cuda")
    with pytest.raises(_lib.NotPositiveDefinite):
        eng.gp_factor(bad, 0.0, torch.zeros(2, dtype=torch.float64, device="cuda"))
    bad = torch.tensor(Kh - 5.0 * np.eye(1000), device="cuda")        # breaks down in a later panel
    with pytest.raises(_lib.NotPositiveDefinite):
        eng.gp_factor(bad, 0.0, torch.zeros(1000, dtype=torch.float64, device="cuda"))


@pytest.mark.gpu
@pytest.mark.parametrize("fam,n0,h_grid", [(0, 60, (1, 2)), (1, 150, tuple(range(6))), (2, 120, (2,))])
def test_incremental_refit_matches_from_scratch(fam, n0, h_grid):
    """a BO loop appends a batch per iteration (nasbench_optimisation.py:209-213) and refits.
    reset_XY + fit on examples that EXTEND the previous ones appends to the previous fit — append-only
    dictionary, rank-k Cholesky / inverse append, rank-k update of the pool operators — and must agree with a
    from-scratch fit at 1e-9, iteration after iteration."""
    from nasbowl_b200 import synth
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.cells import CellBatch
    from nasbowl_b200.kernels import WeisfilerLehman
    total = n0 + 4 * 5
    cells = synth.generate_cells_host(fam, 8, 0, total)
    y_all = torch.randn(total, generator=torch.Generator().manual_seed(fam)).double()
    pool = synth.generate_cells_host(fam, 9, 5000, 3000)
    gp = GraphGP(cells[:n0], y_all[:n0], [WeisfilerLehman(h=2)], n_max=cells.n_max)
    gp.incremental_min_n = 0           # (by default only training sets of >= 768 cells are extended: below, refitting is faster)
    gp.fit(wl_subtree_candidates=h_grid, optimize_lik=True, max_lik=0.01)
    assert not gp.last_fit_incremental
    used = 0
    for it in range(1, 5):
        n = n0 + 5 * it
        gp.reset_XY(cells[:n], y_all[:n])
        gp.fit(wl_subtree_candidates=h_grid, optimize_lik=True, max_lik=0.01)
        used += int(gp.last_fit_incremental)
        ref = GraphGP(cells[:n], y_all[:n], [WeisfilerLehman(h=2)], n_max=cells.n_max)
        ref.incremental = False
        if it > 1:
            ref.likelihood = 0.01                  # the noise the loop carries over from the previous fit (Q5)
            ref.likelihood = float(np.float32(0.01))
        ref.fit(wl_subtree_candidates=h_grid, optimize_lik=True, max_lik=0.01)
        assert gp.kernels[0].h == ref.kernels[0].h and gp.likelihood == ref.likelihood
        np.testing.assert_allclose(float(gp.nlml), float(ref.nlml), rtol=1e-10)
        np.testing.assert_allclose(float(gp.logDetK), float(ref.logDetK), rtol=1e-10)
        np.testing.assert_allclose(gp.K.numpy(), ref.K.numpy(), rtol=0, atol=0)
        np.testing.assert_allclose(gp.K_i.numpy(), ref.K_i.numpy(), rtol=1e-9, atol=1e-9)
        mu_a, var_a = gp.predict_diag(pool)
        mu_b, var_b = ref.predict_diag(pool)
        np.testing.assert_allclose(mu_a.numpy(), mu_b.numpy(), rtol=1e-9, atol=1e-11)
        np.testing.assert_allclose(var_a.numpy(), var_b.numpy(), rtol=1e-9, atol=1e-12)
        ia = GraphExpectedImprovement(gp).propose_location(pool, top_n=5)[2]
        ib = GraphExpectedImprovement(ref).propose_location(pool, top_n=5)[2]
        assert set(int(i) for i in ia) == set(int(i) for i in ib)
        # same vocabulary sizes as the from-scratch dictionary, old labels kept their ids
        assert gp._dev['fits'][0].dims == ref._dev['fits'][0].dims
    assert used >= 3, "the incremental path was not taken"
    # examples that do not extend the old ones: from scratch again
    gp.reset_XY(cells[5:n0 + 5], y_all[5:n0 + 5])
    gp.fit(wl_subtree_candidates=h_grid, optimize_lik=True, max_lik=0.01)
    assert not gp.last_fit_incremental


def test_reference_output_dtypes():
    """out_dtype=torch.float32 hands out K, K_i, mu and cov in the reference's dtype (combine_kernels.py:36,
    utils.py:140, gp.py:271-283) — rounded from the fp64 evaluation, which is what the default returns."""
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    z, batch = load_golden("cfg1_nb101")
    n = int(z["n_train"])
    train, test = batch[:n], batch[n:n + 64]
    out = {}
    for dt in (None, torch.float32):
        gp = GraphGP(train, torch.tensor(z["y"]), [WeisfilerLehman(h=2)], out_dtype=dt)
        gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True, max_lik=0.01)
        out[dt] = (gp.K, gp.K_i) + tuple(gp.predict(test))
    for a, b in zip(out[None], out[torch.float32]):
        assert a.dtype == torch.float64 and b.dtype == torch.float32 and a.shape == b.shape
        assert torch.equal(a.to(torch.float32), b)
    # usable next to the reference's own fp32 tensors
    assert torch.cat([out[torch.float32][2], torch.zeros(3)]).dtype == torch.float32
