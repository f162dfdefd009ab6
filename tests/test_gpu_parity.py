"""GPU parity tests: the CUDA path (through the C-ABI) against the CPU oracle and the golden
vectors of the unmodified reference.  Integer / index work is compared bit-exactly; floating
point against the fp64 oracle at 1e-6 relative (fp64 outputs) and 1e-4 (fp32 scores), the
tolerances BASELINE.json's north_star states."""
import numpy as np
import pytest
import torch

from conftest import golden_kraw, load_golden
from nasbowl_b200 import synth
from nasbowl_b200.cells import NO_NODE, CellBatch
from oracle import gp_oracle, wl_oracle

pytestmark = pytest.mark.gpu

RTOL64 = 1e-6
RTOL32 = 1e-4


@pytest.fixture(scope="module")
def eng():
    from nasbowl_b200.engine import Engine
    return Engine.get()


def _sorted_columns(phi):
    phi = np.asarray(phi)
    cols = [tuple(c) for c in phi.T.tolist() if any(c)]
    return sorted(cols)


# ---------------------------------------------------------------------------- WL + Gram
@pytest.mark.parametrize("oa", [False, True])
@pytest.mark.parametrize("h", [0, 1, 2, 3, 5])
def test_raw_gram_bit_exact_vs_reference(eng, family, h, oa):
    _, z, train, _ = family
    fit = eng.wl_fit(eng.upload_cells(train), train.n_max, h, oa)
    _, dims = wl_oracle.wl_labels_fast(train, h)
    assert fit.dims == dims                                  # vocabulary sizes per level
    K = eng.gram(fit.phi, fit.phi, 0, fit.k_end(h)).cpu().numpy()
    assert np.array_equal(K, z["Kraw_h%d_oa%d" % (h, int(oa))])     # golden, from the reference
    K_simt = eng.gram(fit.phi, fit.phi, 0, fit.k_end(h), impl=1).cpu().numpy()
    assert np.array_equal(K, K_simt)
    assert np.array_equal(fit.self_sim.cpu().numpy(), np.diag(K))
    # prefix property: one fit at depth h serves every shallower depth (used by the h grid search)
    for hh in range(h):
        Kh = eng.gram(fit.phi, fit.phi, 0, fit.k_end(hh)).cpu().numpy()
        assert np.array_equal(Kh, wl_oracle.gram_raw(train, hh, oa))


def test_histograms_match_oracle_up_to_column_order(eng, family):
    _, z, train, _ = family
    fit = eng.wl_fit(eng.upload_cells(train), train.n_max, 3, False)
    phi = fit.phi.cpu().numpy()
    ref = wl_oracle.wl_features(train, 3)
    for l in range(4):
        block = phi[:, fit.col_off[l]:fit.col_off[l + 1]]
        assert _sorted_columns(block) == _sorted_columns(ref[l])
        assert not block[:, fit.widths[l]:].any()            # pad columns are zero


def test_joint_gram_and_pool_features(eng, family):
    _, z, train, pool = family
    n = len(train)
    from nasbowl_b200.pool_model import PoolModel
    for oa in (False, True):
        ref = z["Kraw_joint_h2_oa%d" % int(oa)]
        joint = CellBatch.concat([train, pool])
        fitj = eng.wl_fit(eng.upload_cells(joint), joint.n_max, 2, oa)
        assert np.array_equal(eng.gram(fitj.phi, fitj.phi, 0, fitj.k_end(2)).cpu().numpy(), ref)
        # pool side: features restricted to the training vocabulary + self-similarity over all labels
        fit = eng.wl_fit(eng.upload_cells(train), train.n_max, 2, oa)
        pm = PoolModel.from_fit(eng, fit, 2)     # owns the device tables the descriptor points into
        model = pm.descriptor()
        phi_p, ss = eng.pool_features(model, eng.upload_cells(pool), len(pool))
        D = fit.k_end(2)
        K_s = fit.phi.cpu().numpy()[:, :D].astype(np.int64) @ phi_p.cpu().numpy()[:, :D].astype(np.int64).T
        assert np.array_equal(K_s, ref[:n, n:])
        assert np.array_equal(ss.cpu().numpy(), np.diag(ref)[n:])
        # the training cells themselves must map onto their own rows
        phi_t, ss_t = eng.pool_features(model, eng.upload_cells(train), n)
        assert np.array_equal(phi_t.cpu().numpy()[:, :D], fit.phi.cpu().numpy()[:, :D])
        assert np.array_equal(ss_t.cpu().numpy(), np.diag(ref)[:n])


def test_gram_tcgen05_random_shapes(eng):
    rng = np.random.RandomState(0)
    for (m, n, k) in [(1, 1, 128), (50, 70, 256), (128, 128, 128), (129, 257, 384), (300, 40, 1280), (513, 515, 640)]:
        A = torch.from_numpy(rng.randint(-128, 128, size=(m, k)).astype(np.int8)).cuda()
        B = torch.from_numpy(rng.randint(-128, 128, size=(n, k)).astype(np.int8)).cuda()
        ref = A.cpu().numpy().astype(np.int64) @ B.cpu().numpy().astype(np.int64).T
        for k0, k1 in [(0, k), (128 * ((k // 128) // 2), k)]:
            r = A.cpu().numpy().astype(np.int64)[:, k0:k1] @ B.cpu().numpy().astype(np.int64)[:, k0:k1].T
            assert np.array_equal(eng.gram(A, B, k0, k1, impl=0).cpu().numpy(), r), (m, n, k, k0)
            assert np.array_equal(eng.gram(A, B, k0, k1, impl=1).cpu().numpy(), r), (m, n, k, k0)
        assert np.array_equal(eng.gram(A, B, 0, k).cpu().numpy(), ref)


# ---------------------------------------------------------------------------- dictionary build
def test_sort_unique_large(eng):
    import ctypes as C
    rng = np.random.RandomState(1)
    for n, distinct in [(1, 1), (37, 5), (4096, 100), (4097, 4097), (300000, 20000),
                        (10000, 9000), (70000, 65536), (300000, 100000), (2000000, 23493)]:
        pool = rng.randint(0, 2 ** 63 - 1, size=distinct, dtype=np.int64)
        keys = pool[rng.randint(0, distinct, size=n)]
        invalid = rng.rand(n) < 0.1
        keys_u = keys.astype(np.uint64)
        keys_u[invalid] = np.uint64(0xFFFFFFFFFFFFFFFF)
        chks = (keys_u * np.uint64(3)) ^ np.uint64(0x1234)
        dk = torch.from_numpy(keys_u.view(np.int64)).cuda()
        dc = torch.from_numpy(chks.view(np.int64)).cuda()
        ids = torch.empty(n, dtype=torch.int32, device="cuda")
        uk = torch.empty(n, dtype=torch.int64, device="cuda")
        uc = torch.empty(n, dtype=torch.int64, device="cuda")
        d = C.c_int64(0)
        eng._check(eng.lib.nbw_dict_build(eng.ctx, eng._p(dk), eng._p(dc), n, 64, None, 8, None, eng._p(ids),
                                          eng._p(uk), eng._p(uc), n, C.byref(d), eng.stream))
        valid = ~invalid
        uniq, inv = np.unique(keys_u[valid], return_inverse=True)
        assert d.value == len(uniq)
        got = ids.cpu().numpy().view(np.uint32)
        assert np.all(got[invalid] == 0xFFFFFFFF)
        assert np.array_equal(got[valid].astype(np.int64), inv.reshape(-1))
        assert np.array_equal(uk.cpu().numpy()[:d.value].view(np.uint64), uniq)
        assert np.array_equal(uc.cpu().numpy()[:d.value].view(np.uint64), (uniq * np.uint64(3)) ^ np.uint64(0x1234))


def test_dict_build_hash_path_matches_full_sort(eng, monkeypatch):
    """Above 8192 keys the dictionary is built by hash pre-dedup + a sort of the distinct keys only; the ids, the
    dictionary and the collision check must be those of the full radix sort (NBW_DICT_RADIX=1)."""
    cells = eng.generate_cells(0, 5, 0, 30000)
    a = eng.wl_fit(cells, 8, 3, build_features=False)
    monkeypatch.setenv("NBW_DICT_RADIX", "1")
    b = eng.wl_fit(cells, 8, 3, build_features=False)
    assert a.dims == b.dims
    assert torch.equal(a.ids, b.ids)
    for l in range(4):
        assert torch.equal(a.uniq_keys[l], b.uniq_keys[l]) and torch.equal(a.uniq_chks[l], b.uniq_chks[l])


def test_hash_collision_is_detected_large(eng):
    import ctypes as C
    n = 50000
    rng = np.random.RandomState(3)
    keys = rng.randint(0, 1000, size=n).astype(np.int64)
    chks = keys * 7
    chks[31337] += 1
    dk, dc = torch.from_numpy(keys).cuda(), torch.from_numpy(chks).cuda()
    ids = torch.empty(n, dtype=torch.int32, device="cuda")
    uk = torch.empty(n, dtype=torch.int64, device="cuda")
    uc = torch.empty(n, dtype=torch.int64, device="cuda")
    d = C.c_int64(0)
    rc = eng.lib.nbw_dict_build(eng.ctx, eng._p(dk), eng._p(dc), n, 64, None, 8, None, eng._p(ids),
                                eng._p(uk), eng._p(uc), n, C.byref(d), eng.stream)
    assert rc == 4



@pytest.mark.parametrize("fam,n", [(0, 7), (1, 150), (2, 333)])
def test_level_grams_small_kernel_matches_tensor_core_path(eng, fam, n, monkeypatch):
    """nbw_level_grams_i32 (one SIMT launch for all WL levels of a small training set) against one tcgen05 Gram per
    level and against the oracle: exact integers."""
    from nasbowl_b200 import synth
    train = synth.generate_cells_host(fam, 4, 0, n)
    fit = eng.wl_fit(eng.upload_cells(train), train.n_max, 3, False)
    a = eng.level_grams_i32(fit, 3)[:, :, :n].clone()
    monkeypatch.setenv("NBW_LEVEL_GRAMS_TC", "1")
    b = eng.level_grams_i32(fit, 3)[:, :, :n].clone()
    assert torch.equal(a, b)
    tot = a.sum(0).cpu().numpy()
    assert np.array_equal(tot, wl_oracle.gram_raw(train, 3, False))


def test_hash_collision_is_detected(eng):
    """Equal keys with different check words must be reported, not merged."""
    import ctypes as C
    from nasbowl_b200 import _lib
    keys = torch.tensor([5, 9, 5, 7], dtype=torch.int64, device="cuda")
    chks = torch.tensor([1, 2, 3, 4], dtype=torch.int64, device="cuda")
    ids = torch.empty(4, dtype=torch.int32, device="cuda")
    uk = torch.empty(4, dtype=torch.int64, device="cuda")
    uc = torch.empty(4, dtype=torch.int64, device="cuda")
    d = C.c_int64(0)
    rc = eng.lib.nbw_dict_build(eng.ctx, eng._p(keys), eng._p(chks), 4, 64, None, 8, None, eng._p(ids),
                                eng._p(uk), eng._p(uc), 4, C.byref(d), eng.stream)
    assert rc == 4
    with pytest.raises(_lib.HashCollision):
        eng._check(rc)


# ---------------------------------------------------------------------------- linear algebra
@pytest.mark.parametrize("n", [1, 5, 32, 33, 100, 257, 600])
def test_cholesky_inverse_gemm(eng, n):
    rng = np.random.RandomState(n)
    A = rng.randn(n, n)
    K = A @ A.T / n + 0.5 * np.eye(n)
    y = rng.randn(n)
    Kd = torch.from_numpy(K).cuda()
    yd = torch.from_numpy(y).cuda()
    L, Linv, alpha, logdet, yKy, trKinv, a2 = eng.gp_factor(Kd, 0.01, yd)
    Kl = K + 0.01 * np.eye(n)
    Lr = np.linalg.cholesky(Kl)
    np.testing.assert_allclose(L.cpu().numpy(), Lr, rtol=1e-10, atol=1e-12)
    np.testing.assert_allclose(Linv.cpu().numpy(), np.linalg.inv(Lr), rtol=1e-8, atol=1e-10)
    Ki = np.linalg.inv(Kl)
    np.testing.assert_allclose(alpha.cpu().numpy(), Ki @ y, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(logdet, np.linalg.slogdet(Kl)[1], rtol=1e-10)
    np.testing.assert_allclose(yKy, y @ Ki @ y, rtol=1e-9)
    np.testing.assert_allclose(trKinv, np.trace(Ki), rtol=1e-9)
    np.testing.assert_allclose(a2, (Ki @ y) @ (Ki @ y), rtol=1e-8)
    # GEMM in all four transpose modes
    m, k = 70, 45
    X, Y = rng.randn(m, k), rng.randn(k, n)
    for ta in (False, True):
        for tb in (False, True):
            Xa = torch.from_numpy(np.ascontiguousarray(X.T if ta else X)).cuda()
            Yb = torch.from_numpy(np.ascontiguousarray(Y.T if tb else Y)).cuda()
            Cd = torch.from_numpy(np.ones((m, n))).cuda()
            eng.gemm(ta, tb, m, n, k, 2.0, Xa, Xa.shape[1], Yb, Yb.shape[1], 0.5, Cd, n)
            np.testing.assert_allclose(Cd.cpu().numpy(), 2.0 * X @ Y + 0.5, rtol=1e-12, atol=1e-12)


def test_not_positive_definite_is_reported(eng):
    from nasbowl_b200 import _lib
    K = torch.from_numpy(np.array([[1.0, 2.0], [2.0, 1.0]])).cuda()
    y = torch.zeros(2, dtype=torch.float64, device="cuda")
    with pytest.raises(_lib.NotPositiveDefinite):
        eng.gp_factor(K, 0.0, y)
    with pytest.raises(RuntimeError, match="not positive definite despite of jitter"):
        eng.pd_inverse(K, 1e-3, y)        # utils.py:135-137


# ---------------------------------------------------------------------------- GP + acquisitions
def _fit_both(train, y, cands, oa, opt):
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    gp = GraphGP(train, torch.tensor(y), [WeisfilerLehman(h=2, oa=oa)])
    gp.fit(wl_subtree_candidates=cands, optimize_lik=opt, max_lik=0.01)
    st = gp_oracle.gp_fit(train, y.astype(np.float64), h_candidates=cands, oa=oa, optimize_lik=opt, max_lik=0.01)
    return gp, st


@pytest.mark.parametrize("opt", [False, True])
@pytest.mark.parametrize("oa", [False, True])
def test_gp_fit_predict_acquisitions(family, oa, opt):
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphUpperConfidentBound
    _, z, train, pool = family
    tag = "oa%d_opt%d" % (int(oa), int(opt))
    cands = tuple(int(c) for c in z["gp_cands_" + tag])
    gp, st = _fit_both(train, z["y"], cands, oa, opt)
    k = gp.kernels[0]
    # --- fit: exact discrete quantities, fp64 tolerance elsewhere
    assert k.h == st.h == int(z["gp_h_" + tag])
    assert np.float32(gp.likelihood) == np.float32(st.lik) == np.float32(z["gp_lik_" + tag])
    grid = [gp.nlml_grid[id(k)][c] for c in cands]
    np.testing.assert_allclose(grid, st.nlml_grid, rtol=1e-7)
    np.testing.assert_allclose(grid, z["gp_grid_" + tag], rtol=2e-5)          # fp32 reference
    np.testing.assert_allclose(gp.K.numpy(), st.K, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(gp.K_i.numpy(), st.K_i, rtol=RTOL64, atol=1e-8)
    np.testing.assert_allclose(float(gp.logDetK), st.logDetK, rtol=1e-10)
    if opt:
        np.testing.assert_allclose(float(gp.nlml), st.nlml, rtol=1e-9)
    # --- posterior: fused pool path and reference-shaped dense path vs the fp64 oracle
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool)
    mu_d, var_d = gp.predict_diag(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    # the prefix form (H over node pairs, linear kernel) and the general form (G over
    # (node, level) pairs) are two evaluation orders of the same quadratic form
    pm = gp._score_state()
    assert (pm.H is not None) != (oa and pm.oa_exc is None)     # oa: item form (chains + exceptions) or general G
    pm.use_prefix = False
    mu_g, var_g = gp.predict_diag(pool)
    pm.use_prefix = True
    np.testing.assert_allclose(mu_g.numpy(), mu_d.numpy(), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(var_g.numpy(), var_d.numpy(), rtol=1e-8, atol=1e-13)
    np.testing.assert_allclose(var_g.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    mu_f, cov_f = gp.predict(pool)
    mu_fo, cov_fo = gp_oracle.gp_predict_full(st, train, pool)
    np.testing.assert_allclose(mu_f.numpy(), mu_fo, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(cov_f.numpy(), cov_fo, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(mu_f.numpy(), z["gp_mu_" + tag], rtol=2e-3, atol=2e-3)     # fp32 reference
    # --- acquisitions
    inc = gp_oracle.incumbent(st)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, inc)
    acq = GraphExpectedImprovement(gp)
    assert float(acq._get_incumbent()) == pytest.approx(inc, rel=1e-7) == pytest.approx(float(z["incumbent_" + tag]), rel=1e-6)
    np.testing.assert_allclose(acq.eval(pool).numpy(), ei_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(acq.eval(pool).numpy(), z["ei_" + tag], rtol=5e-2, atol=2e-3)
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    np.testing.assert_allclose(eis.reshape(-1), ei_o, rtol=RTOL32, atol=1e-9)
    assert eis.dtype == np.float32 and eis.shape == (len(pool), 1)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(eis.reshape(-1), 5))
    if opt:
        aei = GraphExpectedImprovement(gp, augmented_ei=True, in_fill='posterior')
        np.testing.assert_allclose(aei.eval(pool).numpy(),
                                   gp_oracle.expected_improvement(mu_o, var_o, inc, augmented=True, lik=st.lik),
                                   rtol=RTOL64, atol=1e-12)
        ucb = GraphUpperConfidentBound(gp)
        ucb.iters = 1
        np.testing.assert_allclose(ucb.eval(pool).numpy(),
                                   gp_oracle.upper_confidence_bound(mu_o, var_o, gp_oracle.ucb_beta(1)),
                                   rtol=RTOL64)
        assert float(ucb.beta) == gp_oracle.ucb_beta(1)      # beta is an fp32 torch scalar in the reference
        np.testing.assert_allclose(ucb.eval(pool).numpy(), z["ucb_" + tag], rtol=5e-3, atol=5e-3)
        ref_top = set(int(i) for i in z["top5_" + tag])
        got = set(int(i) for i in idx)
        if got != ref_top:       # only candidates the fp32 reference cannot tell apart may swap
            vals = [z["ei_" + tag][i] for i in got ^ ref_top]
            assert max(vals) - min(vals) < 1e-4 * max(1.0, abs(max(vals)))


def test_networkx_entry_points_match_packed(family):
    """list[nx.DiGraph] in, the reference's call shapes out (gp.py:240-283, acquisitions.py:59-113)."""
    from nasbowl_b200 import OpVocab
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import SumKernel, WeisfilerLehman
    name, z, train, pool = family
    vocab = OpVocab(list(z["op_names"]))
    g_train, g_pool = train.to_networkx(vocab), pool.to_networkx(vocab)
    k = WeisfilerLehman(h=2)
    K = k.fit_transform(g_train, rebuild_model=True)
    assert K.dtype == torch.float64 and np.array_equal(K.numpy(), z["Kraw_h2_oa0"])
    Kn = SumKernel(WeisfilerLehman(h=2)).fit_transform([1.0], g_train)
    np.testing.assert_allclose(Kn.numpy(), gp_oracle.normalize_gram(z["Kraw_h2_oa0"]), rtol=1e-13)
    y = torch.tensor(z["y"])
    gp = GraphGP(g_train, y, [WeisfilerLehman(h=2)])
    with pytest.raises(ValueError):
        gp.predict(g_pool)                                   # predict before fit (gp.py:245-247)
    gp.fit(wl_subtree_candidates=(1, 2, 3), optimize_lik=False)
    mu, cov = gp.predict(g_pool)
    assert mu.shape == (len(pool),) and cov.shape == (len(pool), len(pool))
    mu1, cov1 = gp.predict(g_pool[4])                        # single graph (gp.py:242-244)
    np.testing.assert_allclose(mu1.numpy(), mu.numpy()[4:5], rtol=1e-9)
    np.testing.assert_allclose(cov1.numpy()[0, 0], cov.numpy()[4, 4], rtol=1e-9)
    acq = GraphExpectedImprovement(gp)
    e = acq.eval(g_pool[4])
    xs, eis, idx = acq.propose_location(g_pool, top_n=3)
    assert len(xs) == 3 and all(x is g_pool[int(i)] for x, i in zip(xs, idx))
    np.testing.assert_allclose(float(e), eis[4, 0], rtol=1e-5)
    gp.reset_XY(g_train[:20], y[:20])
    assert gp.K_i is None
    with pytest.raises(ValueError):
        gp.predict(g_pool)
    gp.fit(wl_subtree_candidates=(0, 1), optimize_lik=True)
    assert gp.K.shape == (20, 20)
    # growing the list again re-uses the records of the 20 cells already packed (GraphGP._pack_train): same fit as
    # a GP that sees the list for the first time
    gp.reset_XY(g_train[:30], y[:30])
    gp.likelihood = 1e-3                                      # (the grid search starts from the current noise, Q5)
    gp.fit(wl_subtree_candidates=(0, 1, 2), optimize_lik=True)
    fresh = GraphGP(g_train[:30], y[:30], [WeisfilerLehman(h=2)])
    fresh.fit(wl_subtree_candidates=(0, 1, 2), optimize_lik=True)
    assert gp.kernels[0].h == fresh.kernels[0].h and gp.likelihood == fresh.likelihood
    np.testing.assert_allclose(gp.K.numpy(), fresh.K.numpy(), rtol=1e-12)
    mu2, _ = gp.predict(g_pool)
    mu3, _ = fresh.predict(g_pool)
    np.testing.assert_allclose(mu2.numpy(), mu3.numpy(), rtol=1e-9, atol=1e-12)


def test_appendix_b_known_answer():
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP, GraphUpperConfidentBound
    from nasbowl_b200.kernels import WeisfilerLehman
    z, batch = load_golden("appendix_b")
    train, test = batch[:4], batch[4:]
    gp = GraphGP(train, torch.tensor(z["y"]), [WeisfilerLehman(h=2)])
    gp.fit(optimize_lik=False, wl_subtree_candidates=(0, 1, 2))
    assert gp.kernels[0].h == 2 and np.float32(gp.likelihood) == np.float32(1e-3)
    np.testing.assert_allclose(gp.K.numpy(), z["gp_K"], rtol=2e-6)
    np.testing.assert_allclose(gp.K_i.numpy(), z["gp_Ki"], rtol=2e-4)
    np.testing.assert_allclose(float(gp.logDetK), z["gp_logDetK"], rtol=1e-5)
    mu, cov = gp.predict(test)
    np.testing.assert_allclose(mu.numpy(), z["gp_mu"], rtol=1e-4)
    np.testing.assert_allclose(cov.numpy(), z["gp_cov"], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(GraphExpectedImprovement(gp).eval(test).numpy(), z["ei"], rtol=5e-3, atol=1e-7)
    ucb = GraphUpperConfidentBound(gp)
    ucb.iters = 1
    np.testing.assert_allclose(ucb.eval(test).numpy(), z["ucb"], rtol=1e-4)


# ---------------------------------------------------------------------------- top-k
def test_topk_distinct_and_plain(eng):
    rng = np.random.RandomState(3)
    for n in [1, 5, 4096, 4097, 100000, 1 << 20]:
        s = np.round(rng.randn(n), 2).astype(np.float32)      # many duplicates
        s[rng.rand(n) < 0.01] = np.nan
        d = torch.from_numpy(s).cuda()
        for k in [1, 5, 64]:
            vals, idx = eng.topk(d, k, distinct=True, index_base=7)
            ok = ~np.isnan(s)
            uniq, first = np.unique(s[ok], return_index=True)
            first = np.nonzero(ok)[0][first]
            order = np.argsort(-uniq, kind="stable")[:k]
            assert np.array_equal(vals, uniq[order])
            assert np.array_equal(idx, first[order] + 7)
            vals, idx = eng.topk(d, k, distinct=False)
            cand = np.nonzero(ok)[0]
            order = cand[np.lexsort((cand, -s[cand]))][:k]
            assert np.array_equal(idx, order) and np.array_equal(vals, s[order])


# ---------------------------------------------------------------------------- synthetic pools
@pytest.mark.parametrize("fam", [0, 1, 2])
def test_device_generator_matches_host(eng, fam):
    host = synth.generate_cells_host(fam, 99, 12345, 3000)
    dev = eng.generate_cells(fam, 99, 12345, 3000).cpu().numpy()
    assert np.array_equal(dev, host.to_records())
    # shards concatenate to the whole
    a = eng.generate_cells(fam, 99, 12345, 1000).cpu().numpy()
    b = eng.generate_cells(fam, 99, 13345, 2000).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), dev)


# ---------------------------------------------------------------------------- edge cases
def test_edge_cases(eng):
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(0, 5, 0, 24)
    y = np.random.RandomState(0).randn(24)
    gp = GraphGP(train, torch.tensor(y), [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(2,), optimize_lik=False)
    st = gp_oracle.gp_fit(train, y, h_candidates=(2,), optimize_lik=False)
    lab = np.full((5, 8), NO_NODE, np.uint8)
    suc = np.zeros((5, 8), np.uint16)
    lab[0, :2] = [0, 1]; suc[0, 0] = 2                        # input -> output (smallest legal cell)
    lab[1, :3] = [0, 200, 1]; suc[1, 0] = 0b110; suc[1, 1] = 0b100   # op code never seen in training
    lab[2, :3] = [0, 3, 1]; suc[2, 0] = 0b100                 # node 1 is isolated: counted at level 0 only (Q4)
    lab[3, :1] = [2]                                          # a single node, no edge
    lab[4] = train.labels[0]; suc[4] = train.succ[0]          # a training cell: variance at the clamp
    pool = CellBatch(8, lab, suc)
    mu, var = gp.predict_diag(pool)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool)
    np.testing.assert_allclose(mu.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    # a duplicated training cell: predictive variance lam + lam*(1 - ...) lies in (lam, 2 lam]
    assert st.lik * st.y_std ** 2 <= var.numpy()[4] <= 2.0 * st.lik * st.y_std ** 2
    # cells without any label known to the dictionary fall back to the prior
    assert var.numpy()[3] == pytest.approx((1.0 + st.lik) * st.y_std ** 2, rel=1e-12) or lab[3, 0] in train.labels
    # empty pool
    scores, _, _, _ = eng.score_pool(gp.pool_model("EI"), torch.empty((0, 16), dtype=torch.uint8, device="cuda"), 0)
    assert scores.numel() == 0
    with pytest.raises(ValueError):
        eng.wl_fit(torch.empty((0, 16), dtype=torch.uint8, device="cuda"), 8, 2)
    with pytest.raises(TypeError):
        WeisfilerLehman(h=-1)


# ---------------------------------------------------------------------------- full-size properties
def test_million_candidate_pool_properties(eng):
    """BASELINE config 2 size (NB201-shaped, n_train=150, 1M candidates): size-independent
    properties instead of an oracle pass — shard invariance, invariance under renumbering of a
    cell's nodes, duplicates of training cells sit at the variance clamp, top-k is consistent."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    n, M = 150, 1_000_000
    train = CellBatch.from_records(eng.generate_cells(1, 0, 0, n).cpu().numpy(), 8)
    y = torch.randn(n, generator=torch.Generator().manual_seed(0))
    gp = GraphGP(train, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=tuple(range(6)), optimize_lik=True, max_lik=0.01)
    pool = eng.generate_cells(1, 1234, 10_000, M)
    acq = GraphExpectedImprovement(gp)
    scores, mu, var, sc = acq.score_pool(pool, want64=True)
    s = scores.cpu().numpy()
    assert np.isfinite(s).all() and (s >= 0).all()
    # (1) shard invariance: bitwise equal scores whichever way the pool is split
    cut = 333_337
    a, _, _, _ = acq.score_pool(pool[:cut].contiguous())
    b, _, _, _ = acq.score_pool(pool[cut:].contiguous())
    assert torch.equal(torch.cat([a, b]), scores)
    # (2) node renumbering: reverse the node order of every cell
    sub = CellBatch.from_records(pool[:20000].cpu().numpy(), 8)
    k = sub.n_nodes
    lab = np.full_like(sub.labels, NO_NODE)
    suc = np.zeros_like(sub.succ)
    for v in range(8):
        src = k - 1 - v
        okv = src >= 0
        lab[okv, v] = sub.labels[okv, src[okv]]
        m = sub.succ[okv, src[okv]].astype(np.uint32)
        out = np.zeros_like(m)
        for j in range(8):
            tgt = k[okv] - 1 - j
            bit = (m >> j) & 1
            out |= np.where(tgt >= 0, bit << np.maximum(tgt, 0), 0).astype(np.uint32)
        suc[okv, v] = out
    perm, _, _, _ = acq.score_pool(CellBatch(8, lab, suc))
    np.testing.assert_allclose(perm.cpu().numpy(), s[:20000], rtol=1e-5, atol=1e-9)
    # (3) against the oracle on a slice the CPU finishes in seconds
    st = gp_oracle.gp_fit(train, y.numpy().astype(np.float64), h_candidates=tuple(range(6)), optimize_lik=True)
    assert st.h == gp.kernels[0].h
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, sub[:4000])
    np.testing.assert_allclose(mu.cpu().numpy()[:4000], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy()[:4000], var_o, rtol=RTOL64, atol=1e-12)
    # (4) top-5 distinct agrees with numpy on the full score vector
    _, _, idx = acq.propose_location(pool, top_n=5)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(s, 5))


def test_sharded_topk_device_merge_equals_global(eng):
    """The multi-GPU selection path on one GPU: per-shard device top-k lists, concatenated the
    way the all-gather delivers them, merged by nbw_topk_merge == top-k of the whole pool."""
    from nasbowl_b200 import parallel
    rng = np.random.RandomState(11)
    s = np.round(rng.randn(300_000), 2).astype(np.float32)
    s[::1013] = np.nan
    d = torch.from_numpy(s).cuda()
    for distinct in (True, False):
        ref_v, ref_i = eng.topk(d, 5, distinct=distinct)
        for world in (1, 2, 3, 8):
            parts = []
            for r in range(world):
                lo, hi = parallel.shard_range(len(s), r, world)
                parts.append(eng.topk_device(d[lo:hi], 5, distinct=distinct, index_base=lo).clone())
            parts.append(eng.topk_device(d[:0], 5, distinct=distinct))          # an empty shard
            v, i = eng.topk_merge(torch.cat(parts), 5, distinct=distinct)
            assert np.array_equal(i, ref_i) and np.array_equal(v, ref_v)
            hv, hi_ = parallel.merge_topk(s, np.arange(len(s)), 5, distinct)     # host restatement
            assert np.array_equal(i, hi_) and np.array_equal(v, hv)


def test_host_pool_pipeline_matches_resident_path(eng):
    """propose_location on packed records in HOST memory (chunked H2D / kernel / D2H pipeline)
    returns the same scores and selection as the resident-pool path."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(0, 3, 0, 60)
    gp = GraphGP(train, torch.randn(60, generator=torch.Generator().manual_seed(1)), [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True)
    acq = GraphExpectedImprovement(gp)
    for M in (5, 1000, 100_003):
        dev = eng.generate_cells(0, 9, 500, M)
        host = dev.cpu()
        xs_d, eis_d, idx_d = acq.propose_location(dev, top_n=5)
        xs_h, eis_h, idx_h = acq.propose_location(host, top_n=5)
        xs_p, eis_p, idx_p = acq.propose_location(host.pin_memory(), top_n=5)
        assert np.array_equal(eis_d, eis_h) and np.array_equal(eis_d, eis_p)
        assert np.array_equal(idx_d, idx_h) and np.array_equal(idx_d, idx_p)
        assert eis_h.shape == (M, 1) and eis_h.dtype == np.float32


def test_mapped_and_staged_host_entry_points_agree(eng):
    """nbw_score_pool_mapped (kernel reads pinned records over PCIe, mirrors the scores into pinned
    host memory) == nbw_score_pool_host (staged) == nbw_score_pool (resident), bit for bit; pageable
    records are refused by the mapped entry point."""
    from nasbowl_b200 import _lib
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    for family, h in ((1, 3), (2, 2)):
        train = synth.generate_cells_host(family, 5, 0, 40)
        gp = GraphGP(train, torch.randn(40, generator=torch.Generator().manual_seed(2)), [WeisfilerLehman(h=h)],
                     n_max=train.n_max)
        gp.fit(wl_subtree_candidates=(h,), optimize_lik=False)
        acq = GraphExpectedImprovement(gp)
        model = acq._model()
        for M in (1, 33, 50_001):
            dev = eng.generate_cells(family, 17, 100, M)
            resident = acq.score_pool(dev)[0].cpu()
            pinned = dev.cpu().pin_memory()
            for pieces in (1, 3):
                s_dev, s_host = eng.score_host_pool(model, pinned, chunks=pieces, mapped=True)
                torch.cuda.synchronize()
                assert torch.equal(s_dev.cpu(), resident) and torch.equal(s_host, resident)
            s_dev, s_host = eng.score_host_pool(model, pinned, chunks=4)
            torch.cuda.synchronize()
            assert torch.equal(s_dev.cpu(), resident) and torch.equal(s_host, resident)
    with pytest.raises(_lib.NbwError, match="pinned"):
        eng.score_host_pool(model, dev.cpu(), chunks=1, mapped=True)


def test_device_mutation_pool_matches_host(eng):
    par = synth.generate_cells_host(0, 3, 0, 12)
    dev_par = eng.upload_cells(par)
    host = synth.mutate_cells_host(0, par, 7, 500, 3000)
    dev = eng.mutate_cells(0, dev_par, 7, 500, 3000).cpu().numpy()
    assert np.array_equal(dev, host.to_records())
    a = eng.mutate_cells(0, dev_par, 7, 500, 1000).cpu().numpy()
    b = eng.mutate_cells(0, dev_par, 7, 1500, 2000).cpu().numpy()
    assert np.array_equal(np.concatenate([a, b]), dev)


@pytest.mark.parametrize("family,edits", [(1, 1), (2, 1), (2, 4)])
def test_device_encoded_generation_and_mutation_match_host(eng, family, edits):
    """nbw_generate_encoded / nbw_mutate_encoded (NB201 and DARTS architectures edited on their
    encodings) are bit-identical to the host twins, shard by shard, and feed the scoring path."""
    cells, enc = eng.generate_encoded(family, 21, 1000, 300)
    hb, henc = synth.generate_encoded_host(family, 21, 1000, 300)
    assert np.array_equal(cells.cpu().numpy(), hb.to_records()) and np.array_equal(enc.cpu().numpy(), henc)
    assert np.array_equal(cells.cpu().numpy(), eng.generate_cells(family, 21, 1000, 300).cpu().numpy())
    par = enc[:12]
    hc, he = synth.mutate_encoded_host(family, henc[:12], 7, 500, 2000, edits)
    dc, de = eng.mutate_encoded(family, par, 7, 500, 2000, edits)
    assert np.array_equal(dc.cpu().numpy(), hc.to_records()) and np.array_equal(de.cpu().numpy(), he)
    a, ea = eng.mutate_encoded(family, par, 7, 500, 700, edits)
    b, eb = eng.mutate_encoded(family, par, 7, 1200, 1300, edits)
    assert torch.equal(torch.cat([a, b]), dc) and torch.equal(torch.cat([ea, eb]), de)
    # second generation: children become parents
    g2, e2 = eng.mutate_encoded(family, de[:50], 9, 0, 500, edits)
    h2, he2 = synth.mutate_encoded_host(family, he[:50], 9, 0, 500, edits)
    assert np.array_equal(g2.cpu().numpy(), h2.to_records()) and np.array_equal(e2.cpu().numpy(), he2)
    # the mutated pool goes straight into the scoring kernel
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(family, 3, 0, 30)
    gp = GraphGP(train, torch.randn(30, generator=torch.Generator().manual_seed(1)), [WeisfilerLehman(h=1)],
                 n_max=train.n_max)
    gp.fit(wl_subtree_candidates=(1,), optimize_lik=False)
    acq = GraphExpectedImprovement(gp)
    s_dev = acq.score_pool(dc)[0].cpu().numpy()
    s_host = acq.score_pool(hc)[0].cpu().numpy()
    assert np.array_equal(s_dev, s_host) and np.isfinite(s_dev).all()


@pytest.mark.parametrize("family", [0, 2])
def test_small_sort_path_equals_radix_path(eng, family):
    """The dictionary build sorts up to 8192 keys in one CTA and larger inputs by radix passes: both
    give the same ids, dictionary and features (and each size class is exercised here)."""
    import os
    for n_cells in (1, 7, 500, 1024, 2500):
        cells = eng.generate_cells(family, 5, 77, n_cells)
        n_max = 16 if family == 2 else 8
        a = eng.wl_fit(cells, n_max, 3, False)
        os.environ["NBW_DICT_RADIX"] = "1"
        try:
            b = eng.wl_fit(cells, n_max, 3, False)
        finally:
            del os.environ["NBW_DICT_RADIX"]
        assert a.dims == b.dims and torch.equal(a.ids, b.ids) and torch.equal(a.phi, b.phi)
        for l in range(4):
            assert torch.equal(a.uniq_keys[l], b.uniq_keys[l]) and torch.equal(a.uniq_chks[l], b.uniq_chks[l])


def _random_dags(rng, n_cells, n_max, n_ops, p_edge, allow_isolated=True):
    """Arbitrary labelled DAGs (not sampler-shaped): random sizes, dense / sparse adjacency,
    isolated nodes, duplicate cells — the edge cases of the WL relabelling rules."""
    lab = np.full((n_cells, n_max), NO_NODE, np.uint8)
    suc = np.zeros((n_cells, n_max), np.uint16)
    for i in range(n_cells):
        k = rng.randint(1, n_max + 1)
        lab[i, :k] = rng.randint(0, n_ops, size=k)
        perm = rng.permutation(k)                  # edges follow a hidden topological order
        for a in range(k):
            for b in range(a + 1, k):
                if rng.rand() < p_edge:
                    suc[i, perm[a]] |= np.uint16(1 << perm[b])
        if not allow_isolated:
            continue
    return CellBatch(n_max, lab, suc)


@pytest.mark.parametrize("n_max,p_edge", [(8, 0.15), (8, 0.6), (8, 1.0), (16, 0.1), (16, 0.45)])
def test_random_dags_exact_wl_and_posterior(eng, n_max, p_edge):
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    rng = np.random.RandomState(int(n_max * 100 + p_edge * 10))
    train = _random_dags(rng, 120, n_max, 6, p_edge)
    pool = CellBatch.concat([_random_dags(rng, 400, n_max, 7, p_edge), train[:20]])   # op 6 is unseen in training
    for oa in (False, True):
        for h in (1, 4, 7):
            fit = eng.wl_fit(eng.upload_cells(train), n_max, h, oa)
            _, dims = wl_oracle.wl_labels_fast(train, h)
            assert fit.dims == dims
            K = eng.gram(fit.phi, fit.phi, 0, fit.k_end(h)).cpu().numpy()
            assert np.array_equal(K, wl_oracle.gram_raw(train, h, oa))
        y = rng.randn(len(train))
        gp = GraphGP(train, torch.tensor(y), [WeisfilerLehman(h=2, oa=oa)])
        gp.fit(wl_subtree_candidates=(0, 1, 2, 3, 4), optimize_lik=True)
        st = gp_oracle.gp_fit(train, y, h_candidates=(0, 1, 2, 3, 4), oa=oa, optimize_lik=True)
        assert gp.kernels[0].h == st.h
        mu, var = gp.predict_diag(pool)
        mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool)
        np.testing.assert_allclose(mu.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
        np.testing.assert_allclose(var.numpy(), var_o, rtol=RTOL64, atol=1e-12)


def test_cfg1_regression_case():
    """BASELINE.json configs[0]: GraphGP(...).fit(optimize_lik=True); predict(400 test cells)."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    z, batch = load_golden("cfg1_nb101")
    n = int(z["n_train"])
    train, test = batch[:n], batch[n:]
    gp = GraphGP(train, torch.tensor(z["y"]), [WeisfilerLehman(h=2)])
    gp.fit(optimize_lik=True)                                      # defaults: h in range(5), 20 Adam iterations
    st = gp_oracle.gp_fit(train, z["y"].astype(np.float64), h_candidates=tuple(range(5)), optimize_lik=True)
    assert gp.kernels[0].h == st.h == int(z["gp_h"])
    assert np.float32(gp.likelihood) == np.float32(z["gp_lik"])
    np.testing.assert_allclose(float(gp.nlml), st.nlml, rtol=1e-9)
    mu, cov = gp.predict(test)                                      # reference-shaped: full 400 x 400 covariance
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, test)
    np.testing.assert_allclose(mu.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(np.diag(cov.numpy()), var_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(mu.numpy(), z["gp_mu"], rtol=2e-3, atol=2e-3)          # fp32 reference
    mu_t, _ = gp.predict(train)
    np.testing.assert_allclose(mu_t.numpy(), z["gp_mu_train"], rtol=2e-3, atol=2e-3)
    acq = GraphExpectedImprovement(gp)
    _, eis, idx = acq.propose_location(test, top_n=5)
    np.testing.assert_allclose(eis.reshape(-1), gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent(st)),
                               rtol=RTOL32, atol=1e-9)
    np.testing.assert_allclose(eis.reshape(-1), z["ei"], rtol=5e-2, atol=2e-3)
    got, ref_top = set(int(i) for i in idx), set(int(i) for i in z["top5"])
    if got != ref_top:
        vals = [z["ei"][i] for i in got ^ ref_top]
        assert max(vals) - min(vals) < 1e-4 * max(1.0, abs(max(vals)))


@pytest.mark.parametrize("name", ["cfg2_nb201_n150", "cfg3_nb101_n500", "cfg4_darts_n1000"])
def test_baseline_training_set_sizes_vs_reference_golden(eng, name):
    """BASELINE configs[1] / configs[2] at their training-set sizes against vectors of the UNMODIFIED reference
    (tests/golden/cfg2_nb201_n150.npz, cfg3_nb101_n500.npz; make_golden.py::baseline_size_case): chosen depth, fitted
    noise, integer Gram bit-exact, marginal-likelihood grid; posterior / EI at the reference's fp32 tolerance and at
    1e-6 against the fp64 oracle; the reference's own propose_location top-5."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    z, batch = load_golden(name)
    n = int(z["n_train"])
    train, pool = batch[:n], batch[n:]
    cands = tuple(int(c) for c in z["cands"])
    gp = GraphGP(train, torch.tensor(z["y"]), [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=cands, optimize_lik=True, max_lik=0.01)
    st = gp_oracle.gp_fit(train, z["y"].astype(np.float64), h_candidates=cands, optimize_lik=True, max_lik=0.01)
    assert gp.kernels[0].h == st.h == int(z["gp_h"])
    assert np.float32(gp.likelihood) == np.float32(z["gp_lik"])
    np.testing.assert_allclose(float(gp.nlml), st.nlml, rtol=1e-9)
    np.testing.assert_allclose(float(gp.nlml), float(z["gp_nlml"]), rtol=1e-5)
    fit = eng.wl_fit(eng.upload_cells(train), train.n_max, st.h, False)
    assert np.array_equal(eng.gram(fit.phi, fit.phi, 0, fit.k_end(st.h)).cpu().numpy(), golden_kraw(z))
    acq = GraphExpectedImprovement(gp)
    s, mu, var, s64 = acq.score_pool(pool, want64=True)
    mu, var, s64 = mu.cpu().numpy(), var.cpu().numpy(), s64.cpu().numpy()
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent(st))
    np.testing.assert_allclose(mu, mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var, var_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(s64, ei_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(mu, z["gp_mu"], rtol=2e-3, atol=1e-2)               # fp32 reference
    np.testing.assert_allclose(var, z["gp_var"], rtol=5e-2, atol=5e-3)
    ne = int(z["n_eval"])
    np.testing.assert_allclose(s64[:ne], z["ei"], rtol=5e-2, atol=2e-3)
    _, eis, idx = acq.propose_location(pool[:ne], top_n=5)
    assert set(int(i) for i in idx) == set(int(i) for i in z["top5"])


@pytest.mark.parametrize("family,h,oa", [(1, 5, False), (1, 5, True), (0, 5, False), (2, 4, False)])
def test_stable_suffix_fast_path_is_exact(eng, family, h, oa):
    """When the training dictionary stops refining, deeper WL levels inherit membership and take the
    child label instead of hashing and probing (nbw_model.child).  Same scores, features and
    self-similarities as the hashed path, bit for bit, and agreement with the oracle."""
    import os
    from nasbowl_b200.bayesopt import GraphExpectedImprovement, GraphGP
    from nasbowl_b200.kernels import WeisfilerLehman
    from nasbowl_b200.pool_model import PoolModel
    train = synth.generate_cells_host(family, 3, 0, 150)
    y = torch.randn(150, generator=torch.Generator().manual_seed(7))
    M = 60_000
    pool = eng.generate_cells(family, 9, 500, M)

    def run(no_stable):
        if no_stable:
            os.environ["NBW_NO_STABLE"] = "1"
        try:
            gp = GraphGP(train, y, [WeisfilerLehman(h=h, oa=oa)], n_max=train.n_max)
            gp.fit(wl_subtree_candidates=(h,), optimize_lik=False)
            pm = gp._score_state()
            acq = GraphExpectedImprovement(gp)
            s, mu, var, s64 = acq.score_pool(pool, want64=True)
            out = [s.clone(), mu.clone(), var.clone()]
            if not oa:
                pm.use_prefix = False                       # the general G path through the same levels
                out.append(acq.score_pool(pool)[0].clone())
                pm.use_prefix = True
            fitm = PoolModel.from_fit(eng, gp._dev['fits'][0], h)
            phi, ss = eng.pool_features(fitm.descriptor(), pool, M)
            out += [phi.clone(), ss.clone()]
            return pm.first_stable, gp, out
        finally:
            os.environ.pop("NBW_NO_STABLE", None)
    first, gp, fast = run(False)
    first_off, _, slow = run(True)
    assert first_off == 0
    dims = gp._dev['fits'][0].dims
    if first == 0:
        pytest.skip("the dictionary of this training set keeps refining: dims %s" % dims)
    assert first >= 3 and dims[first - 1] == dims[first - 2]
    for a, b in zip(fast, slow):
        assert torch.equal(a, b)
    sub = CellBatch.from_records(pool[:300].cpu().numpy(), train.n_max)
    st = gp_oracle.gp_fit(train, y.numpy().astype(np.float64), h_candidates=(h,), oa=oa, optimize_lik=False)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, sub)
    np.testing.assert_allclose(fast[1][:300].cpu().numpy(), mu_o, rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(fast[2][:300].cpu().numpy(), var_o, rtol=1e-6, atol=1e-12)
