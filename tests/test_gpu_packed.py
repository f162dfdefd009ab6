"""GPU parity tests of the packed scoring path (csrc/score_packed.cu): dense lane packing, fetched hashes,
exact in-cell class refinement, tag-bucket probes, sums of kernels over one or several records per
candidate, weighted WL levels, undirected kernels and the fused top-n — against the CPU oracle, against
the first-generation kernel (NBW_SCORE_LEGACY=1) and at the sizes of BASELINE configs[2] / configs[3]."""
import os

import numpy as np
import pytest
import torch

from nasbowl_b200 import synth
from nasbowl_b200.cells import NO_NODE, CellBatch, CellSlots
from oracle import gp_oracle

pytestmark = pytest.mark.gpu

RTOL64 = 1e-6
RTOL32 = 1e-4


@pytest.fixture(scope="module")
def eng():
    from nasbowl_b200.engine import Engine
    return Engine.get()


def _gp(train, y, kernels, **kw):
    from nasbowl_b200.bayesopt import GraphGP
    return GraphGP(train, torch.as_tensor(np.asarray(y, dtype=np.float64)), kernels, **kw)


def _legacy(fn):
    os.environ["NBW_SCORE_LEGACY"] = "1"
    try:
        return fn()
    finally:
        os.environ.pop("NBW_SCORE_LEGACY", None)


# ---------------------------------------------------------------------------- packed vs legacy vs oracle
@pytest.mark.parametrize("fam,h", [(0, 0), (0, 1), (0, 2), (0, 3), (1, 2), (1, 5), (2, 2), (2, 4)])
def test_packed_kernel_matches_first_generation_and_oracle(eng, fam, h):
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    n, M = 120, 30_011                       # a pool size that is a multiple of nothing
    train = synth.generate_cells_host(fam, 5, 0, n)
    y = np.random.RandomState(fam * 10 + h).randn(n)
    gp = _gp(train, y, [WeisfilerLehman(h=h)], n_max=train.n_max)
    gp.fit(wl_subtree_candidates=(h,), optimize_lik=False)
    pool = eng.generate_cells(fam, 77, 1000, M)
    acq = GraphExpectedImprovement(gp)
    assert eng.fused_topk_ok(acq._model(), 5)
    s, mu, var, s64 = [t.clone() for t in acq.score_pool(pool, want64=True)]
    s_l, mu_l, var_l, s64_l = _legacy(lambda: [t.clone() for t in acq.score_pool(pool, want64=True)])
    if train.n_max == 8:
        # 8-node records are scored one THREAD per cell; the lane-per-node kernel must give the same numbers
        os.environ["NBW_SCORE_LANES"] = "1"
        try:
            s_n, mu_n, var_n, s64_n = [t.clone() for t in acq.score_pool(pool, want64=True)]
            _, _, idx_n = acq.propose_location(pool, top_n=5)
        finally:
            os.environ.pop("NBW_SCORE_LANES", None)
        np.testing.assert_allclose(mu.cpu().numpy(), mu_n.cpu().numpy(), rtol=1e-11, atol=1e-13)
        np.testing.assert_allclose(var.cpu().numpy(), var_n.cpu().numpy(), rtol=1e-9, atol=1e-14)
        _, _, idx_c = acq.propose_location(pool, top_n=5)
        if torch.equal(s, s_n):
            assert np.array_equal(np.sort(idx_c), np.sort(idx_n))
    # two evaluation orders of the same sums of fp64 terms
    np.testing.assert_allclose(mu.cpu().numpy(), mu_l.cpu().numpy(), rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(var.cpu().numpy(), var_l.cpu().numpy(), rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(s64.cpu().numpy(), s64_l.cpu().numpy(), rtol=1e-8, atol=1e-14)
    sub = CellBatch.from_records(pool[:1500].cpu().numpy(), train.n_max)
    st = gp_oracle.gp_fit(train, y, h_candidates=(h,), optimize_lik=False)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, sub)
    np.testing.assert_allclose(mu.cpu().numpy()[:1500], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy()[:1500], var_o, rtol=RTOL64, atol=1e-12)
    # packing must not leak into the numbers: any split of the pool gives bitwise the same scores
    for cut in (1, 4097, M - 3):
        a = acq.score_pool(pool[:cut].contiguous())[0].clone()
        b = acq.score_pool(pool[cut:].contiguous())[0].clone()
        assert torch.equal(torch.cat([a, b]), s)
    # fused top-5 == tournament on the score vector == numpy
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    ref = gp_oracle.top_n_distinct(s.cpu().numpy(), 5)
    assert set(int(i) for i in idx) == set(int(i) for i in ref)
    v_t, i_t = eng.topk(s, 5, distinct=True)
    assert np.array_equal(np.sort(idx), np.sort(i_t))
    np.testing.assert_array_equal(eis.reshape(-1), s.cpu().numpy())



# ---------------------------------------------------------------------------- histogram intersection, item form
@pytest.mark.parametrize("fam,h", [(0, 0), (0, 2), (0, 3), (1, 1), (1, 5), (2, 1), (2, 3)])
def test_oa_item_form_matches_general_path_and_oracle(eng, fam, h, monkeypatch):
    """oa=True on the packed path (rank-1 chains + exception items, nbw_model.oa_exc) against the general G path of
    the first-generation kernel ((node, level)^2 gathers over thermometer columns) and against the oracle."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    monkeypatch.setenv("NBW_OA_FORM", "items")         # also where the default would pick the general path
    from nasbowl_b200.kernels import WeisfilerLehman
    n, M = 100, 20_003
    train = synth.generate_cells_host(fam, 9, 0, n)
    y = np.random.RandomState(fam * 7 + h).randn(n)
    gp = _gp(train, y, [WeisfilerLehman(h=h, oa=True)], n_max=train.n_max)
    gp.fit(wl_subtree_candidates=(h,), optimize_lik=False)
    pool = eng.generate_cells(fam, 31, 5000, M)
    acq = GraphExpectedImprovement(gp)
    pm = gp._score_state()
    assert pm.oa_exc is not None and pm.total_labels > pm.n_labels
    assert eng.fused_topk_ok(acq._model(), 5)
    s, mu, var, s64 = [t.clone() for t in acq.score_pool(pool, want64=True)]
    s_l, mu_l, var_l, s64_l = _legacy(lambda: [t.clone() for t in acq.score_pool(pool, want64=True)])
    np.testing.assert_allclose(mu.cpu().numpy(), mu_l.cpu().numpy(), rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(var.cpu().numpy(), var_l.cpu().numpy(), rtol=1e-8, atol=1e-12)
    if train.n_max == 8:
        os.environ["NBW_SCORE_LANES"] = "1"
        try:
            _, mu_n, var_n, _ = [t.clone() for t in acq.score_pool(pool, want64=True)]
        finally:
            os.environ.pop("NBW_SCORE_LANES", None)
        np.testing.assert_allclose(mu.cpu().numpy(), mu_n.cpu().numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var.cpu().numpy(), var_n.cpu().numpy(), rtol=1e-9, atol=1e-13)
    sub = CellBatch.from_records(pool[:1200].cpu().numpy(), train.n_max)
    st = gp_oracle.gp_fit(train, y, h_candidates=(h,), oa=True, optimize_lik=False)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, sub)
    np.testing.assert_allclose(mu.cpu().numpy()[:1200], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy()[:1200], var_o, rtol=RTOL64, atol=1e-12)
    for cut in (1, 4097):
        a = acq.score_pool(pool[:cut].contiguous())[0].clone()
        b = acq.score_pool(pool[cut:].contiguous())[0].clone()
        assert torch.equal(torch.cat([a, b]), s)
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    ref = gp_oracle.top_n_distinct(s.cpu().numpy(), 5)
    assert set(int(i) for i in idx) == set(int(i) for i in ref)


def test_oa_item_form_on_cells_with_many_twins(eng):
    """Hand-made cells whose nodes repeat labels at every level (twins keep rank > 1 down the whole chain), counts
    above the training maximum, labels outside the dictionary."""
    from nasbowl_b200.kernels import WeisfilerLehman
    rng = np.random.RandomState(0)

    def cells(N, ops):
        labels = np.full((N, 8), NO_NODE, dtype=np.uint8)
        succ = np.zeros((N, 8), dtype=np.uint16)
        for i in range(N):
            k = rng.randint(3, 9)
            labels[i, :k] = rng.choice(ops, size=k)
            for u in range(k - 1):
                # many nodes feed the LAST node only: same op + same successors = twins at every level
                if rng.rand() < 0.7:
                    succ[i, u] |= 1 << (k - 1)
                if rng.rand() < 0.3 and u + 1 < k - 1:
                    succ[i, u] |= 1 << rng.randint(u + 1, k - 1)
        return CellBatch(8, labels, succ)

    train = cells(60, [1, 2])
    pool = cells(4000, [1, 2, 3])            # op 3 never seen; pools with up to 7 twins (training max is lower)
    y = rng.randn(60)
    for h, no_twins in ((1, False), (3, False), (3, True)):
        # rank-2 chain items (nbw_model.oa_d2_base) on and off: two ways of writing the same item sums
        if no_twins:
            os.environ["NBW_OA_NO_TWINS"] = "1"
        try:
            gp = _gp(train, y, [WeisfilerLehman(h=h, oa=True)], n_max=8)
            gp.fit(wl_subtree_candidates=(h,), optimize_lik=False)
            assert (gp._score_state().oa_d2_base == 0) == no_twins
        finally:
            os.environ.pop("NBW_OA_NO_TWINS", None)
        mu, var = gp.predict_diag(pool)
        st = gp_oracle.gp_fit(train, y, h_candidates=(h,), oa=True, optimize_lik=False)
        mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool)
        np.testing.assert_allclose(mu.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
        np.testing.assert_allclose(var.numpy(), var_o, rtol=RTOL64, atol=1e-12)
        os.environ["NBW_SCORE_LANES"] = "1"
        try:
            mu_n, var_n = gp.predict_diag(pool)
        finally:
            os.environ.pop("NBW_SCORE_LANES", None)
        np.testing.assert_allclose(mu.numpy(), mu_n.numpy(), rtol=1e-10, atol=1e-12)
        np.testing.assert_allclose(var.numpy(), var_n.numpy(), rtol=1e-9, atol=1e-13)



# ---------------------------------------------------------------------------- size-binned walk of large pools
@pytest.mark.parametrize("fam", [0, 1])
def test_size_binned_order_gives_the_same_scores_and_selection(eng, fam, monkeypatch):
    """Device pools of >= 65536 cells are scored in size-binned order (cell_bin_kernel: counting sort of the cell
    indices by node count).  Scores are per cell and the top-n is (value desc, index asc), so nothing may change:
    bitwise equal scores, identical selection, for one kernel and for a sum of two (compact records of that size:
    test_compact_wire_records_score_identically)."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    n, M = 80, 150_001
    train = synth.generate_cells_host(fam, 2, 0, n)
    y = np.random.RandomState(fam).randn(n)
    pool = eng.generate_cells(fam, 13, 500, M)
    for kernels in ([WeisfilerLehman(h=3)], [WeisfilerLehman(h=1), WeisfilerLehman(h=2)]):
        gp = _gp(train, y, kernels, weights=[1.0 / len(kernels)] * len(kernels))
        gp.fit(wl_subtree_candidates=(kernels[0].h,) if len(kernels) == 1 else (), optimize_lik=False)
        acq = GraphExpectedImprovement(gp)
        s_b = acq.score_pool(pool)[0].clone()
        _, _, idx_b = acq.propose_location(pool, top_n=5)
        monkeypatch.setenv("NBW_NO_BINNING", "1")
        s_p = acq.score_pool(pool)[0].clone()
        _, _, idx_p = acq.propose_location(pool, top_n=5)
        monkeypatch.delenv("NBW_NO_BINNING")
        assert torch.equal(s_b, s_p)
        assert np.array_equal(np.sort(idx_b), np.sort(idx_p))
        ref = gp_oracle.top_n_distinct(s_b.cpu().numpy(), 5)
        assert set(int(i) for i in idx_b) == set(int(i) for i in ref)


def test_packing_edge_cases(eng):
    """Cells of 1..8 nodes in every order, empty records, isolated nodes, pools of 1..40 cells."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(0, 5, 0, 80)
    y = np.random.RandomState(3).randn(80)
    gp = _gp(train, y, [WeisfilerLehman(h=3)])
    gp.fit(wl_subtree_candidates=(3,), optimize_lik=False)
    acq = GraphExpectedImprovement(gp)
    rng = np.random.RandomState(0)
    N = 3000
    labels = np.full((N, 8), NO_NODE, dtype=np.uint8)
    succ = np.zeros((N, 8), dtype=np.uint16)
    for g in range(N):
        k = int(rng.randint(0, 9))             # 0 = an empty record
        labels[g, :k] = rng.randint(0, 6, size=k)   # op code 5 is outside the training vocabulary
        for v in range(k):
            for j in range(v + 1, k):
                if rng.rand() < 0.35:
                    succ[g, v] |= 1 << j
    pool = CellBatch(8, labels, succ)
    st = gp_oracle.gp_fit(train, y, h_candidates=(3,), optimize_lik=False)
    nonempty = pool.n_nodes > 0
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, pool[np.nonzero(nonempty)[0]])
    mu, var = gp.predict_diag(pool)
    np.testing.assert_allclose(mu.numpy()[nonempty], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.numpy()[nonempty], var_o, rtol=RTOL64, atol=1e-12)
    # an empty record falls back to the prior
    prior_var = (np.sqrt(1.0 + st.lik) * st.y_std) ** 2
    np.testing.assert_allclose(mu.numpy()[~nonempty], st.y_mean, rtol=1e-12)
    np.testing.assert_allclose(var.numpy()[~nonempty], prior_var, rtol=1e-12)
    full = acq.score_pool(pool)[0].clone()
    for m in (1, 2, 7, 8, 9, 31, 40):
        part = acq.score_pool(pool[:m])[0]
        assert torch.equal(part, full[:m])
        xs, eis, idx = acq.propose_location(pool[:m], top_n=5)
        ref = gp_oracle.top_n_distinct(eis.reshape(-1), 5)
        if len(ref) >= 5:
            assert set(int(i) for i in idx) == set(int(i) for i in ref)


# ---------------------------------------------------------------------------- sums of kernels
def test_sum_of_kernels_on_the_same_cells(eng, family):
    """SumKernel of linear WL kernels with fixed weights (kernels/combine_kernels.py:36-56) through the fused
    pool path == dense predict() == oracle."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    y = z["y"].astype(np.float64)
    w = [0.3, 0.7]
    gp = _gp(train, y, [WeisfilerLehman(h=1), WeisfilerLehman(h=3)], weights=w)
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit_sum_fixed(train, y, [(1, False), (3, False)], w)
    mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, train, pool)
    mu_d, var_d = gp.predict_diag(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    mu_f, cov_f = gp.predict(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_f.numpy(), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(var_d.numpy(), np.diag(cov_f.numpy()), rtol=1e-8, atol=1e-13)
    acq = GraphExpectedImprovement(gp)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent_sum(st))
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    np.testing.assert_allclose(eis.reshape(-1), ei_o, rtol=RTOL32, atol=1e-9)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(eis.reshape(-1), 5))
    # three and four kernels (the limit of the fused path)
    for hs in ([0, 1, 2], [2, 2, 1, 0]):
        ws = [1.0 / len(hs)] * len(hs)
        gp = _gp(train, y, [WeisfilerLehman(h=h) for h in hs], weights=ws)
        gp.fit(wl_subtree_candidates=(), optimize_lik=False)
        st = gp_oracle.gp_fit_sum_fixed(train, y, [(h, False) for h in hs], ws)
        mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, train, pool)
        mu_d, var_d = gp.predict_diag(pool)
        np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
        np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)


def test_trained_weights_then_fused_pool(eng, family):
    """Kernel weights trained by fit() (gp.py:139-212) feed the fused path like fixed ones."""
    from nasbowl_b200.kernels import WeisfilerLehman
    name, z, train, pool = family
    y = z["y"].astype(np.float64)
    gp = _gp(train, y, [WeisfilerLehman(h=1), WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(), optimize_lik=True, iters=5)
    mu_d, var_d = gp.predict_diag(pool)
    mu_f, cov_f = gp.predict(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_f.numpy(), rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(var_d.numpy(), np.diag(cov_f.numpy()), rtol=1e-8, atol=1e-13)


def test_sum_over_record_slots_normal_and_reduce(eng):
    """An architecture = (normal cell, reduce cell): k = w1 k_WL(normal, normal') + w2 k_WL(reduce, reduce')
    (BASELINE configs[3]).  Two records per candidate, one kernel per record slot."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    n, M = 90, 2500
    normal = synth.generate_cells_host(2, 11, 0, n + M)
    reduce_ = synth.generate_cells_host(2, 12, 50_000, n + M)
    y = np.random.RandomState(4).randn(n)
    tr = CellSlots([normal[:n], reduce_[:n]])
    po = CellSlots([normal[n:], reduce_[n:]])
    w = [0.5, 0.5]
    gp = _gp(tr, y, [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)], weights=w)
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit_sum_fixed([normal[:n], reduce_[:n]], y, [(2, False), (2, False)], w)
    np.testing.assert_allclose(gp.K.numpy(), st.K, rtol=1e-12)
    mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, [normal[:n], reduce_[:n]], [normal[n:], reduce_[n:]])
    mu_d, var_d = gp.predict_diag(po)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    mu_f, cov_f = gp.predict(po[:300])
    np.testing.assert_allclose(mu_d.numpy()[:300], mu_f.numpy(), rtol=1e-8, atol=1e-11)
    # host records (two per candidate, 96 bytes) through propose_location
    acq = GraphExpectedImprovement(gp)
    rec = torch.from_numpy(po.to_records()).pin_memory()
    assert rec.shape == (M, 96)
    xs, eis, idx = acq.propose_location(rec, top_n=5)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent_sum(st))
    np.testing.assert_allclose(eis.reshape(-1), ei_o, rtol=RTOL32, atol=1e-9)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(eis.reshape(-1), 5))
    # tuples of networkx graphs are the same thing
    from nasbowl_b200.cells import OpVocab
    vocab = OpVocab(synth.FAMILY_OPS[2])

    def to_nx(b):
        return b.to_networkx(vocab)
    pairs = list(zip(to_nx(normal[n:n + 40]), to_nx(reduce_[n:n + 40])))
    gp2 = _gp(list(zip(to_nx(normal[:n]), to_nx(reduce_[:n]))), y,
              [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)], weights=w)
    gp2.fit(wl_subtree_candidates=(), optimize_lik=False)
    mu2, var2 = gp2.predict_diag(pairs)
    np.testing.assert_allclose(mu2.numpy(), mu_o[:40], rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var2.numpy(), var_o[:40], rtol=RTOL64, atol=1e-12)


def test_layer_weights_and_undirected_kernels_on_the_fused_path(eng, family):
    from nasbowl_b200.kernels import WeisfilerLehman
    _, z, train, pool = family
    y = z["y"].astype(np.float64)
    lw = [1.0, 0.5, 0.25]
    gp = _gp(train, y, [WeisfilerLehman(h=2, layer_weights=torch.tensor(lw))])
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit_sum_fixed(train, y, [(2, False)], [1.0], layer_weights=[lw])
    mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, train, pool)
    mu_d, var_d = gp.predict_diag(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    # undirected: the pool kernel symmetrises the adjacency itself (kernels/weisfilerlehman.py:127-128)
    gp = _gp(train, y, [WeisfilerLehman(h=2, undirected=True)])
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit(train.to_undirected(), y, h_candidates=(2,), optimize_lik=False)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train.to_undirected(), pool.to_undirected())
    mu_d, var_d = gp.predict_diag(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)
    # a directed and an undirected kernel of one sum read the same record
    gp = _gp(train, y, [WeisfilerLehman(h=2), WeisfilerLehman(h=1, undirected=True)], weights=[0.6, 0.4])
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit_sum_fixed([train, train.to_undirected()], y, [(2, False), (1, False)], [0.6, 0.4])
    mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, [train, train.to_undirected()], [pool, pool.to_undirected()])
    mu_d, var_d = gp.predict_diag(pool)
    np.testing.assert_allclose(mu_d.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_d.numpy(), var_o, rtol=RTOL64, atol=1e-12)


def test_pool_with_wider_cells_than_the_training_set(eng):
    """A GP fitted on <= 8-node cells scores 16-node records: the device state is re-derived on wide records."""
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(0, 5, 0, 60)
    y = np.random.RandomState(1).randn(60)
    gp = _gp(train, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True)
    h, lik, nlml = gp.kernels[0].h, gp.likelihood, float(gp.nlml)
    small = synth.generate_cells_host(0, 6, 500, 50)
    mu_a, var_a = gp.predict_diag(small)
    wide = synth.generate_cells_host(2, 6, 0, 50)             # DARTS cells, ops outside the training vocabulary
    mu_w, var_w = gp.predict_diag(wide)
    assert gp._dev['n_max'] == 16 and gp.kernels[0].h == h and gp.likelihood == lik and float(gp.nlml) == nlml
    mu_b, var_b = gp.predict_diag(small)
    np.testing.assert_allclose(mu_a.numpy(), mu_b.numpy(), rtol=1e-10)
    np.testing.assert_allclose(var_a.numpy(), var_b.numpy(), rtol=1e-10)
    st = gp_oracle.gp_fit(train, y, h_candidates=(1, 2), optimize_lik=True)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train.widen(16), wide)
    np.testing.assert_allclose(mu_w.numpy(), mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var_w.numpy(), var_o, rtol=RTOL64, atol=1e-12)


def test_fused_topk_distinct_rule_and_host_pipelines(eng):
    """NB201 pools repeat architectures: equal scores must be reported once, at their first index
    (acquisitions.py:101-105).  Device, staged-host and mapped-host entry points agree."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    train = synth.generate_cells_host(1, 5, 0, 100)
    y = np.random.RandomState(2).randn(100)
    gp = _gp(train, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(2,), optimize_lik=False)
    acq = GraphExpectedImprovement(gp)
    M = 200_003
    pool = eng.generate_cells(1, 9, 0, M)
    s = acq.score_pool(pool)[0].cpu().numpy()
    for k in (1, 3, 5, 8):
        for distinct in (True, False):
            xs, eis, idx = acq.propose_location(pool, top_n=k, return_distinct=distinct)
            v_t, i_t = eng.topk(torch.from_numpy(s).cuda(), k, distinct=distinct)
            assert np.array_equal(idx, i_t), (k, distinct)
    host = torch.empty((M, 16), dtype=torch.uint8).pin_memory()
    host.copy_(pool.cpu())
    ref_v, ref_i = eng.topk(torch.from_numpy(s).cuda(), 5, distinct=True)
    model = acq._model()
    for mapped, pieces in ((False, 1), (False, 8), (False, 5), (True, 1), (True, 4)):
        cands, hs = eng.propose_host(model, host, 5, True, chunks=pieces, mapped=mapped)
        v, i = eng.cands_read(cands, 5)
        assert np.array_equal(i, ref_i) and np.array_equal(v, ref_v), (mapped, pieces)
        np.testing.assert_array_equal(hs.numpy(), s)
    xs, eis, idx = acq.propose_location(host, top_n=5)
    assert np.array_equal(idx, ref_i)
    np.testing.assert_array_equal(eis.reshape(-1), s)
    # k > 8 falls back to score + tournament
    xs, eis, idx = acq.propose_location(pool, top_n=12)
    v_t, i_t = eng.topk(torch.from_numpy(s).cuda(), 12, distinct=True)
    assert np.array_equal(idx, i_t)


# ---------------------------------------------------------------------------- BASELINE configs[2] / configs[3] sizes
def test_cfg3_nb101_n500(eng):
    """BASELINE configs[2]: NAS-Bench-101-shaped cells, n_train = 500, WL h = 2, EI + top-5 — at the
    training-set size of the config, 6000 candidates against the oracle, 1M through properties."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    n = 500
    train = synth.generate_cells_host(0, 0, 0, n)
    y = torch.randn(n, generator=torch.Generator().manual_seed(0)).numpy().astype(np.float64)
    gp = _gp(train, y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(2,), optimize_lik=True, max_lik=0.01)
    st = gp_oracle.gp_fit(train, y, h_candidates=(2,), optimize_lik=True, max_lik=0.01)
    assert np.float32(gp.likelihood) == np.float32(st.lik)
    M = 1_000_000
    pool = eng.generate_cells(0, 1234, 10_000, M)
    acq = GraphExpectedImprovement(gp)
    s, mu, var, s64 = acq.score_pool(pool, want64=True)
    S = 6000
    sub = CellBatch.from_records(pool[:S].cpu().numpy(), 8)
    mu_o, var_o = gp_oracle.gp_predict_diag(st, train, sub)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent(st))
    np.testing.assert_allclose(mu.cpu().numpy()[:S], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy()[:S], var_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(s64.cpu().numpy()[:S], ei_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(s.cpu().numpy()[:S], ei_o, rtol=RTOL32, atol=1e-9)
    xs, eis, idx = acq.propose_location(pool[:S].contiguous(), top_n=5)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(ei_o.astype(np.float32), 5))
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(s.cpu().numpy(), 5))


def test_cfg4_darts_n1000(eng):
    """BASELINE configs[3]: DARTS-space normal + reduce cells, summed WL kernels, n_train = 1000, h = 2."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.kernels import WeisfilerLehman
    n = 1000
    normal = synth.generate_cells_host(2, 0, 0, n)
    reduce_ = synth.generate_cells_host(2, 1, 0, n)
    y = torch.randn(n, generator=torch.Generator().manual_seed(0)).numpy().astype(np.float64)
    w = [0.5, 0.5]
    gp = _gp(CellSlots([normal, reduce_]), y, [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)], weights=w)
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    st = gp_oracle.gp_fit_sum_fixed([normal, reduce_], y, [(2, False), (2, False)], w)
    np.testing.assert_allclose(gp.K.numpy(), st.K, rtol=1e-12)
    M = 200_000
    pn = eng.generate_cells(2, 1234, 10_000, M)
    pr = eng.generate_cells(2, 4321, 10_000, M)
    pool = torch.cat([pn, pr], dim=1).contiguous()            # two 48-byte records per candidate
    acq = GraphExpectedImprovement(gp)
    s, mu, var, s64 = acq.score_pool(pool, want64=True)
    S = 2000
    subs = [CellBatch.from_records(pn[:S].cpu().numpy(), 16), CellBatch.from_records(pr[:S].cpu().numpy(), 16)]
    mu_o, var_o = gp_oracle.gp_predict_diag_sum(st, [normal, reduce_], subs, chunk=1000)
    ei_o = gp_oracle.expected_improvement(mu_o, var_o, gp_oracle.incumbent_sum(st))
    np.testing.assert_allclose(mu.cpu().numpy()[:S], mu_o, rtol=RTOL64, atol=1e-9)
    np.testing.assert_allclose(var.cpu().numpy()[:S], var_o, rtol=RTOL64, atol=1e-12)
    np.testing.assert_allclose(s64.cpu().numpy()[:S], ei_o, rtol=RTOL64, atol=1e-12)
    xs, eis, idx = acq.propose_location(pool, top_n=5)
    assert set(int(i) for i in idx) == set(int(i) for i in gp_oracle.top_n_distinct(s.cpu().numpy(), 5))


# ---------------------------------------------------------------------------- Gram epilogues / symmetric schedule
def test_gram_output_modes_and_symmetric_schedule(eng):
    """nbw_gram_i8_ex: u8 / s16 (saturating) / cosine-normalised fp32 epilogues and the upper-tiles-only
    schedule + mirror pass are bit-identical to the int32 product (vertex_histogram.py:243)."""
    rng = np.random.RandomState(5)
    for (m, k, hi) in [(1, 128, 3), (300, 256, 3), (513, 384, 4), (1100, 640, 2), (2049, 128, 9)]:
        A = torch.from_numpy(rng.randint(0, hi, size=(m, k)).astype(np.int8)).cuda()
        ref = A.cpu().numpy().astype(np.int64) @ A.cpu().numpy().astype(np.int64).T
        assert np.array_equal(eng.gram_ex(A, A, 0, k, mode="s32").cpu().numpy(), ref)
        ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
        u8 = eng.gram_ex(A, A, 0, k, mode="u8", overflow=ovf).cpu().numpy()
        assert np.array_equal(u8, np.clip(ref, 0, 255).astype(np.uint8))
        assert bool(ovf.item()) == bool((ref > 255).any())
        s16 = eng.gram_ex(A, A, 0, k, mode="s16").cpu().numpy()
        assert np.array_equal(s16, np.clip(ref, -32768, 32767).astype(np.int16))
        d = np.maximum(np.diag(ref), 1).astype(np.float64)
        sc = torch.from_numpy((1.0 / np.sqrt(d)).astype(np.float32)).cuda()
        f32 = eng.gram_ex(A, A, 0, k, mode="f32norm", scale_a=sc, scale_b=sc).cpu().numpy()
        np.testing.assert_allclose(f32, ref * np.outer(sc.cpu().numpy(), sc.cpu().numpy()), rtol=3e-7)
        # symmetric schedule in row blocks of 512 + mirror == full product
        ld = (m + 15) // 16 * 16
        full = torch.zeros((m, ld), dtype=torch.uint8, device="cuda")
        for r0 in range(0, m, 512):
            rows = min(512, m - r0)
            eng.gram_ex(A[r0:r0 + rows], A, 0, k, mode="u8", upper_only=True, row_offset=r0, out=full[r0:r0 + rows])
        upper = full[:, :m].cpu().numpy()
        iu = np.triu_indices(m)
        assert np.array_equal(upper[iu], np.clip(ref, 0, 255).astype(np.uint8)[iu])
        eng.sym_mirror(full)
        assert np.array_equal(full[:, :m].cpu().numpy(), np.clip(ref, 0, 255).astype(np.uint8))
    # negative operands through the signed narrow path
    A = torch.from_numpy(rng.randint(-128, 128, size=(200, 256)).astype(np.int8)).cuda()
    B = torch.from_numpy(rng.randint(-128, 128, size=(333, 256)).astype(np.int8)).cuda()
    ref = A.cpu().numpy().astype(np.int64) @ B.cpu().numpy().astype(np.int64).T
    ovf = torch.zeros(1, dtype=torch.int32, device="cuda")
    s16 = eng.gram_ex(A, B, 0, 256, mode="s16", overflow=ovf).cpu().numpy()
    assert np.array_equal(s16, np.clip(ref, -32768, 32767).astype(np.int16)) and int(ovf.item()) == 1


def test_gram_stress_small(eng):
    """The configs[4] driver (Engine.gram_stress) at a size the oracle finishes: u8 result == oracle Gram."""
    from oracle import wl_oracle
    N = 3000
    cells = eng.generate_cells(0, 2024, 0, N)
    fit = eng.wl_fit(cells, 8, 3, False)
    res = eng.gram_stress(fit, fit.k_end(3), 1024)
    assert all(res["checks"].values()), res["checks"]
    assert 0.5 < res["computed_fraction"] < 0.75
    K_o = wl_oracle.gram_raw(CellBatch.from_records(cells.cpu().numpy(), 8), 3)
    assert np.array_equal(res["result"][:N, :N].cpu().numpy().astype(np.int64), K_o)


# ---------------------------------------------------------------------------- pool de-duplication (row f1)
@pytest.mark.parametrize("fam", [0, 1, 2])
@pytest.mark.parametrize("labeled", [True, False])
def test_pool_dedup_matches_host_twin_and_networkx(eng, fam, labeled):
    """nbw_pool_dedup (allow_isomorphism=False of generate_test_graphs.py:497-520): the keep-first mask equals the
    host twin's, and on a small pool the classes are exactly networkx's isomorphism classes."""
    import networkx as nx
    from nasbowl_b200.cells import OpVocab
    N = 1500 if fam < 2 else 300
    batch = synth.generate_cells_host(fam, 21, 0, N)
    n_max = batch.n_max
    cells = eng.upload_cells(batch)
    for n_prot in (0, 40):
        keep, n_classes, n_kept = eng.pool_dedup(cells, n_max, labeled=labeled, n_protected=n_prot)
        ref = synth.dedup_keep_first_host(batch, labeled=labeled, n_protected=n_prot)
        got = keep.cpu().numpy().astype(bool)
        assert np.array_equal(got[n_prot:], ref)
        assert n_kept == int(ref.sum()) and n_classes == len(set(synth.isomorphism_classes_host(batch, labeled)))
    # networkx on the first 120 cells: same signature <=> isomorphic
    M = 120
    keep, _, _ = eng.pool_dedup(cells[:M].contiguous(), n_max, labeled=labeled)
    k = keep.cpu().numpy().astype(bool)
    gs = batch[:M].to_networkx(OpVocab(synth.FAMILY_OPS[fam]))
    nm = (lambda a, c: a['op_name'] == c['op_name']) if labeled else None
    reps = [i for i in range(M) if k[i]]
    for i in range(M):
        iso_earlier = any(nx.is_isomorphic(gs[i], gs[j], node_match=nm) for j in reps if j < i)
        assert k[i] == (not iso_earlier), (i, k[i], iso_earlier)


def test_mutation_pool_has_no_isomorphic_members(eng):
    """pools.mutation_pool: distinct children that are not isomorphic to a parent, topped up with random cells."""
    import networkx as nx
    from nasbowl_b200 import pools
    from nasbowl_b200.cells import OpVocab
    parents_b = synth.generate_cells_host(0, 3, 0, 10)
    parents = eng.upload_cells(parents_b)
    pool = pools.mutation_pool(eng, parents, pool_size=120, seed=5, labeled=True)
    assert pool.shape == (120, 16)
    b = CellBatch.from_records(pool.cpu().numpy(), 8)
    vocab = OpVocab(synth.FAMILY_OPS[0])
    kids = b[:60].to_networkx(vocab)
    pg = parents_b.to_networkx(vocab)
    nm = lambda a, c: a['op_name'] == c['op_name']
    for i, g in enumerate(kids):
        assert not any(nx.is_isomorphic(g, p, node_match=nm) for p in pg)
        assert not any(nx.is_isomorphic(g, kids[j], node_match=nm) for j in range(i))
    # the reference's structure-only rule leaves far fewer children per batch, the pool is still filled
    pool2 = pools.mutation_pool(eng, parents, pool_size=40, seed=5, labeled=False)
    assert pool2.shape == (40, 16)
    # allow_isomorphism=True: plain mutation batch
    pool3 = pools.mutation_pool(eng, parents, pool_size=40, seed=5, allow_isomorphism=True)
    assert torch.equal(pool3[:20], eng.mutate_cells(0, parents, 5, 0, 64)[:20])


# ---------------------------------------------------------------------------- compact wire records
def test_compact_wire_records_score_identically(eng):
    """nbw_model.compact: 8-byte records (3-bit palette indices + the upper triangle of the adjacency) are expanded
    in the kernel prologue: bitwise the same scores and selection as the 16-byte records, resident or from host
    memory, one or two record slots."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.cells import CompactRecords
    from nasbowl_b200.kernels import WeisfilerLehman
    for fam in (0, 1):
        train = synth.generate_cells_host(fam, 5, 0, 100)
        y = np.random.RandomState(fam).randn(100)
        gp = _gp(train, y, [WeisfilerLehman(h=3)])
        gp.fit(wl_subtree_candidates=(3,), optimize_lik=False)
        acq = GraphExpectedImprovement(gp)
        M = 100_003
        pool_dev = eng.generate_cells(fam, 7, 0, M)
        batch = CellBatch.from_records(pool_dev.cpu().numpy(), 8)
        rec, pal = batch.to_compact()
        assert rec.shape == (M, 8)
        full = acq.score_pool(pool_dev)[0].clone()
        xs, eis, idx = acq.propose_location(pool_dev, top_n=5)
        host = torch.from_numpy(rec).pin_memory()
        xs_c, eis_c, idx_c = acq.propose_location(CompactRecords(host, [pal]), top_n=5)
        np.testing.assert_array_equal(eis_c.reshape(-1), full.cpu().numpy())
        assert np.array_equal(idx, idx_c)
        xs_d, eis_d, idx_d = acq.propose_location(CompactRecords(host.cuda(), [pal]), top_n=5)      # resident compact pool
        np.testing.assert_array_equal(eis_d.reshape(-1), full.cpu().numpy())
        assert np.array_equal(idx, idx_d)
    # two slots
    a, b = synth.generate_cells_host(0, 11, 0, 2080), synth.generate_cells_host(1, 12, 0, 2080)
    y = np.random.RandomState(4).randn(80)
    gp = _gp(CellSlots([a[:80], b[:80]]), y, [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)], weights=[0.5, 0.5])
    gp.fit(wl_subtree_candidates=(), optimize_lik=False)
    acq = GraphExpectedImprovement(gp)
    po = CellSlots([a[80:], b[80:]])
    mu_f, var_f = gp.predict_diag(po)
    ra, pa = a[80:].to_compact()
    rb, pb = b[80:].to_compact()
    host = torch.from_numpy(np.concatenate([ra, rb], axis=1)).pin_memory()
    xs, eis_c, idx_c = acq.propose_location(CompactRecords(host, [pa, pb]), top_n=5)
    xs, eis_f, idx_f = acq.propose_location(torch.from_numpy(po.to_records()).pin_memory(), top_n=5)
    np.testing.assert_array_equal(eis_c, eis_f)
    assert np.array_equal(idx_c, idx_f)
    # what does not fit the wire form is refused on the host
    many = CellBatch(8, np.arange(8, dtype=np.uint8)[None, :], np.zeros((1, 8), dtype=np.uint16))   # eight distinct ops
    with pytest.raises(ValueError):
        many.to_compact()
    bad = CellBatch(8, np.array([[0, 1, 255, 255, 255, 255, 255, 255]], dtype=np.uint8),
                    np.array([[0, 1, 0, 0, 0, 0, 0, 0]], dtype=np.uint16))  # edge 1 -> 0
    with pytest.raises(ValueError):
        bad.to_compact()


def test_compact_24_byte_records_of_16_node_cells(eng):
    """nbw_model.compact for n_max = 16: sixteen 4-bit palette indices + the 120-bit upper triangle in 24 bytes.  Host
    round trip of the packing, then bitwise the same scores and selection as the 48-byte records: one kernel and the
    normal | reduce sum, resident and from pinned host memory."""
    from nasbowl_b200.bayesopt import GraphExpectedImprovement
    from nasbowl_b200.cells import CompactRecords
    from nasbowl_b200.kernels import WeisfilerLehman
    M = 40_007
    a, b = synth.generate_cells_host(2, 21, 0, M + 60), synth.generate_cells_host(2, 22, 0, M + 60)
    rec, pal = a.to_compact()
    assert rec.shape == (M + 60, 24) and pal.shape == (16,)
    # decode on the host: labels and successor masks come back exactly
    w = rec.view(np.uint64).reshape(-1, 3)
    for v in (0, 3, 9, 14, 15):
        idx = (w[:, 0] >> np.uint64(4 * v)) & np.uint64(15)
        lab = np.where(idx == 15, NO_NODE, pal[np.minimum(idx, 14).astype(np.int64)])
        assert np.array_equal(lab.astype(np.uint8), a.labels[:, v])
        if v < 15:
            off, n_bits = v * (31 - v) // 2, 15 - v
            tri = [int(w[i, 1]) | (int(w[i, 2]) << 64) for i in range(200)]
            row = np.array([((t >> off) & ((1 << n_bits) - 1)) << (v + 1) for t in tri], dtype=np.uint16)
            assert np.array_equal(row, a.succ[:200, v])
    y = np.random.RandomState(8).randn(60)
    # one kernel
    gp = _gp(a[:60], y, [WeisfilerLehman(h=2)])
    gp.fit(wl_subtree_candidates=(2,), optimize_lik=False)
    acq = GraphExpectedImprovement(gp)
    pool = a[60:]
    full = acq.score_pool(pool)[0].clone()
    _, _, idx_f = acq.propose_location(pool, top_n=5)
    rec_p, pal_p = pool.to_compact()
    host = torch.from_numpy(rec_p).pin_memory()
    for data in (host, host.cuda()):
        _, eis, idx_c = acq.propose_location(CompactRecords(data, [pal_p]), top_n=5)
        np.testing.assert_array_equal(eis.reshape(-1), full.cpu().numpy())
        assert np.array_equal(idx_f, idx_c)
    # normal | reduce
    gp2 = _gp(CellSlots([a[:60], b[:60]]), y, [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=1, slot=1)], weights=[0.5, 0.5])
    gp2.fit(wl_subtree_candidates=(), optimize_lik=False)
    acq2 = GraphExpectedImprovement(gp2)
    slots = CellSlots([a[60:], b[60:]])
    _, eis_f, idx_f = acq2.propose_location(slots, top_n=5)
    ra, pa = a[60:].to_compact()
    rb, pb = b[60:].to_compact()
    both = torch.from_numpy(np.concatenate([ra, rb], axis=1)).pin_memory()
    _, eis_c, idx_c = acq2.propose_location(CompactRecords(both, [pa, pb]), top_n=5)
    np.testing.assert_array_equal(eis_c, eis_f)
    assert np.array_equal(idx_f, idx_c)
