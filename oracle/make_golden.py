"""TEST INFRASTRUCTURE — generate golden vectors by running the UNMODIFIED reference.

    python oracle/make_golden.py            # writes tests/golden/*.npz  (build container only)

The reference tree (/root/reference) is imported under the shims in oracle/ref_shims; its own
samplers produce the cells (NB101 / NB201: bayesopt/generate_test_graphs.py:543-641; DARTS:
darts/utils.py:11-98 fed with genotypes drawn the way darts_objectives.py:159-198 does), its own
WeisfilerLehman / GraphGP / acquisitions produce the outputs.  The cells are stored in the
packed record layout of nasbowl_b200/cells.py so the fixtures can be replayed on the GPU box,
where the reference tree does not exist.
"""
import os
import random
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from nasbowl_b200.cells import OpVocab, pack_networkx  # noqa: E402
from oracle.ref_import import load_reference  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
DARTS_OPS = ['max_pool_3x3', 'avg_pool_3x3', 'skip_connect', 'sep_conv_3x3', 'sep_conv_5x5',
             'dil_conv_3x3', 'dil_conv_5x5']


def seed_all(s):
    random.seed(s)
    np.random.seed(s)
    torch.manual_seed(s)


def sample_cells(R, family, n):
    if family in ("nasbench101", "nasbench201"):
        return R.gtg.random_sampling(n, benchmark=family)[0]
    from darts.cnn.genotypes import Genotype
    from darts.utils import darts2graph, is_valid_darts
    out = []
    while len(out) < n:
        normal = []
        for i in range(4):
            ops = np.random.choice(range(len(DARTS_OPS)), 4)
            nodes_in = np.random.choice(range(i + 2), 2, replace=False)
            normal.extend([(DARTS_OPS[ops[0]], int(nodes_in[0])), (DARTS_OPS[ops[1]], int(nodes_in[1]))])
        g = Genotype(normal=normal, normal_concat=range(2, 6), reduce=normal, reduce_concat=range(2, 6))
        if is_valid_darts(g):
            out.append(darts2graph(g)[0])
    return out


def ref_levels(R, graphs, h, oa):
    """Per-level reference histograms (ordered feature layout) and the raw Gram."""
    k = R.WeisfilerLehman(h=h, oa=oa, requires_grad=True)
    K = k.fit_transform(graphs, rebuild_model=True)
    phis = [np.asarray(k.kern.X[i].X) for i in range(h + 1)]
    dims = list(k.kern.feature_dims)
    return K.numpy(), phis, dims


def family_case(R, family, seed, n_train=40, n_pool=30):
    seed_all(seed)
    graphs = sample_cells(R, family, n_train + n_pool)
    vocab = OpVocab()
    batch = pack_networkx(graphs, vocab)
    out = {"records": batch.to_records(), "n_max": np.int64(batch.n_max),
           "op_names": np.array(vocab.names), "n_train": np.int64(n_train)}
    train, pool = graphs[:n_train], graphs[n_train:]
    # --- integer part: raw Gram for several depths, both base kernels; histograms at h=3
    for oa in (False, True):
        for h in (0, 1, 2, 3, 5):
            K, phis, dims = ref_levels(R, train, h, oa)
            out["Kraw_h%d_oa%d" % (h, int(oa))] = np.rint(K).astype(np.int64)
            if h == 3 and not oa:
                out["feature_dims_h3"] = np.array(dims, dtype=np.int64)
                kk = R.WeisfilerLehman(h=3, oa=False, requires_grad=True)
                kk.fit_transform(train, rebuild_model=True)
                fm = kk.feature_map(flatten=True)
                out["featmap_h3_ids"] = np.array(list(fm.keys()), dtype=np.int64)
                out["featmap_h3_names"] = np.array(list(fm.values()))
                emb, names = kk.feature_value(pool)
                out["featval_h3_pool"] = np.rint(emb.numpy()).astype(np.int16)
                for l, p in enumerate(phis):
                    out["phi_h3_l%d" % l] = np.rint(p).astype(np.int16)
        # joint (train ∪ pool) Gram at h=2: pins pool-side features incl. unseen labels
        K, _, _ = ref_levels(R, graphs, 2, oa)
        out["Kraw_joint_h2_oa%d" % int(oa)] = np.rint(K).astype(np.int64)
    # --- GP: fit / predict / acquisitions
    y = torch.randn(n_train)
    out["y"] = y.numpy().astype(np.float32)
    for oa in (False, True):
        for opt in (False, True):
            tag = "oa%d_opt%d" % (int(oa), int(opt))
            gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=2, oa=oa)])
            cands = tuple(range(4)) if family == "darts" else tuple(range(5))
            gp.fit(wl_subtree_candidates=cands, optimize_lik=opt, max_lik=0.01)
            with torch.no_grad():
                out["gp_h_" + tag] = np.int64(gp.kernels[0].h)
                out["gp_cands_" + tag] = np.array(cands, dtype=np.int64)
                out["gp_lik_" + tag] = np.float64(gp.likelihood)
                out["gp_nlml_" + tag] = np.float64(gp.nlml) if gp.nlml is not None else np.float64("nan")
                out["gp_K_" + tag] = gp.K.detach().numpy()
                out["gp_Ki_" + tag] = gp.K_i.detach().numpy()
                out["gp_logDetK_" + tag] = np.float64(gp.logDetK)
                mu, cov = gp.predict(pool)
                out["gp_mu_" + tag] = mu.numpy()
                out["gp_cov_" + tag] = cov.numpy()
                # grid values (recomputed here exactly as _grid_search_wl_kernel does, gp.py:521-527)
                from bayesopt.utils import compute_log_marginal_likelihood, compute_pd_inverse
                grid = []
                kk = R.WeisfilerLehman(h=2, oa=oa)
                for hc in cands:
                    kk.change_kernel_params({'h': hc})
                    Kh = kk.fit_transform(train, rebuild_model=True, save_gram_matrix=True)
                    K_i, ld = compute_pd_inverse(Kh, 1e-3)
                    grid.append(float(-compute_log_marginal_likelihood(K_i, ld, gp.y)))
                out["gp_grid_" + tag] = np.array(grid)
                ei = R.GraphExpectedImprovement(gp)
                out["ei_" + tag] = np.array([float(ei.eval(c)) for c in pool])
                out["incumbent_" + tag] = np.float64(ei._get_incumbent())
                if opt:
                    aei = R.GraphExpectedImprovement(gp, augmented_ei=True, in_fill='posterior')
                    out["aei_" + tag] = np.array([float(aei.eval(c)) for c in pool])
                    ucb = R.GraphUpperConfidentBound(gp)
                    ucb.iters = 1
                    out["ucb_" + tag] = np.array([float(ucb.eval(c)) for c in pool])
                    _, eis, idx = R.GraphExpectedImprovement(gp).propose_location(pool, top_n=5)
                    out["top5_" + tag] = np.array(sorted(int(i) for i in idx), dtype=np.int64)
    return out


def appendix_b(R):
    """The hand-checkable 4-graph known-answer test of SURVEY.md Appendix B, full precision."""
    import networkx as nx
    I, O, C3, C1, MP = 'input', 'output', 'conv3x3-bn-relu', 'conv1x1-bn-relu', 'maxpool3x3'

    def mk(labels, edges):
        g = nx.DiGraph()
        for i, l in enumerate(labels):
            g.add_node(i, op_name=l)
        g.add_edges_from(edges)
        return g
    G = [mk([I, C3, O], [(0, 1), (1, 2)]), mk([I, C3, O], [(0, 1), (0, 2), (1, 2)]),
         mk([I, C3, MP, O], [(0, 1), (0, 2), (1, 3), (2, 3)]),
         mk([I, C1, C3, O], [(0, 1), (0, 2), (1, 2), (1, 3), (2, 3)])]
    T = [mk([I, C3, C3, O], [(0, 1), (1, 2), (2, 3)]), mk([I, O], [(0, 1)])]
    vocab = OpVocab()
    batch = pack_networkx(G + T, vocab)
    out = {"records": batch.to_records(), "n_max": np.int64(batch.n_max),
           "op_names": np.array(vocab.names), "n_train": np.int64(4)}
    K, phis, dims = ref_levels(R, G, 2, False)
    out["Kraw_h2_oa0"] = np.rint(K).astype(np.int64)
    out["feature_dims_h2"] = np.array(dims, dtype=np.int64)
    for l, p in enumerate(phis):
        out["phi_h2_l%d" % l] = np.rint(p).astype(np.int16)
    y = torch.tensor([0.5, 0.1, -0.3, 0.9])
    out["y"] = y.numpy()
    gp = R.GraphGP(G, y, [R.WeisfilerLehman(h=2)])
    gp.fit(optimize_lik=False, wl_subtree_candidates=(0, 1, 2))
    with torch.no_grad():
        out["gp_h"] = np.int64(gp.kernels[0].h)
        out["gp_lik"] = np.float64(gp.likelihood)
        out["gp_K"] = gp.K.numpy()
        out["gp_Ki"] = gp.K_i.numpy()
        out["gp_logDetK"] = np.float64(gp.logDetK)
        mu, cov = gp.predict(T)
        out["gp_mu"], out["gp_cov"] = mu.numpy(), cov.numpy()
        ei = R.GraphExpectedImprovement(gp)
        out["ei"] = np.array([float(ei.eval(t)) for t in T])
        out["incumbent"] = np.float64(ei._get_incumbent())
        ucb = R.GraphUpperConfidentBound(gp)
        ucb.iters = 1
        out["ucb"] = np.array([float(ucb.eval(t)) for t in T])
    return out


def cfg1_case(R, seed=11, n_train=50, n_test=400):
    """BASELINE.json configs[0]: nas_regression.py:78-83 on NB101-shaped cells from the reference's
    own random_sampling — fit(optimize_lik=True) with the default h grid 0..4, predict on 400 test
    cells, plus EI of every test cell and the proposed top-5."""
    seed_all(seed)
    graphs = sample_cells(R, "nasbench101", n_train + n_test)
    vocab = OpVocab()
    batch = pack_networkx(graphs, vocab)
    train, test = graphs[:n_train], graphs[n_train:]
    y = torch.randn(n_train)
    gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=2, oa=False)])
    gp.fit(optimize_lik=True)
    out = {"records": batch.to_records(), "n_max": np.int64(batch.n_max), "op_names": np.array(vocab.names),
           "n_train": np.int64(n_train), "y": y.numpy().astype(np.float32)}
    with torch.no_grad():
        mu, cov = gp.predict(test)
        out["gp_h"], out["gp_lik"] = np.int64(gp.kernels[0].h), np.float64(gp.likelihood)
        out["gp_nlml"] = np.float64(gp.nlml)
        out["gp_mu"], out["gp_var"] = mu.numpy(), np.diag(cov.numpy()).copy()
        mu_t, cov_t = gp.predict(train)
        out["gp_mu_train"] = mu_t.numpy()
        ei = R.GraphExpectedImprovement(gp)
        out["ei"] = np.array([float(ei.eval(c)) for c in test])
        out["incumbent"] = np.float64(ei._get_incumbent())
        _, _, idx = R.GraphExpectedImprovement(gp).propose_location(test, top_n=5)
        out["top5"] = np.array(sorted(int(i) for i in idx), dtype=np.int64)
    return out



def baseline_size_case(R, family, seed, n_train, cands, n_pool, n_eval):
    """A fixture at the TRAINING-SET SIZE of BASELINE.json configs[1] / configs[2] (NB201 n=150 with the
    h grid 0..5; NB101 n=500 at h=2), from the unmodified reference: the integer Gram at the chosen depth, the
    marginal-likelihood grid, the fitted noise, the batched posterior over `n_pool` candidates, and the
    reference's own per-candidate EI (`GraphExpectedImprovement.eval`, acquisitions.py:59-80) and
    `propose_location` top-5 (acquisitions.py:94-113) over the first `n_eval` of them (the reference needs
    0.1-1 s per candidate at these sizes, which is what bounds the slice)."""
    from bayesopt.utils import compute_log_marginal_likelihood, compute_pd_inverse
    seed_all(seed)
    graphs = sample_cells(R, family, n_train + n_pool)
    vocab = OpVocab()
    batch = pack_networkx(graphs, vocab)
    train, pool = graphs[:n_train], graphs[n_train:]
    # a target the cells explain (edges + 3x3 operations) plus noise: with pure noise every pool EI sits 4-5 sigma
    # in the tail (1e-8 and below), where the reference's fp32 posterior decides the ranking by rounding
    y = torch.tensor([0.25 * g.number_of_edges() + 0.5 * sum('3x3' in d['op_name'] for _, d in g.nodes(data=True))
                      for g in train], dtype=torch.float32) + 0.1 * torch.randn(n_train)
    out = {"records": batch.to_records(), "n_max": np.int64(batch.n_max), "op_names": np.array(vocab.names),
           "n_train": np.int64(n_train), "y": y.numpy().astype(np.float32),
           "cands": np.array(cands, dtype=np.int64), "n_eval": np.int64(n_eval)}
    gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=2, oa=False)])
    gp.fit(wl_subtree_candidates=cands, optimize_lik=True, max_lik=0.01)
    with torch.no_grad():
        h = int(gp.kernels[0].h)
        out["gp_h"], out["gp_lik"], out["gp_nlml"] = np.int64(h), np.float64(gp.likelihood), np.float64(gp.nlml)
        out["gp_logDetK"] = np.float64(gp.logDetK)
        kk = R.WeisfilerLehman(h=2, oa=False)
        grid = []
        for hc in cands:                                   # as _grid_search_wl_kernel does, gp.py:521-527
            kk.change_kernel_params({'h': hc})
            Kh = kk.fit_transform(train, rebuild_model=True, save_gram_matrix=True)
            K_i, ld = compute_pd_inverse(Kh, 1e-3)
            grid.append(float(-compute_log_marginal_likelihood(K_i, ld, gp.y)))
            if hc == h:
                K = np.rint(Kh.numpy()).astype(np.int64)      # symmetric, entries <= 255 on these sets: the packed
                assert np.array_equal(K, K.T) and 0 <= K.min() and K.max() < 256       # upper triangle as bytes
                out["Kraw_triu"] = K[np.triu_indices(n_train)].astype(np.uint8)
        out["gp_grid"] = np.array(grid)
        mu, cov = gp.predict(pool)
        out["gp_mu"], out["gp_var"] = mu.numpy(), np.diag(cov.numpy()).copy()
        ei = R.GraphExpectedImprovement(gp)
        out["incumbent"] = np.float64(ei._get_incumbent())
        out["ei"] = np.array([float(ei.eval(c)) for c in pool[:n_eval]])
        _, _, idx = R.GraphExpectedImprovement(gp).propose_location(pool[:n_eval], top_n=5)
        out["top5"] = np.array(sorted(int(i) for i in idx), dtype=np.int64)
    return out


def dmu_case(R, n_train=30, n_pool=20):
    """Interpretability gradients (gp.py:296-476) from the unmodified reference: per family and
    base kernel, GraphGP.dmu_dphi over a pool (raw Jacobian-vector products + the embedding they
    belong to) and its three averaged forms."""
    out = {}
    for family, seed in (("nasbench101", 111), ("nasbench201", 211), ("darts", 311)):
        seed_all(seed)
        graphs = sample_cells(R, family, n_train + n_pool)
        vocab = OpVocab()
        batch = pack_networkx(graphs, vocab)
        train, pool = graphs[:n_train], graphs[n_train:]
        y = torch.randn(n_train)
        out[family + "_records"] = batch.to_records()
        out[family + "_n_max"] = np.int64(batch.n_max)
        out[family + "_op_names"] = np.array(vocab.names)
        out[family + "_y"] = y.numpy().astype(np.float32)
        for oa in (False, True):
            tag = "%s_oa%d_" % (family, int(oa))
            # requires_grad=True = ordered feature columns (vertex_histogram.py:160-162), the setting of
            # every interpretability consumer (interpreter.py:23, transfer_nasbench201.py:89)
            gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=2, oa=oa, requires_grad=True)])
            gp.fit(wl_subtree_candidates=(1, 2), optimize_lik=True, max_lik=0.01)
            out[tag + "h"] = np.int64(gp.kernels[0].h)
            out[tag + "lik"] = np.float64(gp.likelihood)
            raw, _, feat = gp.dmu_dphi(pool, average_across_features=False)
            out[tag + "raw"] = torch.cat(raw).detach().numpy()
            out[tag + "feat"] = feat.detach().numpy()
            for nm, kw in (("occ", dict(average_across_occurrences=False)), ("avg", dict(average_across_occurrences=True))):
                a, v, inc = gp.dmu_dphi(pool, **kw)
                out[tag + nm + "_mu"], out[tag + nm + "_var"], out[tag + nm + "_inc"] = \
                    a.detach().numpy(), v.detach().numpy(), inc.detach().numpy()
            a, v, inc = gp.dmu_dphi()
            out[tag + "train_mu"], out[tag + "train_var"], out[tag + "train_inc"] = \
                a.detach().numpy(), v.detach().numpy(), inc.detach().numpy()
    return out


def multi_kernel_case(R, n_train=40, n_pool=20):
    """GraphGP over a SUM of two WL kernels with trainable weights (gp.py:139-212)."""
    out = {}
    for family, seed in (("nasbench101", 121), ("nasbench201", 221)):
        seed_all(seed)
        graphs = sample_cells(R, family, n_train + n_pool)
        vocab = OpVocab()
        batch = pack_networkx(graphs, vocab)
        train, pool = graphs[:n_train], graphs[n_train:]
        y = torch.randn(n_train)
        out[family + "_records"] = batch.to_records()
        out[family + "_n_max"] = np.int64(batch.n_max)
        out[family + "_op_names"] = np.array(vocab.names)
        out[family + "_y"] = y.numpy().astype(np.float32)
        gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=1, oa=False), R.WeisfilerLehman(h=2, oa=True)])
        gp.fit(wl_subtree_candidates=(0, 1, 2), optimize_lik=True, max_lik=0.01)
        with torch.no_grad():
            out[family + "_hs"] = np.array([k.h for k in gp.kernels], dtype=np.int64)
            out[family + "_weights"] = gp.weights.detach().numpy().astype(np.float64)
            out[family + "_lik"] = np.float64(gp.likelihood)
            out[family + "_nlml"] = np.float64(gp.nlml)
            out[family + "_K"] = gp.K.detach().numpy()
            out[family + "_Ki"] = gp.K_i.detach().numpy()
            mu, cov = gp.predict(pool)
            out[family + "_mu"], out[family + "_cov"] = mu.detach().numpy(), cov.detach().numpy()
        # Adam's steps are sign-driven and saturate the clamp; three small SGD steps pin the gradient VALUES
        gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=1, oa=False), R.WeisfilerLehman(h=2, oa=True)])
        gp.fit(iters=3, optimizer='sgd', optimizer_kwargs={'lr': 0.02}, wl_subtree_candidates=(), optimize_lik=True,
               max_lik=0.01)
        out[family + "_sgd_weights"] = gp.weights.detach().numpy().astype(np.float64)
        out[family + "_sgd_lik"] = np.float64(gp.likelihood)
        out[family + "_sgd_nlml"] = np.float64(gp.nlml)
        # trainable WL layer weights (gp.py:157-165,208,218), one kernel: Adam run + three SGD steps
        for tag, kw in (("lw", {}), ("lwsgd", dict(iters=3, optimizer='sgd', optimizer_kwargs={'lr': 0.02}))):
            gp = R.GraphGP(train, y.clone(), [R.WeisfilerLehman(h=2, oa=False)])
            gp.fit(wl_subtree_candidates=(), optimize_lik=True, max_lik=0.01, optimize_wl_layer_weights=True, **kw)
            out[family + "_%s_layer_weights" % tag] = gp.layer_weights.detach().numpy().astype(np.float64)
            out[family + "_%s_nlml" % tag] = np.float64(gp.nlml)
            with torch.no_grad():
                mu, cov = gp.predict(pool)
                out[family + "_%s_mu" % tag] = mu.detach().numpy()
                out[family + "_%s_var" % tag] = np.diag(cov.detach().numpy()).copy()
    return out


def main():
    R = load_reference()
    if len(sys.argv) > 1 and sys.argv[1] == "multi":      # only the multi-kernel fixture
        np.savez_compressed(os.path.join(OUT, "multi_kernel.npz"), **multi_kernel_case(R))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "dmu":        # only the interpretability fixture
        np.savez_compressed(os.path.join(OUT, "dmu_dphi.npz"), **dmu_case(R))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sizes":      # only the BASELINE-size fixtures (minutes of reference time)
        np.savez_compressed(os.path.join(OUT, "cfg2_nb201_n150.npz"),
                            **baseline_size_case(R, "nasbench201", 202, 150, tuple(range(6)), 400, 64))
        np.savez_compressed(os.path.join(OUT, "cfg3_nb101_n500.npz"),
                            **baseline_size_case(R, "nasbench101", 303, 500, (2,), 400, 32))
        return
    if len(sys.argv) > 1 and sys.argv[1] == "sizes-darts":    # configs[3]'s training-set size, one WL kernel (the
        # reference has no two-cells-per-example form: darts_objectives.py:193 keeps the normal cell's graph only)
        np.savez_compressed(os.path.join(OUT, "cfg4_darts_n1000.npz"),
                            **baseline_size_case(R, "darts", 404, 1000, (2,), 200, 16))
        return
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "appendix_b.npz"), **appendix_b(R))
    np.savez_compressed(os.path.join(OUT, "cfg1_nb101.npz"), **cfg1_case(R))
    np.savez_compressed(os.path.join(OUT, "dmu_dphi.npz"), **dmu_case(R))
    np.savez_compressed(os.path.join(OUT, "multi_kernel.npz"), **multi_kernel_case(R))
    for family, seed in (("nasbench101", 101), ("nasbench201", 201), ("darts", 301)):
        case = family_case(R, family, seed)
        np.savez_compressed(os.path.join(OUT, family + ".npz"), **case)
        print(family, "ok:", {k: (v.shape if hasattr(v, "shape") else v) for k, v in list(case.items())[:4]})


if __name__ == "__main__":
    main()
