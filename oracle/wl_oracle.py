"""TEST INFRASTRUCTURE — CPU restatement of the reference's WL feature extractor.

Not part of the product: only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline /
reference arm import this module.  It restates, on packed cell batches
(nasbowl_b200/cells.py), what the reference computes with Python dicts and strings:

* WL relabelling            grakel_replace/weisfeiler_lehman.py:213-296
* per-level histograms      grakel_replace/vertex_histogram.py:98-209
* per-level Gram + sum      grakel_replace/vertex_histogram.py:211-259, weisfeiler_lehman.py:316-327

Parity pinned: tests/test_oracle_golden.py checks every function here against golden vectors
produced by the *unmodified* reference (oracle/make_golden.py -> tests/golden/*.npz).

Two restatements of the relabelling are provided:
  wl_labels_reference_order : string credentials, ids numbered exactly like the reference
                              (sorted strings, running counter) -- pure Python, small inputs.
  wl_labels_fast            : numpy row-unique on (own, sorted successors) tuples -- same
                              partition of nodes into label classes, different id order.
"""
from __future__ import annotations

import numpy as np

from nasbowl_b200.cells import NO_NODE, CellBatch


def _successors(mask: int, n_max: int):
    return [j for j in range(n_max) if (mask >> j) & 1]


def _endpoint_mask(batch: CellBatch) -> np.ndarray:
    """(N, n_max) bool: node is an endpoint of at least one edge.

    GraKeL's edge dictionary has a key for every endpoint, and WL levels >= 1 iterate over
    those keys only (weisfeiler_lehman.py:265); level 0 counts every labelled node
    (vertex_histogram.py:156).  SURVEY.md Appendix A, Q4."""
    s = batch.succ.astype(np.uint32)
    has_out = s != 0
    incoming = np.bitwise_or.reduce(s, axis=1)                      # (N,)
    has_in = ((incoming[:, None] >> np.arange(batch.n_max)[None, :]) & 1).astype(bool)
    return (has_out | has_in) & (batch.labels != NO_NODE)


def wl_labels_reference_order(batch: CellBatch, h: int, op_names=None):
    """Per-level label ids numbered exactly like the reference.

    Returns (ids, dims): ids[l] is an (N, n_max) int64 array (-1 = node not counted at that
    level), dims[l] = (first id, one-past-last id) of level l.

    Level 0: id = rank of the op *name* in sorted(distinct names) (weisfeiler_lehman.py:226-229).
    Level l>=1: credential string  str(own) + "," + str(sorted(successor ids))  (:266-267);
    new ids = running counter over sorted(set(credentials)) (:271-274) -- lexicographic string
    order, so "10,[..]" < "5,[..]".
    """
    N, n_max = batch.labels.shape
    present = batch.labels != NO_NODE
    codes = sorted(set(int(c) for c in batch.labels[present]))
    if op_names is not None:
        order = sorted(codes, key=lambda c: op_names[c])
    else:
        order = codes
    code_to_id = {c: i for i, c in enumerate(order)}
    ids0 = np.full((N, n_max), -1, dtype=np.int64)
    for g in range(N):
        for v in range(n_max):
            if present[g, v]:
                ids0[g, v] = code_to_id[int(batch.labels[g, v])]
    ids, dims = [ids0], [(0, len(order))]
    label_count = len(order)
    endpoint = _endpoint_mask(batch)
    cur = ids0
    for _level in range(1, h + 1):
        cred = {}
        label_set = set()
        for g in range(N):
            for v in range(n_max):
                if not endpoint[g, v]:
                    continue
                nb = sorted(int(cur[g, j]) for j in _successors(int(batch.succ[g, v]), n_max))
                c = str(int(cur[g, v])) + "," + str(nb)
                cred[(g, v)] = c
                label_set.add(c)
        inv = {}
        start = label_count
        for c in sorted(label_set):
            inv[c] = label_count
            label_count += 1
        nxt = np.full((N, n_max), -1, dtype=np.int64)
        for (g, v), c in cred.items():
            nxt[g, v] = inv[c]
        ids.append(nxt)
        dims.append((start, label_count))
        cur = nxt
    return ids, dims


def wl_embed_reference_order(train: CellBatch, pool: CellBatch, h: int, op_names=None, layout: str = "trimmed"):
    """Histogram embeddings in the reference's feature order: X [n, W] of the training cells and
    Y [M, W] of `pool` over the training vocabulary.

    Training pass = wl_labels_reference_order (weisfeiler_lehman.py:223-296).  Transform pass
    (:364-516), ONE pool cell at a time as GraphGP.dmu_dphi calls it (gp.py:358-362): a node's
    credential is looked up in the training dictionary of its level; unknown credentials are numbered
    feature_dims[l+1] + rank in sorted(unknown credentials of that cell) (:452-466).

    layout:
      "trimmed"  W = sum D_l: unseen labels dropped (feature_value, kernels/weisfilerlehman.py:259-261).
      "padded"   W = sum (D_l + 1): each level block followed by one all-zero spare column — the
                 shape of forward_t's embeddings (vertex_histogram.py:184) with unseen labels dropped.
      "forward_t" exactly what forward_t hands to autograd (kernels/weisfilerlehman.py:206-215): the
                 level block of a cell with u >= 1 unknown labels is D_l + u wide (unknown label of
                 rank r sits in column D_l + r, so rank 0 lands in the spare column), the blocks are
                 concatenated and only THEN cut to the training width — a cell with two or more
                 unknown labels at a level shifts all its deeper levels.  Identical to "padded" for
                 cells without unknown labels (every consumer passes the training cells)."""
    assert layout in ("trimmed", "padded", "forward_t")
    ids_t, dims = wl_labels_reference_order(train, h, op_names)
    D = dims[-1][1]
    n_max = train.labels.shape[1]
    assert pool.labels.shape[1] == n_max
    X = np.zeros((len(train), D))
    for l in range(h + 1):
        for g in range(len(train)):
            for v in range(n_max):
                if ids_t[l][g, v] >= 0:
                    X[g, ids_t[l][g, v]] += 1
    # training dictionaries, rebuilt from the labelled training nodes
    present_t = train.labels != NO_NODE
    name_of = (lambda c: op_names[c]) if op_names is not None else (lambda c: c)
    level0 = {int(train.labels[g, v]): int(ids_t[0][g, v]) for g in range(len(train)) for v in range(n_max)
              if present_t[g, v]}
    cred_maps = []
    end_t = _endpoint_mask(train)
    for l in range(1, h + 1):
        m = {}
        for g in range(len(train)):
            for v in range(n_max):
                if end_t[g, v]:
                    nb = sorted(int(ids_t[l - 1][g, j]) for j in _successors(int(train.succ[g, v]), n_max))
                    m[str(int(ids_t[l - 1][g, v])) + "," + str(nb)] = int(ids_t[l][g, v])
        cred_maps.append(m)
    M = len(pool)
    present_p = pool.labels != NO_NODE
    end_p = _endpoint_mask(pool)
    rows = []
    for g in range(M):
        blocks = []
        # level 0
        unknown = sorted({name_of(int(pool.labels[g, v])) for v in range(n_max)
                          if present_p[g, v] and int(pool.labels[g, v]) not in level0})
        cur = {}
        seen_cnt = np.zeros(dims[0][1] - dims[0][0])
        unk_cnt = np.zeros(len(unknown))
        for v in range(n_max):
            if not present_p[g, v]:
                continue
            c = int(pool.labels[g, v])
            if c in level0:
                cur[v] = level0[c]
                seen_cnt[cur[v] - dims[0][0]] += 1
            else:
                r = unknown.index(name_of(c))
                cur[v] = dims[0][1] + r
                unk_cnt[r] += 1
        blocks.append((seen_cnt, unk_cnt))
        for l in range(1, h + 1):
            creds = {}
            for v in range(n_max):
                if end_p[g, v]:
                    nb = sorted(cur[j] for j in _successors(int(pool.succ[g, v]), n_max))
                    creds[v] = str(cur[v]) + "," + str(nb)
            unknown = sorted({c for c in creds.values() if c not in cred_maps[l - 1]})
            seen_cnt = np.zeros(dims[l][1] - dims[l][0])
            unk_cnt = np.zeros(len(unknown))
            nxt = {}
            for v, c in creds.items():
                if c in cred_maps[l - 1]:
                    nxt[v] = cred_maps[l - 1][c]
                    seen_cnt[nxt[v] - dims[l][0]] += 1
                else:
                    r = unknown.index(c)
                    nxt[v] = dims[l][1] + r
                    unk_cnt[r] += 1
            blocks.append((seen_cnt, unk_cnt))
            cur = nxt
        if layout == "trimmed":
            rows.append(np.concatenate([b[0] for b in blocks]))
        elif layout == "padded":
            rows.append(np.concatenate([np.concatenate([b[0], [0.0]]) for b in blocks]))
        else:
            full = np.concatenate([np.concatenate([b[0], b[1] if len(b[1]) else [0.0]]) for b in blocks])
            rows.append(full[:D + h + 1])
    Y = np.array(rows).reshape(M, -1)
    if layout != "trimmed":
        X = np.concatenate([np.concatenate([X[:, a:b], np.zeros((len(train), 1))], axis=1) for (a, b) in dims], axis=1)
    return X, Y


def wl_labels_fast(batch: CellBatch, h: int):
    """Same partition as wl_labels_reference_order, ids local to each level (0..D_l-1) in
    ascending tuple order.  Vectorised: one np.unique over (own, sorted successor ids) rows."""
    N, n_max = batch.labels.shape
    present = batch.labels != NO_NODE
    lab = batch.labels.astype(np.int64)
    uniq0, inv0 = np.unique(lab[present], return_inverse=True)
    ids0 = np.full((N, n_max), -1, dtype=np.int64)
    ids0[present] = inv0
    ids, dims = [ids0], [len(uniq0)]
    endpoint = _endpoint_mask(batch)
    adj = ((batch.succ.astype(np.uint32)[:, :, None] >> np.arange(n_max)[None, None, :]) & 1).astype(bool)
    cur = ids0
    BIG = np.iinfo(np.int64).max
    for _level in range(1, h + 1):
        nb = np.where(adj, cur[:, None, :], BIG)            # (N, v, j): successor ids or BIG
        nb = np.sort(nb, axis=2)
        rows = np.concatenate([cur[:, :, None], nb], axis=2)[endpoint]   # (n_ep, 1+n_max)
        nxt = np.full((N, n_max), -1, dtype=np.int64)
        if rows.shape[0]:
            uniq, inv = np.unique(rows, axis=0, return_inverse=True)
            nxt[endpoint] = inv.reshape(-1)
            dims.append(uniq.shape[0])
        else:
            dims.append(0)
        ids.append(nxt)
        cur = nxt
    return ids, dims


def histogram(ids_l: np.ndarray, start: int, stop: int) -> np.ndarray:
    """Dense per-level vertex histogram Phi_l [N, stop-start] (vertex_histogram.py:156-201),
    columns in id order (the `requires_ordered_features` layout without the spurious last
    column)."""
    N = ids_l.shape[0]
    phi = np.zeros((N, stop - start), dtype=np.int64)
    g, v = np.nonzero(ids_l >= 0)
    np.add.at(phi, (g, ids_l[g, v] - start), 1)
    return phi


def level_gram(phi_a: np.ndarray, phi_b: np.ndarray, oa: bool) -> np.ndarray:
    """K_l between two histogram blocks: linear Phi Phi^T (vertex_histogram.py:243,254) or
    optimal-assignment / histogram intersection sum_v min (:233-238,:245-249)."""
    if not oa:
        return phi_a @ phi_b.T
    out = np.empty((phi_a.shape[0], phi_b.shape[0]), dtype=np.int64)
    for i in range(phi_a.shape[0]):
        out[i] = np.minimum(phi_a[i][None, :], phi_b).sum(axis=1)
    return out


def wl_features(batch: CellBatch, h: int, fast: bool = True, op_names=None):
    """List of per-level dense histograms [Phi_0..Phi_h] over the batch's own vocabulary."""
    if fast:
        ids, dims = wl_labels_fast(batch, h)
        return [histogram(ids[l], 0, dims[l]) for l in range(h + 1)]
    ids, dims = wl_labels_reference_order(batch, h, op_names)
    return [histogram(ids[l], dims[l][0], dims[l][1]) for l in range(h + 1)]


def gram_raw_levels(batch: CellBatch, h: int, oa: bool = False, fast: bool = True) -> np.ndarray:
    """(h+1, N, N) int64: K_l for l = 0..h.  K_raw(h') = prefix sum over l <= h' with unit
    layer weights (weisfeiler_lehman.py:126-129,316-327); never normalised (:353-362)."""
    phis = wl_features(batch, h, fast=fast)
    return np.stack([level_gram(p, p, oa) for p in phis])


def gram_raw(batch: CellBatch, h: int, oa: bool = False, fast: bool = True) -> np.ndarray:
    """K_raw = sum_{l<=h} K_l  -- what WeisfilerLehman.fit_transform returns
    (kernels/weisfilerlehman.py:122-152)."""
    return gram_raw_levels(batch, h, oa, fast).sum(axis=0)


def pool_features(train: CellBatch, pool: CellBatch, h: int, oa: bool = False):
    """Joint WL on train ∪ pool, the way GraphGP.predict does it (bayesopt/gp.py:255-270).

    Returns (K_s_raw [n, M], kss_raw [M], ktt_raw [n]): the raw cross Gram, the pool's own
    self-similarities (which include labels unseen in training) and the training
    self-similarities.  Never forms the M x M block."""
    n, M = len(train), len(pool)
    joint = CellBatch.concat([train, pool])
    phis = wl_features(joint, h, fast=True)
    K_s = np.zeros((n, M), dtype=np.int64)
    kss = np.zeros(M, dtype=np.int64)
    ktt = np.zeros(n, dtype=np.int64)
    for p in phis:
        pt, pp = p[:n], p[n:]
        if oa:
            K_s += level_gram(pt, pp, True)
            kss += pp.sum(axis=1)
            ktt += pt.sum(axis=1)
        else:
            K_s += pt @ pp.T
            kss += (pp * pp).sum(axis=1)
            ktt += (pt * pt).sum(axis=1)
    return K_s, kss, ktt


def pool_features_weighted(train: CellBatch, pool: CellBatch, h: int, oa: bool, level_weights):
    """pool_features with one weight per WL level (weisfeiler_lehman.py:316-327: K_raw = sum_l w_l K_l):
    float64 (K_s [n, M], kss [M], ktt [n])."""
    n, M = len(train), len(pool)
    joint = CellBatch.concat([train, pool])
    phis = wl_features(joint, h, fast=True)
    K_s = np.zeros((n, M))
    kss = np.zeros(M)
    ktt = np.zeros(n)
    for l, p in enumerate(phis):
        w = float(level_weights[l])
        pt, pp = p[:n], p[n:]
        if oa:
            K_s += w * level_gram(pt, pp, True)
            kss += w * pp.sum(axis=1)
            ktt += w * pt.sum(axis=1)
        else:
            K_s += w * (pt @ pp.T)
            kss += w * (pp * pp).sum(axis=1)
            ktt += w * (pt * pt).sum(axis=1)
    return K_s, kss, ktt
