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
