"""Synthetic cell pools shaped like the reference's samplers (host side).

Same construction rules as the reference, driven by a counter-based RNG so that cell i depends
only on (seed, i) and the CUDA generator (csrc/cellgen.cu, ``nbw_generate_cells``) produces
bit-identical records:

* family 0 — NAS-Bench-101: bayesopt/generate_test_graphs.py:559-603 (7x7 upper-triangular
  adjacency, each of 21 edges ~ Bernoulli(1/2); 5 interior ops uniform over 3; `prune`
  :26-78; reject if no path input->output, 0 edges or > 9 edges).
* family 1 — NAS-Bench-201: :95-177 + :605-624 (6 ops uniform over 5 on the fixed 8-node
  template; `none` nodes deleted, `skip_connect` nodes bypassed; one sweep of dangling-node
  removal; reject empty / edge-less graphs).
* family 2 — DARTS: darts/darts_objectives.py:159-198 (4 towers x 2 ops uniform over 7, two
  distinct inputs among earlier states, both cell inputs used) + darts/utils.py:11-98
  (`add` nodes, skip_connect contraction, removal sweep).

Not reproduced: the reference samplers also drop *duplicates within one pool* (two-node NB101
cells, repeated NB201 op strings), which is a global operation on a host list.
"""
from __future__ import annotations

import numpy as np

from .cells import NO_NODE, CellBatch

NB101_OPS = ["input", "output", "conv1x1-bn-relu", "conv3x3-bn-relu", "maxpool3x3"]
NB201_OPS = ["input", "output", "nor_conv_3x3", "nor_conv_1x1", "avg_pool_3x3"]
DARTS_OPS = ["input1", "input2", "output", "add", "max_pool_3x3", "avg_pool_3x3", "skip_connect",
             "sep_conv_3x3", "sep_conv_5x5", "dil_conv_3x3", "dil_conv_5x5"]
FAMILY_OPS = {0: NB101_OPS, 1: NB201_OPS, 2: DARTS_OPS}
FAMILY_NMAX = {0: 8, 1: 8, 2: 16}
FAMILY_ID = {"nasbench101": 0, "nasbench201": 1, "darts": 2}

_M64 = (1 << 64) - 1
_GOLD = 0x9E3779B97F4A7C15
_GOLD2 = 0xD1B54A32D192ED03


def _mix_a(x):
    x &= _M64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & _M64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & _M64
    x ^= x >> 31
    return x


def _mix_b(x):
    x &= _M64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & _M64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & _M64
    x ^= x >> 33
    return x


def draw(seed: int, index: int, attempt: int, t: int) -> int:
    """64 random bits for (seed, cell index, rejection attempt, draw number)."""
    s = _mix_a(seed + _GOLD * (index + 1))
    return _mix_b(s + _GOLD2 * (attempt * 64 + t + 1))


def _below(x: int, n: int, field: int) -> int:
    """(near-)uniform integer in [0, n) from byte `field` (0..7) of x."""
    return (((x >> (8 * field)) & 0xFF) * n) >> 8


# ------------------------------------------------------------------------------- NB101
_TRIU = [(r, c) for r in range(7) for c in range(r + 1, 7)]


def _nb101_cell(seed, idx):
    attempt = 0
    while True:
        x = draw(seed, idx, attempt, 0)
        y = draw(seed, idx, attempt, 1)
        attempt += 1
        succ = [0] * 7
        for e, (r, c) in enumerate(_TRIU):
            if (x >> e) & 1:
                succ[r] |= 1 << c
        ops = [0] + [2 + _below(y, 3, i) for i in range(5)] + [1]
        fwd = 1
        for v in range(1, 7):
            if any((fwd >> u) & 1 and (succ[u] >> v) & 1 for u in range(v)):
                fwd |= 1 << v
        bwd = 1 << 6
        for v in range(5, -1, -1):
            if succ[v] & bwd:
                bwd |= 1 << v
        keep = fwd & bwd
        if keep == 0:
            continue
        n_edges = sum(bin(succ[v] & keep).count("1") for v in range(7) if (keep >> v) & 1)
        if n_edges == 0 or n_edges > 9:
            continue
        return _compact(ops, succ, keep, 7)


def _compact(ops, succ, keep, n):
    """Renumber the kept nodes 0..k-1 in ascending order of their ids."""
    pos, k = {}, 0
    for v in range(n):
        if (keep >> v) & 1:
            pos[v] = k
            k += 1
    labels, masks = [NO_NODE] * 16, [0] * 16
    for v, p in pos.items():
        labels[p] = ops[v]
        m = 0
        for j in range(n):
            if (succ[v] >> j) & 1 and (keep >> j) & 1:
                m |= 1 << pos[j]
        masks[p] = m
    return labels, masks


# ------------------------------------------------------------------------------- NB201
_T201 = [(0, 1), (0, 2), (0, 4), (1, 3), (1, 5), (2, 6), (3, 6), (4, 7), (5, 7), (6, 7)]


def nb201_from_choice(choice):
    """Packed NB201 cell from its six edge choices (0..2 real ops, 3 skip_connect, 4 none), or None
    when nothing connects input and output (create_nasbench201_graph, generate_test_graphs.py:95-177)."""
    succ0 = [0] * 8
    pred0 = [0] * 8
    for u, v in _T201:
        succ0[u] |= 1 << v
        pred0[v] |= 1 << u
    succ = list(succ0)
    removed = 0
    for i in range(1, 7):
        c = choice[i - 1]
        if c >= 3:
            removed |= 1 << i
            if c == 3:       # bypass: original predecessors x original successors (:151-153)
                for u in range(8):
                    if (pred0[i] >> u) & 1:
                        succ[u] |= succ0[i]
    keep = 0xFF & ~removed
    succ = [(succ[v] & keep) if (keep >> v) & 1 else 0 for v in range(8)]
    indeg = [sum((succ[u] >> v) & 1 for u in range(8)) for v in range(8)]
    drop = 0
    for v in range(8):
        if not (keep >> v) & 1:
            continue
        if v != 0 and indeg[v] == 0:
            drop |= 1 << v
        elif v != 7 and succ[v] == 0:
            drop |= 1 << v
    keep &= ~drop
    succ = [(succ[v] & keep) if (keep >> v) & 1 else 0 for v in range(8)]
    if keep == 0 or not any(succ):
        return None
    ops = [0] + [2 + min(c, 2) for c in choice] + [1]
    return _compact(ops, succ, keep, 8)


def _nb201_cell(seed, idx, want_encoding=False):
    attempt = 0
    while True:
        x = draw(seed, idx, attempt, 0)
        attempt += 1
        choice = [_below(x, 5, i) for i in range(6)]
        cell = nb201_from_choice(choice)
        if cell is None:
            continue
        return (cell, choice + [0, 0]) if want_encoding else cell


# ------------------------------------------------------------------------------- DARTS
def darts_from_genotype(ops8, ins8):
    """Packed DARTS cell from eight (op 0..6, input 0..5) pairs (darts2graph, darts/utils.py:11-98), or
    None when a cell input is unused (is_valid_darts, :148-153).  Inputs may point at a later block:
    the reference's _mutate draws them from all six sources (darts_objectives.py:217-227)."""
    if 0 not in ins8 or 1 not in ins8:
        return None
    n = 15
    ops = [NO_NODE] * n
    ops[0], ops[1], ops[14] = 0, 1, 2
    succ = [0] * n
    for i in range(4):
        ops[3 * i + 2] = 4 + ops8[2 * i]
        ops[3 * i + 3] = 4 + ops8[2 * i + 1]
        ops[3 * i + 4] = 3
        succ[3 * i + 2] |= 1 << (3 * i + 4)
        succ[3 * i + 3] |= 1 << (3 * i + 4)
    for i in range(4):
        for off in range(2):
            k = ins8[2 * i + off]
            src = k if k <= 1 else (k - 2) * 3 + 4
            succ[src] |= 1 << (3 * i + 2 + off)
    for i in range(4):
        succ[3 * i + 4] |= 1 << 14
    keep = (1 << n) - 1
    skip = 4 + 2
    for j in range(n):                             # skip_connect contraction (:53-64)
        if (keep >> j) & 1 and ops[j] == skip:
            out = (succ[j] & -succ[j]).bit_length() - 1
            for u in range(n):
                if (keep >> u) & 1 and (succ[u] >> j) & 1:
                    succ[u] = (succ[u] | (1 << out)) & ~(1 << j)
            succ[j] = 0
            keep &= ~(1 << j)
    for j in range(n):                             # removal sweep, sequential (:67-86)
        if not (keep >> j) & 1:
            continue
        if ops[j] not in (0, 1):
            if not any((keep >> u) & 1 and (succ[u] >> j) & 1 for u in range(n)):
                keep &= ~(1 << j)
                succ[j] = 0
        else:
            if succ[j] & keep == 0:
                keep &= ~(1 << j)
    return _compact(ops, succ, keep, n)


def _darts_cell(seed, idx, want_encoding=False):
    attempt = 0
    while True:
        x = draw(seed, idx, attempt, 0)      # 8 op choices
        y = draw(seed, idx, attempt, 1)      # 8 input choices
        attempt += 1
        ops8, ins8 = [], []
        for i in range(4):
            a = (((y >> (16 * i)) & 0xFF) * (i + 2)) >> 8
            b = (((y >> (16 * i + 8)) & 0xFF) * (i + 1)) >> 8
            if b >= a:
                b += 1
            ops8 += [_below(x, 7, 2 * i), _below(x, 7, 2 * i + 1)]
            ins8 += [a, b]
        cell = darts_from_genotype(ops8, ins8)
        if cell is None:
            continue
        if want_encoding:
            return cell, [v for pair in zip(ops8, ins8) for v in pair]
        return cell


# ------------------------------------------------------------------------------- encoded mutation (NB201, DARTS)
ENC_BYTES = {1: 8, 2: 16}


NB201_CHOICES = ['nor_conv_3x3', 'nor_conv_1x1', 'avg_pool_3x3', 'skip_connect', 'none']
DARTS_CHOICES = ['max_pool_3x3', 'avg_pool_3x3', 'skip_connect', 'sep_conv_3x3', 'sep_conv_5x5', 'dil_conv_3x3',
                 'dil_conv_5x5']


def nb201_arch_string(enc) -> str:
    """The NAS-Bench-201 query string of an encoding — what the reference stores as G.name
    (generate_test_graphs.py:171-176) and hands to the benchmark look-up."""
    o = [NB201_CHOICES[int(c)] for c in enc[:6]]
    return '|%s~0|+|%s~0|%s~1|+|%s~0|%s~1|%s~2|' % tuple(o)


def nb201_encoding(arch_string: str):
    """Inverse of nb201_arch_string (how the reference's mutation reads its parents, :413-414)."""
    ops = [tok[:-2] for tok in arch_string.split('|') if '~' in tok]
    assert len(ops) == 6, arch_string
    return [NB201_CHOICES.index(o) for o in ops] + [0, 0]


def darts_genotype_pairs(enc):
    """[(op_name, input)] x 8: the `normal` list of a DARTS Genotype (darts_objectives.py:159-198)."""
    return [(DARTS_CHOICES[int(enc[2 * k])], int(enc[2 * k + 1])) for k in range(8)]


def darts_encoding(pairs):
    out = []
    for op, src in pairs:
        out += [DARTS_CHOICES.index(op), int(src)]
    assert len(out) == 16
    return out


def generate_encoded_host(family: int, seed: int, first_index: int, n_cells: int):
    """Host twin of nbw_generate_encoded: (CellBatch, encodings uint8 [n, 8 | 16])."""
    assert family in (1, 2)
    n_max = FAMILY_NMAX[family]
    labels = np.full((n_cells, n_max), NO_NODE, dtype=np.uint8)
    succ = np.zeros((n_cells, n_max), dtype=np.uint16)
    enc = np.zeros((n_cells, ENC_BYTES[family]), dtype=np.uint8)
    gen = _nb201_cell if family == 1 else _darts_cell
    for i in range(n_cells):
        (lab, msk), e = gen(seed, first_index + i, want_encoding=True)
        labels[i], succ[i], enc[i] = lab[:n_max], msk[:n_max], e
    return CellBatch(n_max, labels, succ), enc


def mutate_encoded_host(family: int, parent_enc: np.ndarray, seed: int, first_index: int, n_children: int,
                        edits: int = 1):
    """Host twin of nbw_mutate_encoded (rules in csrc/cellgen.cu: generate_test_graphs.py:412-431 for
    NB201, darts/darts_objectives.py:201-245,258-260 for DARTS).  Returns (CellBatch, child encodings)."""
    assert family in (1, 2)
    n_max = FAMILY_NMAX[family]
    labels = np.full((n_children, n_max), NO_NODE, dtype=np.uint8)
    succ = np.zeros((n_children, n_max), dtype=np.uint16)
    enc = np.zeros((n_children, ENC_BYTES[family]), dtype=np.uint8)
    P = len(parent_enc)
    for i in range(n_children):
        idx = first_index + i
        par = [int(v) for v in parent_enc[idx % P]]
        mseed = seed ^ 0x6D75746174650000
        cell = e = None
        for attempt in range(MUTATION_ATTEMPTS):
            if family == 1:
                d0, d1 = draw(mseed, idx, attempt, 0), draw(mseed, idx, attempt, 1)
                choice = par[:6]
                for k in range(6):
                    if _below(d0, 5, k) == 0:
                        pick = _below(d1, 4, k)
                        choice[k] = pick if pick < par[k] else pick + 1
                cell = nb201_from_choice(choice)
                e = choice + [0, 0]
            else:
                ops8, ins8 = par[0::2], par[1::2]
                ok = True
                for ed in range(edits):
                    d0, d1 = draw(mseed, idx, attempt, 2 * ed), draw(mseed, idx, attempt, 2 * ed + 1)
                    coin, pos = _below(d0, 2, 0), _below(d0, 8, 1)
                    if coin == 1:
                        ops8[pos] = _below(d0, 7, 2)
                    else:
                        sibling, found = ins8[pos ^ 1], False
                        for t in range(13):
                            c = _below(d0, 6, 3 + t) if t < 5 else _below(d1, 6, t - 5)
                            if c != sibling and c != ins8[pos]:
                                ins8[pos], found = c, True
                                break
                        if not found:
                            ok = False
                            break
                cell = darts_from_genotype(ops8, ins8) if ok else None
                e = [v for pair in zip(ops8, ins8) for v in pair]
            if cell is not None:
                break
        if cell is None:
            gen = _nb201_cell if family == 1 else _darts_cell       # 64 failed attempts: a random cell
            cell, e = gen(seed ^ 0x66616C6C6261636B, idx, want_encoding=True)
        labels[i], succ[i], enc[i] = cell[0][:n_max], cell[1][:n_max], e
    return CellBatch(n_max, labels, succ), enc


# ------------------------------------------------------------------------------- NB101 mutation
MUTATION_ATTEMPTS = 64


def apply_mutation_nb101(ops, succ, vertice, flip_mask, new_ops):
    """One BANANAS-style edit of a pruned NB101 cell (generate_test_graphs.py:437-456): XOR the
    upper-triangular adjacency with `flip_mask` (bit e = e-th (src<dst) pair in row-major order),
    replace interior ops by `new_ops` (dict node -> op code), prune.  Returns the packed child
    (labels, masks) or None when input and output are disconnected."""
    succ = list(succ[:vertice])
    ops = list(ops[:vertice])
    e = 0
    for src in range(vertice - 1):
        for dst in range(src + 1, vertice):
            if (flip_mask >> e) & 1:
                succ[src] ^= 1 << dst
            e += 1
    for v, o in new_ops.items():
        ops[v] = o
    out = vertice - 1
    fwd = 1
    for v in range(1, vertice):
        if any((fwd >> u) & 1 and (succ[u] >> v) & 1 for u in range(v)):
            fwd |= 1 << v
    bwd = 1 << out
    for v in range(out - 1, -1, -1):
        if succ[v] & bwd:
            bwd |= 1 << v
    keep = fwd & bwd
    if keep == 0:
        return None
    return _compact(ops, succ, keep, vertice)


def _mutate_nb101_cell(seed, idx, p_labels, p_succ):
    vertice = sum(1 for l in p_labels if l != NO_NODE)
    if vertice > 2:
        for attempt in range(MUTATION_ATTEMPTS):
            d = [draw(seed ^ 0x6D75746174650000, idx, attempt, t) for t in range(5)]
            flip = 0
            n_pairs = vertice * (vertice - 1) // 2
            for e in range(n_pairs):                       # each possible edge flips w.p. ~1/vertice
                if _below(d[e // 8], vertice, e % 8) == 0:
                    flip |= 1 << e
            new_ops = {}
            for ind in range(1, vertice - 1):              # each interior op mutates w.p. ~1/(vertice-2)
                if _below(d[3], vertice - 2, ind - 1) == 0:
                    avail = [o for o in (2, 3, 4) if o != p_labels[ind]]
                    new_ops[ind] = avail[_below(d[4], 2, ind - 1)]
            child = apply_mutation_nb101(p_labels, p_succ, vertice, flip, new_ops)
            if child is None:
                continue
            labels, masks = child
            n_nodes = sum(1 for l in labels if l != NO_NODE)
            n_edges = sum(bin(m).count("1") for m in masks)
            if n_nodes <= 3 or n_edges > 9:                # generate_test_graphs.py:486-492
                continue
            return labels, masks
    return _nb101_cell(seed ^ 0x66616C6C6261636B, idx)     # patience exhausted: a random cell (:529-535)


def mutate_cells_host(family: int, parents: CellBatch, seed: int, first_index: int, n_children: int) -> CellBatch:
    """Host twin of nbw_mutate_cells: child i is an edit of parent i % n_parents."""
    if family != 0:
        raise NotImplementedError("mutation needs the unpruned encoding for NB201 / DARTS cells")
    labels = np.full((n_children, 8), NO_NODE, dtype=np.uint8)
    succ = np.zeros((n_children, 8), dtype=np.uint16)
    npar = len(parents)
    for i in range(n_children):
        p = (first_index + i) % npar
        l, m = _mutate_nb101_cell(seed, first_index + i, [int(x) for x in parents.labels[p]],
                                  [int(x) for x in parents.succ[p]])
        labels[i] = l[:8]
        succ[i] = m[:8]
    return CellBatch(8, labels, succ)


_GEN = {0: _nb101_cell, 1: _nb201_cell, 2: _darts_cell}


def generate_cells_host(family: int, seed: int, first_index: int, n_cells: int) -> CellBatch:
    """Host restatement of nbw_generate_cells (bit-identical records)."""
    n_max = FAMILY_NMAX[family]
    labels = np.full((n_cells, n_max), NO_NODE, dtype=np.uint8)
    succ = np.zeros((n_cells, n_max), dtype=np.uint16)
    gen = _GEN[family]
    for i in range(n_cells):
        l, m = gen(seed, first_index + i)
        labels[i] = l[:n_max]
        succ[i] = m[:n_max]
    return CellBatch(n_max, labels, succ)


# --------------------------------------------------------------------------------------------- pool de-duplication
def isomorphism_classes_host(batch: CellBatch, labeled: bool = True, rounds: int = None):
    """Host twin of nbw_pool_dedup's partition: colour refinement over BOTH edge directions (a node's colour is
    refined by the multiset of its successors' colours and, separately, of its predecessors' colours), run to
    depth n_max; a cell's class is the multiset of its final node colours (+ node and edge counts).  Returns an
    int64 array of class ids numbered by first occurrence.

    This is the Weisfeiler-Lehman isomorphism test the reference's mutation() comments on
    (generate_test_graphs.py:497-520): `allow_isomorphism=False` drops a child that is isomorphic to its parent or
    to an earlier member of the pool.  labeled=False ignores the op names, like the reference's
    `iso.is_isomorphic(child.arch, prev_edit)` call (no node_match); labeled=True also requires equal ops."""
    N, n_max = batch.labels.shape
    rounds = n_max if rounds is None else rounds
    out = np.empty(N, dtype=np.int64)
    seen = {}
    for g in range(N):
        lab = batch.labels[g]
        nodes = [v for v in range(n_max) if lab[v] != NO_NODE]
        succ = {v: [j for j in range(n_max) if (int(batch.succ[g, v]) >> j) & 1] for v in nodes}
        pred = {v: [u for u in nodes if v in succ[u]] for v in nodes}
        # colours are hashes of (own colour, sorted successor colours, sorted predecessor colours): canonical across
        # cells (Python hashes tuples of ints deterministically), so the multiset of final colours is the cell's class
        sig = {v: (int(lab[v]) if labeled else 0) for v in nodes}
        for _ in range(rounds):
            sig = {v: hash((sig[v], tuple(sorted(sig[s] for s in succ[v])), tuple(sorted(sig[p] for p in pred[v]))))
                   for v in nodes}
        key = (len(nodes), sum(len(s) for s in succ.values()), tuple(sorted(sig[v] for v in nodes)))
        out[g] = seen.setdefault(key, len(seen))
    return out


def dedup_keep_first_host(batch: CellBatch, labeled: bool = True, n_protected: int = 0) -> np.ndarray:
    """keep[i] for the cells after the first n_protected: True iff no earlier cell (protected ones included)
    belongs to the same isomorphism class — the keep-first rule of mutation()'s evaluation pool."""
    cls = isomorphism_classes_host(batch, labeled)
    first = {}
    keep = np.zeros(len(cls), dtype=bool)
    for i, c in enumerate(cls):
        if c not in first:
            first[c] = i
            keep[i] = True
    return keep[n_protected:]
