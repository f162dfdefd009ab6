"""CPU-only tests: the C-ABI library loads and exports every symbol include/nbw.h declares (no
compute calls without a GPU), the packed cell format, the synthetic generators against the
reference's samplers, and the multi-rank host logic (gloo, world_size 2)."""
import os
import re
import socket
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT, load_golden
from nasbowl_b200 import _lib, synth
from nasbowl_b200.cells import NO_NODE, CellBatch, OpVocab, pack_networkx
from oracle import wl_oracle


# ---------------------------------------------------------------------------- C-ABI surface
def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "nbw.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nbw_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), name
    # ... and the ctypes prototypes cover the header exactly
    assert sorted(_lib.EXPORTED) == declared
    assert lib.nbw_version() >= 100
    assert lib.nbw_last_error(None) == b"null context"
    assert lib.nbw_launch_count(None) == 0


def test_model_struct_layout_matches_header():
    """ctypes mirror of nbw_model: field order from the header, natural alignment."""
    text = open(os.path.join(ROOT, "include", "nbw.h")).read()
    body = text[text.index("typedef struct {"):text.index("} nbw_model;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = re.findall(r"\b([a-zA-Z_][a-zA-Z0-9_]*)\s*(?:\[[A-Z_0-9]+\])?\s*[;,]", body)
    names = [n for n in names if n not in ("NBW_MAX_LEVELS",)]
    assert names == [f[0] for f in _lib.Model._fields_]
    assert _lib.Model.G.offset % 8 == 0 and _lib.Model.H.offset % 8 == 0 and _lib.Model.lik.offset % 8 == 0


def test_no_cpu_fallback_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("needs a CUDA-less host")
    from nasbowl_b200.engine import Engine
    from nasbowl_b200.kernels import WeisfilerLehman
    with pytest.raises(RuntimeError, match="no CPU path"):
        Engine.get()
    z, batch = load_golden("appendix_b")
    with pytest.raises(RuntimeError):
        WeisfilerLehman(h=1).fit_transform(batch)


# ---------------------------------------------------------------------------- cell records
def test_pack_roundtrip_and_node_renumbering():
    import networkx as nx
    g = nx.DiGraph()
    for n, op in [(0, "input"), (3, "conv"), (7, "output")]:      # non-contiguous ids (NB201/DARTS)
        g.add_node(n, op_name=op)
    g.add_edges_from([(0, 3), (3, 7), (0, 7)])
    vocab = OpVocab()
    b = pack_networkx([g], vocab)
    assert b.n_max == 8 and list(b.labels[0, :3]) == [0, 1, 2] and b.labels[0, 3] == NO_NODE
    assert list(b.succ[0, :3]) == [0b110, 0b100, 0]
    assert int(b.n_nodes[0]) == 3 and int(b.n_edges[0]) == 3
    rec = b.to_records()
    assert rec.shape == (1, 16)
    b2 = CellBatch.from_records(rec, 8)
    assert np.array_equal(b2.labels, b.labels) and np.array_equal(b2.succ, b.succ)
    w = b.widen(16)
    assert w.to_records().shape == (1, 48)
    w2 = CellBatch.from_records(w.to_records(), 16)
    assert np.array_equal(w2.labels[:, :8], b.labels) and np.array_equal(w2.succ[:, :8], b.succ)
    u = b.to_undirected()
    assert list(u.succ[0, :3]) == [0b110, 0b101, 0b011]
    g2 = b.to_networkx(vocab)[0]
    assert sorted(g2.edges()) == [(0, 1), (0, 2), (1, 2)]
    with pytest.raises(ValueError):
        pack_networkx([nx.path_graph(20, create_using=nx.DiGraph)], vocab, node_label="missing")
    g.nodes[3].pop("op_name")
    with pytest.raises(ValueError):
        pack_networkx([g], vocab)
    # widening a batch does not change its WL features
    z, batch = load_golden("nasbench101")
    assert np.array_equal(wl_oracle.gram_raw(batch, 2), wl_oracle.gram_raw(batch.widen(16), 2))


def test_vocab_pickles():
    import pickle
    v = OpVocab(["a", "b"])
    v2 = pickle.loads(pickle.dumps(v))
    assert v2.names == ["a", "b"] and v2.code("b") == 1 and v2.code("c") == 2


# ---------------------------------------------------------------------------- synthetic pools


def test_synthetic_generators_are_deterministic_and_well_formed():
    for fam in (0, 1, 2):
        a = synth.generate_cells_host(fam, 7, 100, 300)
        b = synth.generate_cells_host(fam, 7, 100, 300)
        c = synth.generate_cells_host(fam, 7, 250, 50)
        assert np.array_equal(a.to_records(), b.to_records())
        assert np.array_equal(a.to_records()[150:200], c.to_records())       # index-addressable
        k = a.n_nodes
        assert k.min() >= 2 and k.max() <= (15 if fam == 2 else 8 if fam == 1 else 7)
        # successor masks only reference existing nodes, no isolated nodes, first label is an input
        for i in range(len(a)):
            present = (1 << int(k[i])) - 1
            assert all(int(m) & ~present == 0 for m in a.succ[i])
            inc = 0
            for m in a.succ[i]:
                inc |= int(m)
            assert all(((int(a.succ[i, v]) != 0) or ((inc >> v) & 1)) for v in range(int(k[i])))
        assert len(synth.FAMILY_OPS[fam]) > int(a.labels[a.labels != NO_NODE].max())


def test_synthetic_construction_matches_reference_samplers():
    """Same choices in -> same cell out, for the reference's own construction code
    (generate_test_graphs.py prune / create_nasbench201_graph, darts/utils.py darts2graph)."""
    from oracle.ref_import import load_reference, reference_available
    if not reference_available():
        pytest.skip("reference tree not present on this machine")
    import random
    import networkx as nx
    R = load_reference()
    from darts.cnn.genotypes import Genotype
    from darts.utils import darts2graph, is_valid_darts
    rng = random.Random(5)

    def field(c, n):
        f = (c * 256 + n - 1) // n
        while (f * n) >> 8 < c:
            f += 1
        return f

    def run(gen, x, y):
        orig = synth.draw

        def fake(seed, idx, attempt, t):
            if attempt > 0:
                raise StopIteration
            return x if t == 0 else y
        synth.draw = fake
        try:
            return gen(0, 0)
        except StopIteration:
            return None
        finally:
            synth.draw = orig
    ops201 = ['nor_conv_3x3', 'nor_conv_1x1', 'avg_pool_3x3', 'skip_connect', 'none']
    for _ in range(400):
        # NB201
        choice = [rng.randrange(5) for _ in range(6)]
        x = sum(field(c, 5) << (8 * i) for i, c in enumerate(choice))
        g = R.gtg.create_nasbench201_graph([ops201[c] for c in choice])
        mine = run(synth._nb201_cell, x, 0)
        if len(g) == 0 or g.number_of_edges() == 0:
            assert mine is None
        else:
            rb = pack_networkx([g], OpVocab(synth.NB201_OPS))
            assert list(rb.labels[0]) == mine[0][:8] and list(rb.succ[0]) == mine[1][:8]
        # NB101
        x = rng.getrandbits(21)
        ops = [rng.randrange(3) for _ in range(5)]
        y = sum(field(c, 3) << (8 * i) for i, c in enumerate(ops))
        adj = np.zeros((7, 7), dtype=np.int8)
        for e, (r, c) in enumerate(synth._TRIU):
            adj[r, c] = (x >> e) & 1
        res = R.gtg.prune(adj, ['input'] + [synth.NB101_OPS[2 + o] for o in ops] + ['output'])
        mine = run(synth._nb101_cell, x, y)
        if res is None or not (0 < res[0].sum() <= 9):
            assert mine is None
        else:
            G = nx.from_numpy_array(res[0], create_using=nx.DiGraph)
            for i, n in enumerate(res[1]):
                G.nodes[i]['op_name'] = n
            rb = pack_networkx([G], OpVocab(synth.NB101_OPS))
            assert list(rb.labels[0]) == mine[0][:8] and list(rb.succ[0]) == mine[1][:8]
        # DARTS
        x = y = 0
        normal = []
        dops = synth.DARTS_OPS[4:]
        for i in range(4):
            a, b = rng.randrange(i + 2), rng.randrange(i + 1)
            y |= field(a, i + 2) << (16 * i) | field(b, i + 1) << (16 * i + 8)
            o0, o1 = rng.randrange(7), rng.randrange(7)
            x |= field(o0, 7) << (16 * i) | field(o1, 7) << (16 * i + 8)
            normal += [(dops[o0], a), (dops[o1], b + 1 if b >= a else b)]
        gt = Genotype(normal=normal, normal_concat=range(2, 6), reduce=normal, reduce_concat=range(2, 6))
        mine = run(synth._darts_cell, x, y)
        if not is_valid_darts(gt):
            assert mine is None
        else:
            rb = pack_networkx([darts2graph(gt)[0]], OpVocab(synth.DARTS_OPS), n_max=16)
            assert list(rb.labels[0]) == mine[0] and list(rb.succ[0]) == mine[1]


# ---------------------------------------------------------------------------- multi-rank host logic
def test_merge_topk_matches_reference_rule():
    from nasbowl_b200 import parallel
    from oracle import gp_oracle
    rng = np.random.RandomState(0)
    s = np.round(rng.randn(5000), 1).astype(np.float32)
    s[::97] = np.nan
    for world in (1, 2, 3, 8):
        vals, idx = [], []
        for r in range(world):
            lo, hi = parallel.shard_range(len(s), r, world)
            v, i = parallel.merge_topk(s[lo:hi], np.arange(lo, hi), 5, True)     # local distinct top-5
            vals.append(v)
            idx.append(i)
        v, i = parallel.merge_topk(np.concatenate(vals), np.concatenate(idx), 5, True)
        ok = ~np.isnan(s)
        ref = gp_oracle.top_n_distinct(np.where(ok, s, -np.inf), 5)
        assert list(i) == list(ref) and np.array_equal(v, s[ref])
    assert parallel.shard_range(10, 0, 3) == (0, 3) and parallel.shard_range(10, 2, 3) == (6, 10)


_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from nasbowl_b200 import parallel
from nasbowl_b200.pool_model import PoolModel
rank, world, _ = parallel.init_distributed("gloo")
# 1. pool-model broadcast
pm = None
if rank == 0:
    pm = PoolModel()
    pm.n_max, pm.h, pm.oa, pm.seed = 8, 2, False, 77
    pm.level0 = torch.arange(256, dtype=torch.int16)
    pm.tables = [None, torch.full((64,), 3, dtype=torch.uint8), torch.full((128,), 5, dtype=torch.uint8)]
    pm.masks, pm.col_off, pm.label_off = [0, 3, 7], [0, 128, 256, 384], [0, 5, 20, 60]
    pm.n_features, pm.n_labels = 384, 60
    pm.G = torch.arange(385 * 385, dtype=torch.float64).reshape(385, 385)
    pm.w_mu = torch.ones(385, dtype=torch.float64)
    pm.H = torch.full((61, 61), 2.5, dtype=torch.float64)
    pm.w_mu_prefix = torch.zeros(61, dtype=torch.float64)
    pm.lik, pm.y_mean, pm.y_std, pm.incumbent, pm.incumbent_idx = 0.01, 0.5, 2.0, 1.25, 9
    # packed-path tables and a second part of a sum (reads record slot 1, undirected, weighted levels)
    pm.id_hash = [None, torch.full((80,), 7, dtype=torch.uint8), None]
    pm.bucket_tags = [None, torch.full((256,), 1, dtype=torch.uint8), None]
    pm.bucket_entries = [None, torch.full((1024,), 2, dtype=torch.uint8), None]
    pm.bucket_masks = [0, 15, 0]
    pm.n_slots, pm.total_labels = 2, 100
    p2 = PoolModel()
    p2.n_max, p2.h, p2.seed, p2.slot, p2.undirected = 8, 1, 78, 1, True
    p2.level0 = torch.zeros(256, dtype=torch.int16)
    p2.tables, p2.masks = [None, torch.full((32,), 9, dtype=torch.uint8)], [0, 1]
    p2.col_off, p2.label_off, p2.n_features, p2.n_labels = [0, 128, 256], [0, 5, 40], 256, 40
    p2.level_weight, p2.label_base = [0.5, 0.25], 60
    p2.child = torch.arange(40, dtype=torch.int32)
    pm.more = [p2]
pm = parallel.broadcast_pool_model(pm, src=0, device="cpu")
assert len(pm.more) == 1 and pm.n_slots == 2 and pm.total_labels == 100 and pm.bucket_masks == [0, 15, 0]
assert pm.id_hash[0] is None and int(pm.id_hash[1].sum()) == 7 * 80 and int(pm.bucket_entries[1].sum()) == 2048
q = pm.more[0]
assert (q.slot, q.undirected, q.level_weight, q.label_base, q.h) == (1, True, [0.5, 0.25], 60, 1)
assert q.child.dtype == torch.int32 and q.child.tolist() == list(range(40)) and int(q.tables[1].sum()) == 9 * 32
assert pm.h == 2 and pm.seed == 77 and pm.masks == [0, 3, 7] and pm.incumbent == 1.25
assert pm.tables[0] is None and int(pm.tables[2].sum()) == 5 * 128 and float(pm.G[384, 384]) == 385 * 385 - 1
assert float(pm.H.sum()) == 2.5 * 61 * 61 and pm.level0[255] == 255
# 2. sharded top-k agrees on every rank with the single-rank answer
rng = np.random.RandomState(3)
s = np.round(rng.randn(4001), 1).astype(np.float32)
lo, hi = parallel.shard_range(len(s), rank, world)
v, i = parallel.merge_topk(s[lo:hi], np.arange(lo, hi), 5, True)
gv, gi = parallel.allgather_topk(v, i, 5, True, device="cpu")
rv, ri = parallel.merge_topk(s, np.arange(len(s)), 5, True)
assert list(gi) == list(ri) and np.array_equal(gv, rv), (gi, ri)
dist.barrier()
print("rank", rank, "ok")
"""


def test_two_rank_gloo_broadcast_and_topk(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    script = tmp_path / "worker.py"
    script.write_text(_WORKER % ROOT)
    procs = []
    for rank in range(2):
        env = dict(os.environ, RANK=str(rank), WORLD_SIZE="2", LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                   MASTER_PORT=str(port), CUDA_VISIBLE_DEVICES="")
        procs.append(subprocess.Popen([sys.executable, str(script)], env=env, stdout=subprocess.PIPE,
                                      stderr=subprocess.STDOUT, text=True))
    outs = [p.communicate(timeout=240)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o
    assert "rank 0 ok" in outs[0] and "rank 1 ok" in outs[1]


def test_mutation_pool_host_twin():
    """BANANAS-style NB101 mutation: children are valid cells, the edit+prune step agrees with
    the reference's `prune` (generate_test_graphs.py:26-78,437-456) for arbitrary edits."""
    import random
    par = synth.generate_cells_host(0, 3, 0, 12)
    ch = synth.mutate_cells_host(0, par, 7, 100, 400)
    again = synth.mutate_cells_host(0, par, 7, 100, 400)
    assert np.array_equal(ch.to_records(), again.to_records())
    assert ch.n_edges.max() <= 9 and ch.n_edges.min() >= 1 and ch.n_nodes.max() <= 7
    assert (ch.n_nodes >= 4).mean() > 0.9                     # the rest are random fall-back cells
    assert not np.array_equal(ch.to_records()[:12], par.to_records())
    with pytest.raises(NotImplementedError):       # NB201 / DARTS mutate on encodings (mutate_encoded_host)
        synth.mutate_cells_host(1, par, 7, 0, 4)
    from oracle.ref_import import load_reference, reference_available
    if not reference_available():
        return
    import networkx as nx
    R = load_reference()
    rng = random.Random(0)
    for _ in range(300):
        p = rng.randrange(len(par))
        labels, succ = [int(x) for x in par.labels[p]], [int(x) for x in par.succ[p]]
        v = sum(1 for l in labels if l != NO_NODE)
        if v <= 2:
            continue
        pairs = v * (v - 1) // 2
        flip = rng.getrandbits(pairs) & rng.getrandbits(pairs)
        new_ops = {i: rng.choice([o for o in (2, 3, 4) if o != labels[i]]) for i in range(1, v - 1) if rng.random() < 0.4}
        mine = synth.apply_mutation_nb101(labels, succ, v, flip, new_ops)
        adj = np.array([[(succ[s] >> d) & 1 for d in range(v)] for s in range(v)], dtype=np.int8)
        e = 0
        for s in range(v - 1):
            for d in range(s + 1, v):
                if (flip >> e) & 1:
                    adj[s, d] = 1 - adj[s, d]
                e += 1
        ops = [synth.NB101_OPS[new_ops.get(i, labels[i])] for i in range(v)]
        res = R.gtg.prune(adj, ops)
        if res is None:
            assert mine is None
        else:
            G = nx.from_numpy_array(res[0], create_using=nx.DiGraph)
            for i, nm in enumerate(res[1]):
                G.nodes[i]['op_name'] = nm
            rb = pack_networkx([G], OpVocab(synth.NB101_OPS))
            assert mine is not None and list(rb.labels[0]) == mine[0][:8] and list(rb.succ[0]) == mine[1][:8]


def test_encoded_mutation_host_twin_and_reference_rules():
    """NB201 / DARTS mutation on encodings: deterministic, shard-invariant, children valid; the
    encoding -> cell construction agrees with the reference's own constructors for MUTATED
    architectures too (DARTS wiring edits may point at later blocks); the edit statistics agree with
    the reference's _banana_mutate / _mutate."""
    import random
    for family, edits in ((1, 1), (2, 1), (2, 3)):
        _, par = synth.generate_encoded_host(family, 3, 0, 10)
        ch, enc = synth.mutate_encoded_host(family, par, 11, 0, 300, edits)
        a, ea = synth.mutate_encoded_host(family, par, 11, 0, 120, edits)
        b, eb = synth.mutate_encoded_host(family, par, 11, 120, 180, edits)
        assert np.array_equal(np.concatenate([a.to_records(), b.to_records()]), ch.to_records())
        assert np.array_equal(np.concatenate([ea, eb]), enc)
        assert ch.n_nodes.min() >= 2 and ch.n_edges.min() >= 1
        diffs = np.array([(enc[i] != par[i % 10]).sum() for i in range(300)])
        if family == 1:
            assert 0.9 < diffs.mean() < 1.5                      # six ops x ~1/5
            assert enc[:, :6].max() <= 4 and not enc[:, 6:].any()
        else:
            assert diffs.max() <= edits and diffs.mean() > 0.6 * edits
            assert enc[:, 0::2].max() <= 6 and enc[:, 1::2].max() <= 5
            assert all(0 in e[1::2] and 1 in e[1::2] for e in enc)                      # is_valid_darts
            assert all(e[1] != e[3] and e[5] != e[7] and e[9] != e[11] and e[13] != e[15] for e in enc)
        # children rebuilt from their encodings are the cells that were returned
        for i in range(0, 300, 7):
            e = [int(v) for v in enc[i]]
            cell = synth.nb201_from_choice(e[:6]) if family == 1 else synth.darts_from_genotype(e[0::2], e[1::2])
            n_max = ch.n_max
            assert cell[0][:n_max] == list(ch.labels[i]) and cell[1][:n_max] == list(ch.succ[i])
    # encodings <-> the architecture descriptions the reference's drivers evaluate
    _, e201 = synth.generate_encoded_host(1, 3, 0, 20)
    for e in e201:
        assert synth.nb201_encoding(synth.nb201_arch_string(e)) == [int(v) for v in e]
    _, ed = synth.generate_encoded_host(2, 3, 0, 20)
    for e in ed:
        assert synth.darts_encoding(synth.darts_genotype_pairs(e)) == [int(v) for v in e]
    from oracle.ref_import import load_reference, reference_available
    if not reference_available():
        return
    R = load_reference()
    for e in e201[:5]:      # the string is the reference's own G.name for that architecture
        g = R.gtg.create_nasbench201_graph([synth.NB201_CHOICES[int(c)] for c in e[:6]])
        assert g.name == synth.nb201_arch_string(e)
    from darts.cnn.genotypes import Genotype
    from darts.utils import darts2graph
    ops201 = ['nor_conv_3x3', 'nor_conv_1x1', 'avg_pool_3x3', 'skip_connect', 'none']
    dops = synth.DARTS_OPS[4:]
    _, par = synth.generate_encoded_host(2, 5, 0, 20)
    _, enc = synth.mutate_encoded_host(2, par, 13, 0, 200, 4)
    forward_refs = 0
    for e in enc:
        ops8, ins8 = [int(v) for v in e[0::2]], [int(v) for v in e[1::2]]
        forward_refs += any(ins8[k] >= k // 2 + 2 for k in range(8))
        normal = [(dops[o], i) for o, i in zip(ops8, ins8)]
        gt = Genotype(normal=normal, normal_concat=range(2, 6), reduce=normal, reduce_concat=range(2, 6))
        rb = pack_networkx([darts2graph(gt)[0]], OpVocab(synth.DARTS_OPS), n_max=16)
        mine = synth.darts_from_genotype(ops8, ins8)
        assert list(rb.labels[0]) == mine[0] and list(rb.succ[0]) == mine[1]
    assert forward_refs > 10                                    # the case random sampling never produces
    # edit statistics against the reference's own mutation code
    import numpy.random as npr
    import networkx as nx
    random.seed(0)
    npr.seed(0)
    from oracle.ref_import import load_reference_functions
    import networkx
    # the module itself imports the NAS-Bench APIs and the CNN trainer: run just its two functions
    _mutate = load_reference_functions("darts/darts_objectives.py", ["get_ops", "_mutate"],
                                       dict(np=np, nx=networkx, Genotype=Genotype, darts2graph=darts2graph))["_mutate"]
    gt0 = Genotype(normal=[(dops[int(o)], int(i)) for o, i in zip(par[0][0::2], par[0][1::2])], normal_concat=range(2, 6),
                   reduce=[], reduce_concat=range(2, 6))
    ref_same = ref_op = 0
    N = 1500
    for _ in range(N):
        _, g = _mutate(gt0, 'darts', 1)
        changed = [k for k in range(8) if g.normal[k] != gt0.normal[k]]
        ref_same += not changed
        ref_op += bool(changed) and g.normal[changed[0]][1] == gt0.normal[changed[0]][1]
    _, enc1 = synth.mutate_encoded_host(2, par[:1], 17, 0, N, 1)
    mine_same = sum(bool((e == par[0]).all()) for e in enc1)
    mine_op = sum(bool((e != par[0]).any() and (e[1::2] == par[0][1::2]).all()) for e in enc1)
    # (children that leave an input unused are redrawn on our side, so allow a wide band)
    assert abs(ref_same - mine_same) < 0.05 * N and abs(ref_op - mine_op) < 0.08 * N
    # NB201: number of changed ops per child, reference _banana_mutate vs ours
    g0 = R.gtg.create_nasbench201_graph([ops201[int(c)] for c in synth.generate_encoded_host(1, 5, 0, 1)[1][0][:6]])
    _, p201 = synth.generate_encoded_host(1, 5, 0, 1)
    ref_counts = np.zeros(7)
    for _ in range(N):
        pool, _ = R.gtg.mutation([g0], [0.0], n_best=1, n_mutate=1, pool_size=1, allow_isomorphism=True,
                                 benchmark='nasbench201')
        child_ops = [s[:-2] for s in pool[0].name.split("|") if "~" in s]
        ref_counts[sum(ops201.index(o) != int(c) for o, c in zip(child_ops, p201[0][:6]))] += 1
    _, e201 = synth.mutate_encoded_host(1, p201, 19, 0, N)
    mine_counts = np.bincount([(e[:6] != p201[0][:6]).sum() for e in e201], minlength=7)
    assert np.abs(ref_counts - mine_counts).max() < 0.06 * N


# ---------------------------------------------------------------------------- oracle invariants
def test_oracle_invariants_on_random_cells():
    """Properties any correct WL subtree kernel has; they guard the oracle itself (hypothesis)."""
    from hypothesis import given, settings, strategies as st_

    @settings(max_examples=25, deadline=None)
    @given(st_.integers(0, 2 ** 31 - 1), st_.sampled_from([0, 1, 2]), st_.booleans())
    def prop(seed, fam, oa):
        batch = synth.generate_cells_host(fam, seed, seed % 1000, 12)
        K = wl_oracle.gram_raw_levels(batch, 3, oa)
        assert (K >= 0).all() and np.array_equal(K, K.transpose(0, 2, 1))
        Ksum = K.sum(0)
        d = np.diag(Ksum)
        assert (Ksum.astype(np.float64) ** 2 <= np.outer(d, d) + 1e-9).all()          # Cauchy-Schwarz
        if oa:
            assert (Ksum <= wl_oracle.gram_raw(batch, 3, False)).all()                 # min-sum <= dot product
            assert np.array_equal(d, 4 * batch.n_nodes)                                # (h+1) * n_nodes, SURVEY Q16
        # renumbering the nodes of every cell leaves the kernel unchanged
        rng = np.random.RandomState(seed % 9973)
        lab = np.full_like(batch.labels, NO_NODE)
        suc = np.zeros_like(batch.succ)
        for i in range(len(batch)):
            k = int(batch.n_nodes[i])
            perm = rng.permutation(k)                  # old node v -> new index perm[v]
            for v in range(k):
                lab[i, perm[v]] = batch.labels[i, v]
                m = 0
                for j in range(k):
                    if (int(batch.succ[i, v]) >> j) & 1:
                        m |= 1 << int(perm[j])
                suc[i, perm[v]] = m
        assert np.array_equal(wl_oracle.gram_raw(CellBatch(batch.n_max, lab, suc), 3, oa), Ksum)
        # the two restatements of the relabelling agree
        assert np.array_equal(wl_oracle.gram_raw(batch, 3, oa, fast=False), Ksum)
    prop()


def test_weight_gradient_closed_form_matches_finite_differences():
    """The closed form the oracle (and nbw_kernel_weight_grad) uses for d nlml'/d w_i of
    K = normalise(sum_i w_i R_i) + lik I, against central differences of the reference's
    marginal-likelihood formula (utils.py:102-140, incl. the sign of its log-det term)."""
    from oracle import gp_oracle
    rng = np.random.RandomState(3)
    n, m = 25, 3
    raws = []
    for _ in range(m):
        F = rng.randint(0, 4, size=(n, 12)).astype(np.float64)
        raws.append(F @ F.T + np.eye(n))                      # integer Grams with a positive diagonal
    y = rng.randn(n)
    w = np.array([0.2, 0.5, 0.3])
    lik = 0.01

    def nlml(wv):
        K = gp_oracle.normalize_gram(sum(wv[i] * raws[i] for i in range(m)))
        K_i, ld, _ = gp_oracle.pd_inverse(K, lik)
        return gp_oracle.nlml_prime(K_i, ld, y)
    S = sum(w[i] * raws[i] for i in range(m))
    K_i, _, _ = gp_oracle.pd_inverse(gp_oracle.normalize_gram(S), lik)
    g = gp_oracle.weight_gradient(K_i, K_i @ y, S, raws, n)
    for i in range(m):
        e = np.zeros(m)
        e[i] = 1e-6
        fd = (nlml(w + e) - nlml(w - e)) / 2e-6
        assert abs(fd - g[i]) < 1e-6 * max(1.0, abs(fd)), (i, fd, g[i])


def test_stable_suffix_detection_and_child_table():
    """PoolModel._find_stable_suffix: the first level l >= 3 with D_{l-1} == D_{l-2}, only if every deeper
    level keeps that size; child[parent] = label for the levels on the fast path."""
    from types import SimpleNamespace
    from nasbowl_b200.pool_model import PoolModel

    def fake(dims, perms):
        off = [0]
        for d in dims:
            off.append(off[-1] + d)
        parents = torch.zeros(off[-1], dtype=torch.int32)
        for l, p in perms.items():
            parents[off[l]:off[l + 1]] = torch.tensor(p, dtype=torch.int32)
        return SimpleNamespace(dims=dims, label_off=off, parents=parents)
    # levels 2, 3, 4 have four labels each: level 4 is the first with D_3 == D_2
    fit = fake([3, 3, 4, 4, 4], {3: [2, 0, 3, 1], 4: [1, 3, 0, 2]})
    pm = PoolModel()
    pm._find_stable_suffix(fit, 4)
    assert pm.first_stable == 4
    child = pm.child.tolist()
    off = fit.label_off
    for z, par in enumerate([1, 3, 0, 2]):                    # label z of level 4 has parent `par` at level 3
        assert child[off[3] + par] == z
    # final_label: a stable label -> global stacked index of its descendant at the deepest level
    assert pm.final_label.tolist()[off[4]:off[5]] == [off[4] + z for z in range(4)]
    pm.set_label_base(100)
    assert pm.final_label.tolist()[off[4]:off[5]] == [100 + off[4] + z for z in range(4)]
    fit5 = fake([3, 3, 4, 4, 4, 4], {3: [2, 0, 3, 1], 4: [1, 3, 0, 2], 5: [3, 2, 1, 0]})
    pm5 = PoolModel()
    pm5._find_stable_suffix(fit5, 5)
    o5 = fit5.label_off
    fl = pm5.final_label.tolist()
    for z4 in range(4):                                        # level-4 label z4 -> its child at level 5
        kid = [z for z, par in enumerate([3, 2, 1, 0]) if par == z4][0]
        assert fl[o5[4] + z4] == o5[5] + kid
    # depth 3 only: level 3 needs D_2 == D_1 (4 vs 4 here -> stable from 3)
    pm = PoolModel()
    pm._find_stable_suffix(fake([3, 4, 4, 4], {3: [2, 0, 3, 1]}), 3)
    assert pm.first_stable == 3 and pm.child.tolist()[3 + 4 + 2] == 0        # D_2 == D_1: level 3 already stable
    # still refining, or too shallow: off
    for dims in ([3, 5, 9, 12, 14], [3, 4, 4], [3, 4, 6, 6]):
        pm = PoolModel()
        pm._find_stable_suffix(fake(dims, {}), len(dims) - 1)
        assert pm.first_stable == 0 and pm.child is None
    os.environ["NBW_NO_STABLE"] = "1"
    try:
        pm = PoolModel()
        pm._find_stable_suffix(fit, 4)
        assert pm.first_stable == 0
    finally:
        del os.environ["NBW_NO_STABLE"]


def test_dmu_dphi_closed_form_matches_finite_differences():
    """gp_oracle.dmu_dphi (the closed form GraphGP.dmu_dphi evaluates) against central differences of
    mu(phi*) = sum_i alpha_i (phi* . x_i) / (|phi*| |x_i|)."""
    from oracle import gp_oracle
    rng = np.random.RandomState(5)
    n, D = 12, 9
    X = rng.randint(0, 4, size=(n, D)).astype(np.float64) + np.eye(n, D)
    Y = rng.randint(0, 3, size=(4, D)).astype(np.float64) + 1.0
    A = rng.randn(n, n)
    st = gp_oracle.GPState(h=0, oa=False, lik=0.01, y_mean=0.0, y_std=1.0, y_norm=rng.randn(n), y_raw=None,
                           K=None, K_i=A @ A.T / n + np.eye(n), logDetK=0.0, L=None, diag_raw=None)
    alpha = st.K_i @ st.y_norm
    nx_ = np.sqrt((X * X).sum(1))

    def mu(yv):
        return float(alpha @ ((X @ yv) / (np.sqrt(yv @ yv) * nx_)))
    g = gp_oracle.dmu_dphi(st, X, Y)
    for j in range(Y.shape[0]):
        for c in range(D):
            e = np.zeros(D)
            e[c] = 1e-6
            fd = (mu(Y[j] + e) - mu(Y[j] - e)) / 2e-6
            assert abs(fd - g[j, c]) < 1e-7, (j, c, fd, g[j, c])


def test_cpulist_parser_and_numa_binding_is_best_effort():
    from nasbowl_b200 import parallel
    assert parallel.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert parallel.parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    assert parallel.bind_to_gpu_numa_node(0) is None        # no GPU here: nothing read, nothing changed
    assert os.sched_getaffinity(0) == before


def _decode_wire(rec, pal, n_max):
    """Decoder of nbw_model.compact exactly as include/nbw.h states it (what the kernel prologues implement)."""
    import numpy as np
    from nasbowl_b200.cells import NO_NODE
    N = rec.shape[0]
    labels = np.full((N, n_max), NO_NODE, dtype=np.uint8)
    succ = np.zeros((N, n_max), dtype=np.uint16)
    words = rec.view(np.uint64).reshape(N, -1)
    for i in range(N):
        if n_max == 8:
            w = int(words[i, 0])
            idx = [(w >> (3 * v)) & 7 for v in range(8)]
            none, tri, base = 7, w >> 24, lambda u: u * (15 - u) // 2
        else:
            w0 = int(words[i, 0])
            idx = [(w0 >> (4 * v)) & 15 for v in range(16)]
            none, tri, base = 15, int(words[i, 1]) | (int(words[i, 2]) << 64), lambda u: u * (31 - u) // 2
        for v in range(n_max):
            if idx[v] != none:
                labels[i, v] = pal[idx[v]]
            for t in range(v + 1, n_max):
                if (tri >> (base(v) + (t - v - 1))) & 1:
                    succ[i, v] |= 1 << t
    return labels, succ


def test_wire_records_round_trip_on_the_host():
    """CellBatch.to_compact (8 bytes per 8-node cell, 24 per 16-node cell) against the layout written in nbw.h."""
    import numpy as np
    from nasbowl_b200 import synth
    for fam, n_max, width in ((0, 8, 8), (1, 8, 8), (2, 16, 24)):
        batch = synth.generate_cells_host(fam, 3, 0, 300)
        rec, pal = batch.to_compact()
        assert rec.shape == (300, width) and rec.dtype == np.uint8
        labels, succ = _decode_wire(rec, pal, n_max)
        assert np.array_equal(labels, batch.labels) and np.array_equal(succ, batch.succ)
        # a caller-supplied palette (the model's) is honoured; an op outside it is refused
        rec2, pal2 = batch.to_compact(palette=pal)
        labels2, succ2 = _decode_wire(rec2, pal2, n_max)        # (unused palette entries may alias an op: still decodes)
        assert np.array_equal(pal2, pal) and np.array_equal(labels2, batch.labels) and np.array_equal(succ2, batch.succ)
        short = pal.copy()
        present_codes = np.unique(batch.labels[batch.labels != 255])
        short[:] = present_codes[0]
        if len(present_codes) > 1:
            with pytest.raises(ValueError):
                batch.to_compact(palette=short)


def test_gram_row_blocks_are_dealt_by_work():
    """parallel.deal_row_blocks: every block exactly once, deterministic, and — for the symmetric schedule, where a
    block's work is rows x (N - first row) — balanced to within one block of the mean."""
    from nasbowl_b200 import parallel
    N, B = 200_000, 4096
    for world in (1, 2, 3, 8):
        for symmetric in (True, False):
            deal = parallel.deal_row_blocks(N, B, world, symmetric)
            assert len(deal) == world and deal == parallel.deal_row_blocks(N, B, world, symmetric)
            flat = sorted(r0 for m in deal for r0 in m)
            assert flat == list(range(0, N, B))
            if symmetric and world > 1:
                work = [sum(min(B, N - r0) * (N - r0) for r0 in m) for m in deal]
                assert max(work) - min(work) <= B * N
                assert max(work) <= 1.02 * sum(work) / world


def test_training_set_is_packed_incrementally():
    """GraphGP._pack_train: a refit after reset_XY(x + new_x, ...) walks only the new networkx objects (the BO loops
    of the reference append a batch per iteration), an edited example invalidates the cache from its position on,
    and the records always equal those of a full pack."""
    import pickle
    from nasbowl_b200.bayesopt import GraphGP
    from nasbowl_b200.cells import CellSlots
    from nasbowl_b200.kernels import WeisfilerLehman
    graphs = synth.generate_cells_host(1, 5, 0, 60).to_networkx(OpVocab(synth.FAMILY_OPS[1]))
    wide = synth.generate_cells_host(2, 5, 0, 3).to_networkx(OpVocab(synth.FAMILY_OPS[2]))   # DARTS-shaped: 16-node records
    y = torch.randn(70)
    gp = GraphGP(graphs[:40], y[:40], [WeisfilerLehman(h=2)])
    walked = []
    pack = gp._pack
    gp._pack = lambda gr: (walked.append(len(list(gr))), pack(gr))[1]

    def full(x):
        return pack_networkx(x, OpVocab(list(gp.kernels[0].vocab.names))).to_records()

    s = gp._pack_train()
    assert walked == [40] and np.array_equal(s.to_records(), full(graphs[:40]))
    gp._pack_train()
    assert walked == [40]                                                      # nothing new: nothing walked
    gp.reset_XY(graphs[:45], y[:45])
    s = gp._pack_train()
    assert walked == [40, 5] and len(s) == 45 and np.array_equal(s.to_records(), full(graphs[:45]))
    # wider cells in the tail widen the whole batch
    x = graphs[:45] + wide
    gp.reset_XY(x, y[:48])
    s = gp._pack_train()
    assert walked == [40, 5, 3] and s.n_max == 16 and np.array_equal(s.to_records(), full(x))
    # a shorter list that is a prefix of the cached one
    gp.reset_XY(graphs[:30], y[:30])
    s = gp._pack_train()
    assert walked == [40, 5, 3] and np.array_equal(s.to_records(), full(graphs[:30]))
    # an example edited in place (a node removed) is re-walked together with everything after it
    g = graphs[10]
    g.remove_node(max(g.nodes))
    s = gp._pack_train()
    assert walked == [40, 5, 3, 20] and np.array_equal(s.to_records(), full(graphs[:30]))
    # a different object at the same position
    x = graphs[:29] + [graphs[50]]
    gp.reset_XY(x, y[:30])
    s = gp._pack_train()
    assert walked[-1] == 1 and np.array_equal(s.to_records(), full(x))
    # switched off: everything is walked; packed inputs pass through
    gp.pack_cache = False
    gp._pack_train()
    assert walked[-1] == 30
    gp.pack_cache = True
    gp._pack = pack
    assert pickle.loads(pickle.dumps(gp))._pack_memo is None
    # examples made of two cells (record slots)
    pairs = [(a, b) for a, b in zip(graphs[:20], graphs[20:40])]
    gp2 = GraphGP(pairs[:15], y[:15], [WeisfilerLehman(h=2, slot=0), WeisfilerLehman(h=2, slot=1)])
    a = gp2._pack_train()
    gp2.reset_XY(pairs, y[:20])
    b = gp2._pack_train()
    assert isinstance(b, CellSlots) and b.n_slots == 2 and len(b) == 20
    assert np.array_equal(b.to_records()[:15], a.to_records())
    ref = CellSlots([pack_networkx([p[i] for p in pairs], OpVocab(list(gp2.kernels[0].vocab.names))) for i in range(2)])
    assert np.array_equal(b.to_records(), ref.to_records())
