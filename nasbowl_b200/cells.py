"""Packed cell batches: the HBM data layout of the WL hot path.

A *cell* is a small labelled DAG (NAS-Bench-101: <=7 nodes, NAS-Bench-201: <=8, DARTS: <=15;
reference generators bayesopt/generate_test_graphs.py:543-641, darts/utils.py:11-98).  The
reference hands cells around as ``networkx.DiGraph`` objects with an ``op_name`` node
attribute and converts them one at a time into GraKeL's ``[{(u,v):1}, {node:label}, None]``
triple (kernels/weisfilerlehman.py:134-136).  Here a batch of N cells is ONE contiguous array
of fixed-stride records, so a sub-warp reads a whole cell with one coalesced vector load:

    n_max = 8  : 16-byte record  = labels[8]  u8 | succ[8]  u8   (one 128-bit load)
    n_max = 16 : 48-byte record  = labels[16] u8 | succ[16] u16  (three 128-bit loads)

``labels[v]`` is the op code of node v (0xFF = no node); ``succ[v]`` is the bitmask of v's
successors (bit j set <=> edge v->j): a bit-packed CSR row, the neighbourhood the reference's
WL aggregates over (weisfeiler_lehman.py:266-267, Assumption A1 of SURVEY.md).  Nodes are
renumbered 0..k-1 in ascending order of their original ids (NB201/DARTS cells delete nodes
without renumbering: generate_test_graphs.py:155,168).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, List, Optional, Sequence

import numpy as np

NO_NODE = 0xFF
MAX_OPS = 255


class OpVocab:
    """Append-only map op-name -> u8 op code shared by every batch fed to one kernel object."""

    def __init__(self, names: Optional[Sequence[str]] = None):
        self.names: List[str] = []
        self._code = {}
        for n in names or ():
            self.code(n)

    def code(self, name) -> int:
        c = self._code.get(name)
        if c is None:
            c = len(self.names)
            if c >= MAX_OPS:
                raise ValueError("more than %d distinct op names" % MAX_OPS)
            self._code[name] = c
            self.names.append(name)
        return c

    def __len__(self):
        return len(self.names)

    def __getstate__(self):
        return {"names": list(self.names)}

    def __setstate__(self, st):
        self.__init__(st["names"])


def record_stride(n_max: int) -> int:
    if n_max == 8:
        return 16
    if n_max == 16:
        return 48
    raise ValueError("n_max must be 8 or 16, got %r" % (n_max,))


@dataclass
class CellBatch:
    """N packed cells. ``labels``: (N, n_max) u8, ``succ``: (N, n_max) u16 (host-side view)."""
    n_max: int
    labels: np.ndarray
    succ: np.ndarray

    def __post_init__(self):
        record_stride(self.n_max)
        self.labels = np.ascontiguousarray(self.labels, dtype=np.uint8)
        self.succ = np.ascontiguousarray(self.succ, dtype=np.uint16)
        assert self.labels.shape == self.succ.shape and self.labels.shape[1] == self.n_max

    def __len__(self):
        return self.labels.shape[0]

    @property
    def n_nodes(self) -> np.ndarray:
        return (self.labels != NO_NODE).sum(axis=1)

    @property
    def n_edges(self) -> np.ndarray:
        s = self.succ.astype(np.uint32)
        cnt = np.zeros(s.shape, dtype=np.uint32)
        for j in range(self.n_max):
            cnt += (s >> j) & 1
        return cnt.sum(axis=1)

    def to_records(self) -> np.ndarray:
        """Device layout: (N, stride) u8."""
        n = len(self)
        rec = np.empty((n, record_stride(self.n_max)), dtype=np.uint8)
        rec[:, :self.n_max] = self.labels
        if self.n_max == 8:
            rec[:, 8:16] = self.succ.astype(np.uint8)
        else:
            rec[:, 16:48] = self.succ.astype("<u2").view(np.uint8).reshape(n, 32)
        return rec

    @staticmethod
    def from_records(rec: np.ndarray, n_max: int) -> "CellBatch":
        rec = np.ascontiguousarray(rec, dtype=np.uint8).reshape(-1, record_stride(n_max))
        labels = rec[:, :n_max].copy()
        if n_max == 8:
            succ = rec[:, 8:16].astype(np.uint16)
        else:
            succ = rec[:, 16:48].copy().view("<u2").astype(np.uint16)
        return CellBatch(n_max, labels, succ)

    def __getitem__(self, idx) -> "CellBatch":
        if isinstance(idx, (int, np.integer)):
            idx = slice(idx, idx + 1)
        return CellBatch(self.n_max, self.labels[idx], self.succ[idx])

    @staticmethod
    def concat(batches: Iterable["CellBatch"]) -> "CellBatch":
        batches = list(batches)
        n_max = max(b.n_max for b in batches)
        batches = [b.widen(n_max) for b in batches]
        return CellBatch(n_max, np.concatenate([b.labels for b in batches]),
                         np.concatenate([b.succ for b in batches]))

    def to_compact(self, palette: Optional[np.ndarray] = None):
        """(records uint8 [N, 8], palette uint8 [8]): the 8-byte wire form (nbw_model.compact).  Needs 8-node
        records, at most seven distinct op codes and topologically numbered nodes (every edge u -> v has u < v);
        raises ValueError otherwise (such pools travel as full records).  16-node batches give the 24-byte form
        ([N, 24], palette uint8 [16], at most fifteen distinct op codes)."""
        if self.n_max == 16:
            return self._to_compact16(palette)
        present = self.labels != NO_NODE
        if palette is None:
            codes = np.unique(self.labels[present])
            if len(codes) > 7:
                raise ValueError("compact records hold at most 7 distinct op codes (got %d)" % len(codes))
            palette = np.zeros(8, dtype=np.uint8)
            palette[:len(codes)] = codes
            n_codes = len(codes)
        else:
            palette = np.asarray(palette, dtype=np.uint8)
            n_codes = 7
        lut = np.full(256, 255, dtype=np.int64)
        lut[palette[:n_codes]] = np.arange(n_codes)
        idx = np.where(present, lut[self.labels], 7)
        if (idx == 255).any():
            raise ValueError("an op code is not in the palette")
        w = np.zeros(len(self), dtype=np.uint64)
        for v in range(8):
            w |= idx[:, v].astype(np.uint64) << np.uint64(3 * v)
        s = self.succ.astype(np.uint64)
        for u in range(8):
            if (s[:, u] & np.uint64((1 << (u + 1)) - 1)).any():
                raise ValueError("compact records need topologically numbered nodes (an edge u -> v with v <= u)")
        off = 24
        for u in range(7):
            field = (s[:, u] >> np.uint64(u + 1)) & np.uint64((1 << (7 - u)) - 1)
            w |= field << np.uint64(off)
            off += 7 - u
        return w.view(np.uint8).reshape(len(self), 8).copy(), palette

    def _to_compact16(self, palette: Optional[np.ndarray] = None):
        """(records uint8 [N, 24], palette uint8 [16]): the 24-byte wire form of 16-node cells (nbw_model.compact):
        word 0 = sixteen 4-bit palette indices (15 = no node), words 1-2 = the 120 bits of the strict upper triangle."""
        present = self.labels != NO_NODE
        if palette is None:
            codes = np.unique(self.labels[present])
            if len(codes) > 15:
                raise ValueError("24-byte compact records hold at most 15 distinct op codes (got %d)" % len(codes))
            palette = np.zeros(16, dtype=np.uint8)
            palette[:len(codes)] = codes
            n_codes = len(codes)
        else:
            palette = np.asarray(palette, dtype=np.uint8)
            n_codes = 15
        lut = np.full(256, 255, dtype=np.int64)
        lut[palette[:n_codes]] = np.arange(n_codes)
        idx = np.where(present, lut[self.labels], 15)
        if (idx == 255).any():
            raise ValueError("an op code is not in the palette")
        w0 = np.zeros(len(self), dtype=np.uint64)
        for v in range(16):
            w0 |= idx[:, v].astype(np.uint64) << np.uint64(4 * v)
        s = self.succ.astype(np.uint64)
        lo = np.zeros(len(self), dtype=np.uint64)
        hi = np.zeros(len(self), dtype=np.uint64)
        off = 0
        for u in range(15):
            if (s[:, u] & np.uint64((1 << (u + 1)) - 1)).any():
                raise ValueError("compact records need topologically numbered nodes (an edge u -> v with v <= u)")
            n_bits = 15 - u
            field = (s[:, u] >> np.uint64(u + 1)) & np.uint64((1 << n_bits) - 1)
            if off < 64:
                lo |= field << np.uint64(off)          # (bits beyond 63 fall off: they go to `hi` below)
                if off + n_bits > 64:
                    hi |= field >> np.uint64(64 - off)
            else:
                hi |= field << np.uint64(off - 64)
            off += n_bits
        if s[:, 15].any():
            raise ValueError("compact records need topologically numbered nodes (an edge u -> v with v <= u)")
        rec = np.stack([w0, lo, hi], axis=1)
        return rec.view(np.uint8).reshape(len(self), 24).copy(), palette

    def widen(self, n_max: int) -> "CellBatch":
        if n_max == self.n_max:
            return self
        assert n_max > self.n_max
        n = len(self)
        labels = np.full((n, n_max), NO_NODE, dtype=np.uint8)
        succ = np.zeros((n, n_max), dtype=np.uint16)
        labels[:, :self.n_max] = self.labels
        succ[:, :self.n_max] = self.succ
        return CellBatch(n_max, labels, succ)

    def fits_8(self) -> bool:
        return self.n_max == 8 or not len(self) or bool((self.labels[:, 8:] == NO_NODE).all())

    def narrow(self) -> "CellBatch":
        """The 16-byte form of a 16-node batch none of whose cells has more than 8 nodes (else self)."""
        if self.n_max == 8 or not self.fits_8():
            return self
        return CellBatch(8, np.ascontiguousarray(self.labels[:, :8]), np.ascontiguousarray(self.succ[:, :8]))

    def to_undirected(self) -> "CellBatch":
        """Symmetrise adjacency (reference kernels/utils.py:511-521 via `undirected=True`)."""
        s = self.succ.astype(np.uint32)
        pred = np.zeros_like(s)
        for u in range(self.n_max):
            for v in range(self.n_max):
                pred[:, v] |= ((s[:, u] >> v) & 1) << u
        return CellBatch(self.n_max, self.labels, (s | pred).astype(np.uint16))

    def to_networkx(self, vocab: OpVocab, node_label: str = "op_name"):
        import networkx as nx
        out = []
        for i in range(len(self)):
            g = nx.DiGraph()
            for v in range(self.n_max):
                if self.labels[i, v] != NO_NODE:
                    g.add_node(v, **{node_label: vocab.names[self.labels[i, v]]})
            for v in range(self.n_max):
                m = int(self.succ[i, v])
                for j in range(self.n_max):
                    if (m >> j) & 1:
                        g.add_edge(v, j)
            out.append(g)
        return out


class CompactRecords:
    """A pool in the WIRE form of nbw_model.compact: `data` uint8 [N, 8 * n_slots] for 8-node cells, [N, 24 * n_slots]
    for 16-node cells (host tensor, pin it for full PCIe speed), `palettes[s][c]` = op code of palette index c in slot
    s.  Half the bytes of the 16 / 48-byte records — what matters when a pool arrives from host memory.  Built by
    CellBatch.to_compact()."""

    def __init__(self, data, palettes):
        self.data, self.palettes = data, [np.asarray(p, dtype=np.uint8) for p in palettes]

    def __len__(self):
        return self.data.shape[0]

    @property
    def n_slots(self) -> int:
        return len(self.palettes)


class CellSlots:
    """Candidates made of SEVERAL cells each (e.g. DARTS "normal | reduce"): one CellBatch per slot, all of
    the same length.  Kernel i of a sum reads slot `kernel.slot` (default 0, the reference's behaviour:
    every kernel of a SumKernel sees the same graph list, kernels/combine_kernels.py:36-56).  Device layout:
    the slot records of a candidate are concatenated, (N, n_slots * stride) u8."""

    def __init__(self, slots: Sequence[CellBatch]):
        slots = list(slots)
        assert len(slots) >= 1 and all(len(s) == len(slots[0]) for s in slots), "slots must have equal lengths"
        n_max = max(s.n_max for s in slots)
        self.slots = [s.widen(n_max) for s in slots]
        self.n_max = n_max

    def __len__(self):
        return len(self.slots[0])

    @property
    def n_slots(self) -> int:
        return len(self.slots)

    def __getitem__(self, idx) -> "CellSlots":
        return CellSlots([s[idx] for s in self.slots])

    def widen(self, n_max: int) -> "CellSlots":
        return self if n_max == self.n_max else CellSlots([s.widen(n_max) for s in self.slots])

    def narrow(self) -> "CellSlots":
        """8-node records when no slot holds a larger cell (else self)."""
        if self.n_max == 8 or not all(s.fits_8() for s in self.slots):
            return self
        return CellSlots([s.narrow() for s in self.slots])

    def to_records(self) -> np.ndarray:
        return np.concatenate([s.to_records() for s in self.slots], axis=1)

    @staticmethod
    def from_records(rec: np.ndarray, n_max: int, n_slots: int) -> "CellSlots":
        st = record_stride(n_max)
        rec = np.ascontiguousarray(rec, dtype=np.uint8).reshape(-1, n_slots * st)
        return CellSlots([CellBatch.from_records(rec[:, i * st:(i + 1) * st], n_max) for i in range(n_slots)])

    @staticmethod
    def concat(items: Iterable["CellSlots"]) -> "CellSlots":
        items = list(items)
        return CellSlots([CellBatch.concat([it.slots[i] for it in items]) for i in range(items[0].n_slots)])


def pack_networkx(graphs, vocab: OpVocab, node_label: str = "op_name",
                  undirected: bool = False, n_max: Optional[int] = None) -> CellBatch:
    """list[nx.DiGraph] -> CellBatch (the conversion the reference does with
    grakel.utils.graph_from_networkx, kernels/weisfilerlehman.py:134-136)."""
    graphs = list(graphs)
    need = max((g.number_of_nodes() for g in graphs), default=0)
    if n_max is None:
        n_max = 8 if need <= 8 else 16
    if need > n_max or n_max not in (8, 16):
        raise ValueError("cells with more than 16 nodes are not supported (got %d)" % need)
    # rows are assembled as Python ints and converted once: element-wise writes into numpy arrays cost
    # several times more than the graph traversal itself (21 -> 7 us per cell)
    lab_rows, succ_rows = [], []
    pad_l, pad_s = [NO_NODE] * n_max, [0] * n_max
    codes, code = vocab._code, vocab.code
    for i, g in enumerate(graphs):
        nodes, adj = g._node, g._adj                    # node -> attributes, node -> {successor: ...}
        order = sorted(nodes)
        k = len(order)
        pos = None if order == list(range(k)) else {v: p for p, v in enumerate(order)}
        row_l = []
        for v in order:
            try:
                name = nodes[v][node_label]
            except KeyError:
                raise ValueError("node %r of graph %d has no %r attribute" % (v, i, node_label))
            c = codes.get(name)
            row_l.append(code(name) if c is None else c)
        if pos is None:
            row_s = [sum(1 << w for w in adj[v]) for v in order]
        else:
            row_s = [sum(1 << pos[w] for w in adj[v]) for v in order]
        if undirected and g.is_directed():              # symmetrise (an undirected nx.Graph already is)
            for p, m in enumerate(list(row_s)):
                q = 0
                while m:
                    if m & 1:
                        row_s[q] |= 1 << p
                    m >>= 1
                    q += 1
        lab_rows.append(row_l + pad_l[k:])
        succ_rows.append(row_s + pad_s[k:])
    n = len(graphs)
    labels = np.array(lab_rows, dtype=np.uint8).reshape(n, n_max)
    succ = np.array(succ_rows, dtype=np.uint16).reshape(n, n_max)
    return CellBatch(n_max, labels, succ)
