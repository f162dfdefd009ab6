"""The fitted surrogate as the fused pool kernels consume it (nbw_model, include/nbw.h): per WL
kernel of the sum ("part") the level dictionaries — open-addressing tables for the first-generation
kernel, id-hash tables + tag buckets for the packed kernel — and, once for the whole sum, the
feature-space posterior operators and the scalars of posterior and acquisition.  This is the state
that is broadcast once to every GPU before a pool is scored (parallel.py)."""
from __future__ import annotations

import ctypes as C
import json
import os
from typing import List, Optional

import numpy as np
import torch

from . import _lib

ACQ = {"EI": 0, "AEI": 1, "UCB": 2, "MEAN": 3}

_SCALARS = ("n_max", "h", "oa", "seed", "masks", "bucket_masks", "col_off", "label_off", "n_features", "n_labels",
            "first_stable", "use_prefix", "lik", "y_mean", "y_std", "incumbent", "incumbent_idx", "slot", "n_slots",
            "undirected", "level_weight", "label_base", "total_labels", "oa_d2_base")


class PoolModel:
    """One part of the sum (and, on the first part, the operators of the whole sum: `more` lists the
    further parts)."""
    TENSOR_FIELDS = ("level0", "thermo_base", "max_count", "G", "w_mu", "H", "w_mu_prefix", "child", "final_label",
                     "oa_exc")
    LIST_FIELDS = ("tables", "id_hash", "bucket_tags", "bucket_entries")

    def __init__(self):
        self.n_max = 8
        self.h = 0
        self.oa = False
        self.seed = 0
        self.level0: Optional[torch.Tensor] = None          # int16 storage [256] (u16 values)
        self.tables: List[Optional[torch.Tensor]] = []      # per level (index 0 unused) uint8 [cap*16]
        self.masks: List[int] = []
        self.id_hash: List[Optional[torch.Tensor]] = []     # per level l >= 1: uint8 [D_{l-1} * 16]
        self.bucket_tags: List[Optional[torch.Tensor]] = []     # per level: uint8 [n_buckets * 16]
        self.bucket_entries: List[Optional[torch.Tensor]] = []  # per level: uint8 [n_buckets * 64]
        self.bucket_masks: List[int] = []
        self.col_off: List[int] = []
        self.label_off: List[int] = []
        self.thermo_base: Optional[torch.Tensor] = None
        self.max_count: Optional[torch.Tensor] = None
        self.n_features = 0
        self.G: Optional[torch.Tensor] = None                # fp64 [D, D]
        self.w_mu: Optional[torch.Tensor] = None             # fp64 [D + 1]
        self.n_labels = 0                                    # stacked labels of THIS part
        self.total_labels = 0                                # ... of all parts (first part only)
        self.label_base = 0                                  # first stacked index of this part in H
        self.H: Optional[torch.Tensor] = None                # fp64 [T + 1, T + 1] over all parts (linear kernels only)
        self.w_mu_prefix: Optional[torch.Tensor] = None      # fp64 [T + 1]
        self.child: Optional[torch.Tensor] = None            # int32 [T_part]: label -> its only child label (stable levels)
        self.final_label: Optional[torch.Tensor] = None      # int32 [T_part]: stable label -> global index of its level-h descendant
        self.oa_d2_base = 0                                  # oa item form: first rank-2 chain item (0 = none)
        self.oa_exc: Optional[torch.Tensor] = None           # oa item form: int32 [T] (first exception item << 5 | max count)
        self.first_stable = 0                                # first WL level on the stable fast path (0 = none)
        self.use_prefix = True                               # False forces the general G path
        self.general_builder = None                          # builds (G, w_mu) on demand (fitting rank only)
        self.slot = 0                                        # record of the candidate this part reads
        self.n_slots = 1                                     # records per candidate (first part only)
        self.undirected = False
        self.level_weight: Optional[List[float]] = None      # kernel weight x layer weight per level (None = ones)
        self.more: List["PoolModel"] = []                    # further parts of the sum
        self.lik = 0.0
        self.y_mean = 0.0
        self.y_std = 1.0
        self.incumbent = 0.0
        self.incumbent_idx = -1

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_fit(cls, eng, fit, h: int, packed: bool = True) -> "PoolModel":
        """Dictionary side only (enough for nbw_pool_features)."""
        m = cls()
        m.n_max, m.h, m.oa, m.seed = fit.n_max, int(h), bool(fit.oa), int(fit.seed)
        m.level0, m.tables, m.masks = eng.build_tables(fit, h)
        m.col_off = list(fit.col_off[:h + 2])
        m.label_off = list(fit.label_off[:h + 2])
        m.thermo_base, m.max_count = fit.thermo_base, fit.max_count
        m.n_features = fit.k_end(h)
        m.n_labels = m.total_labels = fit.label_off[h + 1]
        m._find_stable_suffix(fit, int(h))
        if packed:
            last_hashed = h if not m.first_stable else m.first_stable - 1
            m.id_hash, m.bucket_tags, m.bucket_entries, m.bucket_masks = eng.build_packed_tables(fit, last_hashed)
        return m

    def _find_stable_suffix(self, fit, h: int):
        """Levels whose training dictionary no longer refines (see nbw_model.child): the first level
        l >= 3 with D_{l-1} == D_{l-2}; WL keeps a stable partition stable, so every deeper level has
        the same size too (checked; otherwise the fast path stays off)."""
        dims = list(fit.dims[:h + 1])
        first = 0
        for l in range(3, h + 1):
            if dims[l - 1] == dims[l - 2]:
                first = l
                break
        if not first or os.environ.get("NBW_NO_STABLE") or fit.parents is None:
            return
        if any(dims[l] != dims[first - 1] for l in range(first, h + 1)):
            return
        T = fit.label_off[h + 1]
        child = torch.zeros((max(T, 1),), dtype=torch.int32, device=fit.parents.device)
        for l in range(first, h + 1):
            par = fit.parents[fit.label_off[l]:fit.label_off[l + 1]].long()          # parent of each level-l label
            z = torch.arange(dims[l], dtype=torch.int32, device=par.device)
            child[fit.label_off[l - 1] + par] = z
        self.child, self.first_stable = child, first
        if not self.label_off:
            self.h, self.label_off = h, list(fit.label_off[:h + 2])
        self._set_final_labels()

    def _set_final_labels(self):
        """final_label[label_off[l] + z] = GLOBAL stacked index of the level-h descendant of the level-l
        label z, for the stable levels l >= first_stable (the child maps composed down to level h)."""
        if not self.first_stable or self.child is None:
            return
        h, lo = self.h, self.label_off
        final = torch.zeros_like(self.child)
        D = lo[h + 1] - lo[h]
        desc = torch.arange(D, dtype=torch.int32, device=self.child.device) + (self.label_base + lo[h])
        final[lo[h]:lo[h + 1]] = desc
        for l in range(h - 1, self.first_stable - 1, -1):
            desc = desc[self.child[lo[l]:lo[l + 1]].long()]
            final[lo[l]:lo[l + 1]] = desc
        self.final_label = final

    def set_label_base(self, base: int):
        self.label_base = int(base)
        self._set_final_labels()

    @property
    def parts(self) -> List["PoolModel"]:
        return [self] + list(self.more)

    # ------------------------------------------------------------------ C descriptor
    def _part_descriptor(self, rec: int, palettes=None) -> "_lib.Model":
        d = _lib.Model()
        if palettes is not None:
            d.compact = 1
            for c in range(8 if self.n_max == 8 else 16):
                d.palette[c] = int(palettes[self.slot][c])
        d.n_max, d.h, d.oa = self.n_max, self.h, int(self.oa)
        d.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        d.level0_table = self.level0.data_ptr()
        for l in range(self.h + 1):
            d.col_off[l] = self.col_off[l]
            d.label_off[l] = self.label_off[l]
            if l >= 1:
                d.table[l] = self.tables[l].data_ptr()
                d.table_mask[l] = self.masks[l]
                if l < len(self.bucket_tags) and self.bucket_tags[l] is not None:
                    d.id_hash[l] = self.id_hash[l].data_ptr()
                    d.bucket_tags[l] = self.bucket_tags[l].data_ptr()
                    d.bucket_entries[l] = self.bucket_entries[l].data_ptr()
                    d.bucket_mask[l] = self.bucket_masks[l]
        if self.oa:
            d.thermo_base = self.thermo_base.data_ptr()
            d.max_count = self.max_count.data_ptr()
        d.n_features = self.n_features
        d.n_labels = self.n_labels
        if self.first_stable and self.child is not None:
            d.child, d.first_stable = self.child.data_ptr(), int(self.first_stable)
            if self.final_label is not None:
                d.final_label = self.final_label.data_ptr()
        lw = self.level_weight
        d.weighted = int(lw is not None and any(float(w) != 1.0 for w in lw))
        for l in range(self.h + 1):
            d.level_weight[l] = 1.0 if lw is None else float(lw[l])
        d.undirected = int(self.undirected)
        d.rec_offset = self.slot * rec
        d.label_base = self.label_base
        return d

    def descriptor(self, acq: str = "MEAN", xi: float = 0.0, beta: float = 0.0,
                   incumbent: Optional[float] = None, palettes=None) -> "_lib.Model":
        """palettes (one uint8[8] / uint8[16] per record slot): the pool arrives as 8 / 24-byte compact records
        (nbw_model.compact)."""
        rec = (16 if self.n_max == 8 else 48) if palettes is None else (8 if self.n_max == 8 else 24)
        d = self._part_descriptor(rec, palettes)
        d.acq = ACQ[acq]
        multi = len(self.more) > 0
        legacy = bool(os.environ.get("NBW_SCORE_LEGACY"))
        prefix = self.H is not None and self.use_prefix and (not self.oa or (self.oa_exc is not None and not legacy))
        if multi and not prefix:
            raise NotImplementedError("a sum of kernels is scored in prefix form (linear WL kernels)")
        if self.G is None and not prefix and not (self.H is None and self.general_builder is None):
            # (a model with neither G, H nor a builder is dictionary-only: nbw_pool_features)
            if self.general_builder is None:
                raise RuntimeError("this pool model carries only the prefix operators (H); the general "
                                   "path needs G, which only the rank that fitted the GP can build")
            self.G, self.w_mu = self.general_builder()
        if self.G is not None:
            d.G, d.ld_g, d.w_mu = self.G.data_ptr(), self.G.stride(0), self.w_mu.data_ptr()
        if prefix:
            d.H, d.ld_h, d.w_mu_prefix = self.H.data_ptr(), self.H.stride(0), self.w_mu_prefix.data_ptr()
            d.n_labels = self.total_labels or self.n_labels
            if self.oa:
                d.oa_exc = self.oa_exc.data_ptr()
                d.oa_d2_base = int(self.oa_d2_base)
        d.rec_stride = self.n_slots * rec
        d.lik, d.y_mean, d.y_std = float(self.lik), float(self.y_mean), float(self.y_std)
        d.incumbent = float(self.incumbent if incumbent is None else incumbent)
        d.xi, d.beta = float(xi), float(beta)
        chain = []
        prev = d
        for part in self.more:
            pd = part._part_descriptor(rec, palettes)
            chain.append(pd)
            prev.next = C.addressof(pd)
            prev = pd
        d._chain = chain            # keeps the chained descriptors alive as long as the first one
        return d

    # ------------------------------------------------------------------ (de)serialisation for broadcast
    def _meta_one(self) -> dict:
        md = {k: getattr(self, k) for k in _SCALARS}
        md["tensors"] = {f: (None if getattr(self, f) is None else
                             (list(getattr(self, f).shape), str(getattr(self, f).dtype).split(".")[-1]))
                         for f in self.TENSOR_FIELDS}
        md["lists"] = {f: [0 if t is None else int(t.numel()) for t in getattr(self, f)] for f in self.LIST_FIELDS}
        return md

    def meta(self) -> dict:
        md = self._meta_one()
        md["more"] = [p._meta_one() for p in self.more]
        return md

    def _tensors_one(self) -> List[torch.Tensor]:
        out = [getattr(self, f) for f in self.TENSOR_FIELDS if getattr(self, f) is not None]
        for f in self.LIST_FIELDS:
            out += [t for t in getattr(self, f) if t is not None]
        return out

    def tensors(self) -> List[torch.Tensor]:
        out = self._tensors_one()
        for p in self.more:
            out += p._tensors_one()
        return out

    @classmethod
    def _empty_one(cls, md: dict, device) -> "PoolModel":
        m = cls()
        for k in _SCALARS:
            setattr(m, k, md[k])
        for f, sd in md["tensors"].items():
            if sd is not None:
                shape, dtype = sd
                setattr(m, f, torch.empty(tuple(shape), dtype=getattr(torch, dtype), device=device))
        for f, sizes in md["lists"].items():
            setattr(m, f, [None if s == 0 else torch.empty(s, dtype=torch.uint8, device=device) for s in sizes])
        return m

    @classmethod
    def empty_from_meta(cls, meta: dict, device) -> "PoolModel":
        m = cls._empty_one(meta, device)
        m.more = [cls._empty_one(md, device) for md in meta.get("more", [])]
        return m

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())

    # ---- flat form: ONE byte buffer (json header + 256-byte aligned tensor payloads) = one collective
    def flat_layout(self):
        """(offsets, total bytes) of the tensors() inside the flat buffer, after `header_bytes`."""
        offs, pos = [], 0
        for t in self.tensors():
            pos = (pos + 255) & ~255
            offs.append(pos)
            pos += t.numel() * t.element_size()
        return offs, (pos + 255) & ~255

    def pack_flat(self) -> torch.Tensor:
        """uint8 device buffer [16 + header + payload]: two int64 (header length, payload length), the json
        meta, then every tensor at its aligned offset."""
        header = json.dumps(self.meta()).encode()
        hlen = (len(header) + 255) & ~255
        offs, plen = self.flat_layout()
        ts = self.tensors()
        dev = ts[0].device
        buf = torch.empty((256 + hlen + plen,), dtype=torch.uint8, device=dev)
        head = np.zeros(256 + hlen, dtype=np.uint8)
        head[:16] = np.array([len(header), plen], dtype=np.int64).view(np.uint8)
        head[256:256 + len(header)] = np.frombuffer(header, dtype=np.uint8)
        buf[:256 + hlen].copy_(torch.from_numpy(head), non_blocking=True)
        base = 256 + hlen
        for t, o in zip(ts, offs):
            nb = t.numel() * t.element_size()
            buf[base + o:base + o + nb].copy_(t.contiguous().view(torch.uint8).reshape(-1), non_blocking=True)
        return buf

    @classmethod
    def unpack_flat(cls, buf: torch.Tensor, header: bytes) -> "PoolModel":
        """Rebuild a model whose tensors are VIEWS into `buf` (the receiver keeps the buffer alive through them)."""
        meta = json.loads(header.decode())
        m = cls.empty_from_meta(meta, "meta")
        hlen = (len(header) + 255) & ~255
        base = 256 + hlen
        offs, _ = m.flat_layout()

        def views(model, it):
            for f in model.TENSOR_FIELDS:
                t = getattr(model, f)
                if t is not None:
                    o = next(it)
                    nb = t.numel() * t.element_size()
                    setattr(model, f, buf[base + o:base + o + nb].view(t.dtype).reshape(t.shape))
            for f in model.LIST_FIELDS:
                lst = getattr(model, f)
                for i, t in enumerate(lst):
                    if t is not None:
                        o = next(it)
                        lst[i] = buf[base + o:base + o + t.numel()]
        it = iter(offs)
        views(m, it)
        for p in m.more:
            views(p, it)
        return m
