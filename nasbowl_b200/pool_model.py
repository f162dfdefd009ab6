"""The fitted surrogate as the fused pool kernel consumes it (nbw_model, include/nbw.h):
per-level WL dictionaries as open-addressing tables, the feature-space posterior operators
G and w_mu, and the scalars of posterior and acquisition.  This is the state that is
broadcast once to every GPU before a pool is scored (parallel.py)."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch

from . import _lib

ACQ = {"EI": 0, "AEI": 1, "UCB": 2, "MEAN": 3}


class PoolModel:
    TENSOR_FIELDS = ("level0", "thermo_base", "max_count", "G", "w_mu", "H", "w_mu_prefix", "child")

    def __init__(self):
        self.n_max = 8
        self.h = 0
        self.oa = False
        self.seed = 0
        self.level0: Optional[torch.Tensor] = None          # int16 storage [256] (u16 values)
        self.tables: List[Optional[torch.Tensor]] = []      # per level (index 0 unused) uint8 [cap*16]
        self.masks: List[int] = []
        self.col_off: List[int] = []
        self.label_off: List[int] = []
        self.thermo_base: Optional[torch.Tensor] = None
        self.max_count: Optional[torch.Tensor] = None
        self.n_features = 0
        self.G: Optional[torch.Tensor] = None                # fp64 [D, D]
        self.w_mu: Optional[torch.Tensor] = None             # fp64 [D + 1]
        self.n_labels = 0
        self.H: Optional[torch.Tensor] = None                # fp64 [T + 1, T + 1] (linear kernel only)
        self.w_mu_prefix: Optional[torch.Tensor] = None      # fp64 [T + 1]
        self.child: Optional[torch.Tensor] = None            # int32 [T]: label -> its only child label (stable levels)
        self.first_stable = 0                                # first WL level on the stable fast path (0 = none)
        self.use_prefix = True                               # False forces the general G path
        self.general_builder = None                          # builds (G, w_mu) on demand (fitting rank only)
        self.lik = 0.0
        self.y_mean = 0.0
        self.y_std = 1.0
        self.incumbent = 0.0
        self.incumbent_idx = -1

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_fit(cls, eng, fit, h: int) -> "PoolModel":
        """Dictionary side only (enough for nbw_pool_features)."""
        m = cls()
        m.n_max, m.h, m.oa, m.seed = fit.n_max, int(h), bool(fit.oa), int(fit.seed)
        m.level0, m.tables, m.masks = eng.build_tables(fit, h)
        m.col_off = list(fit.col_off[:h + 2])
        m.label_off = list(fit.label_off[:h + 2])
        m.thermo_base, m.max_count = fit.thermo_base, fit.max_count
        m.n_features = fit.k_end(h)
        m.n_labels = fit.label_off[h + 1]
        m._find_stable_suffix(fit, int(h))
        return m

    def _find_stable_suffix(self, fit, h: int):
        """Levels whose training dictionary no longer refines (see nbw_model.child): the first level
        l >= 3 with D_{l-1} == D_{l-2}; WL keeps a stable partition stable, so every deeper level has
        the same size too (checked; otherwise the fast path stays off)."""
        import os
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

    # ------------------------------------------------------------------ C descriptor
    def descriptor(self, acq: str = "MEAN", xi: float = 0.0, beta: float = 0.0,
                   incumbent: Optional[float] = None) -> "_lib.Model":
        d = _lib.Model()
        d.n_max, d.h, d.oa, d.acq = self.n_max, self.h, int(self.oa), ACQ[acq]
        d.seed = self.seed & 0xFFFFFFFFFFFFFFFF
        d.level0_table = self.level0.data_ptr()
        for l in range(self.h + 1):
            d.col_off[l] = self.col_off[l]
            d.label_off[l] = self.label_off[l]
            if l >= 1:
                d.table[l] = self.tables[l].data_ptr()
                d.table_mask[l] = self.masks[l]
        if self.oa:
            d.thermo_base = self.thermo_base.data_ptr()
            d.max_count = self.max_count.data_ptr()
        d.n_features = self.n_features
        prefix = self.H is not None and self.use_prefix and not self.oa
        if self.G is None and not prefix and not (self.H is None and self.general_builder is None):
            # (a model with neither G, H nor a builder is dictionary-only: nbw_pool_features)
            if self.general_builder is None:
                raise RuntimeError("this pool model carries only the prefix operators (H); the general "
                                   "path needs G, which only the rank that fitted the GP can build")
            self.G, self.w_mu = self.general_builder()
        if self.G is not None:
            d.G, d.ld_g, d.w_mu = self.G.data_ptr(), self.G.stride(0), self.w_mu.data_ptr()
        d.n_labels = self.n_labels
        if prefix:
            d.H, d.ld_h, d.w_mu_prefix = self.H.data_ptr(), self.H.stride(0), self.w_mu_prefix.data_ptr()
        if self.first_stable and self.child is not None:
            d.child, d.first_stable = self.child.data_ptr(), int(self.first_stable)
        d.lik, d.y_mean, d.y_std = float(self.lik), float(self.y_mean), float(self.y_std)
        d.incumbent = float(self.incumbent if incumbent is None else incumbent)
        d.xi, d.beta = float(xi), float(beta)
        return d

    # ------------------------------------------------------------------ (de)serialisation for broadcast
    def meta(self) -> dict:
        shapes = {f: (None if getattr(self, f) is None else (tuple(getattr(self, f).shape), str(getattr(self, f).dtype)))
                  for f in self.TENSOR_FIELDS}
        return dict(n_max=self.n_max, h=self.h, oa=self.oa, seed=self.seed, masks=list(self.masks),
                    col_off=list(self.col_off), label_off=list(self.label_off), n_features=self.n_features,
                    n_labels=self.n_labels, first_stable=self.first_stable,
                    lik=self.lik, y_mean=self.y_mean, y_std=self.y_std, incumbent=self.incumbent,
                    incumbent_idx=self.incumbent_idx, shapes=shapes,
                    table_sizes=[0 if t is None else int(t.numel()) for t in self.tables])

    def tensors(self) -> List[torch.Tensor]:
        out = [getattr(self, f) for f in self.TENSOR_FIELDS if getattr(self, f) is not None]
        out += [t for t in self.tables if t is not None]
        return out

    @classmethod
    def empty_from_meta(cls, meta: dict, device) -> "PoolModel":
        m = cls()
        for k in ("n_max", "h", "oa", "seed", "masks", "col_off", "label_off", "n_features", "n_labels", "first_stable",
                  "lik", "y_mean", "y_std", "incumbent", "incumbent_idx"):
            setattr(m, k, meta[k])
        for f, sd in meta["shapes"].items():
            if sd is not None:
                shape, dtype = sd
                setattr(m, f, torch.empty(shape, dtype=getattr(torch, dtype.split(".")[-1]), device=device))
        m.tables = [None if s == 0 else torch.empty(s, dtype=torch.uint8, device=device) for s in meta["table_sizes"]]
        return m

    def nbytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())
