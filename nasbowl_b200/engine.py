"""Device engine: drives libnbw's C-ABI (include/nbw.h) with torch-owned device memory.

PyTorch is used here only as plumbing — allocation, streams, host<->device copies and (in
parallel.py) torch.distributed.  Every arithmetic step is a hand-written sm_100a kernel
behind the C-ABI; nothing in this module computes on the CPU or with torch operators.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from .cells import CellBatch, CellSlots, record_stride

ACQ = {"EI": 0, "AEI": 1, "UCB": 2, "MEAN": 3}
DEFAULT_SEED = 0x5EEDB200


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def _pow2_at_least(x: int) -> int:
    p = 1
    while p < x:
        p *= 2
    return p


@dataclass
class WLFit:
    """A fitted WL vocabulary over a batch of cells plus the batch's dense features."""
    n_cells: int
    n_max: int
    h: int
    oa: bool
    seed: int
    ids: torch.Tensor                  # int32 storage [h+1, n_cells * n_max] (u32 values)
    dims: List[int]                    # D_l: distinct labels per level
    uniq_keys: List[torch.Tensor]      # per level, int64 storage (u64) sorted ascending
    uniq_chks: List[torch.Tensor]
    label_off: List[int]               # stacked label offsets, len h+2
    widths: List[int]                  # feature columns per level (D_l, or thermometer width)
    col_off: List[int]                 # feature column offsets (levels padded to 128), len h+2
    phi: torch.Tensor                  # int8 [n_cells, ld_phi]
    self_sim: torch.Tensor             # int32 [n_cells]  = K_raw(g, g) at depth h
    thermo_base: Optional[torch.Tensor] = None   # int32 [total labels] (oa)
    max_count: Optional[torch.Tensor] = None     # uint8 [total labels] (oa)
    op_codes: Optional[np.ndarray] = None        # sorted op codes present (level-0 dictionary)
    parents: Optional[torch.Tensor] = None       # int32 [total labels]: ancestor (level l-1 id) of each label

    @property
    def ld_phi(self) -> int:
        return self.phi.shape[1]

    def k_end(self, h: int) -> int:
        """Feature columns [0, k_end(h)) hold WL levels 0..h."""
        return self.col_off[h + 1]


class Engine:
    """One per process / GPU (the C-ABI context is bound to the current CUDA device)."""
    _instances = {}

    @classmethod
    def get(cls, device: Optional[int] = None) -> "Engine":
        if not torch.cuda.is_available():
            raise RuntimeError("nasbowl_b200 needs a CUDA device (B200, sm_100a); there is no CPU path")
        if device is None:
            device = torch.cuda.current_device()
        eng = cls._instances.get(device)
        if eng is None:
            eng = cls(device)
            cls._instances[device] = eng
        return eng

    def __init__(self, device: int):
        self.lib = _lib.load()
        self.device_index = int(device)
        self.device = torch.device("cuda", self.device_index)
        torch.cuda.set_device(self.device)
        torch.zeros(1, device=self.device)          # make sure the primary context exists
        ctx = C.c_void_p()
        rc = self.lib.nbw_create(C.byref(ctx), self.device_index)
        if rc != 0:
            raise _lib.NbwError(rc, "nbw_create failed on device %d" % self.device_index)
        self.ctx = ctx

    # ------------------------------------------------------------------ helpers
    def _check(self, rc: int):
        if rc == 0:
            return
        msg = self.lib.nbw_last_error(self.ctx).decode("utf-8", "replace")
        if rc == 3:
            raise _lib.NotPositiveDefinite(rc, msg)
        if rc == 4:
            raise _lib.HashCollision(rc, msg)
        raise _lib.NbwError(rc, msg)

    @property
    def stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    @property
    def launches(self) -> int:
        return int(self.lib.nbw_launch_count(self.ctx))

    @staticmethod
    def _p(t: Optional[torch.Tensor]):
        return None if t is None else C.c_void_p(t.data_ptr())

    def empty(self, shape, dtype):
        return torch.empty(shape, dtype=dtype, device=self.device)

    def upload_cells(self, batch) -> torch.Tensor:
        """CellBatch | CellSlots -> device records (N, stride) / (N, n_slots * stride)."""
        rec = torch.from_numpy(batch.to_records())
        return rec.to(self.device, non_blocking=False)

    # ------------------------------------------------------------------ WL fit (train side)
    def wl_fit(self, cells: torch.Tensor, n_max: int, h: int, oa: bool = False,
               seed: int = DEFAULT_SEED, build_features: bool = True) -> WLFit:
        """WL relabelling + dictionary build + dense features for a batch of cells.
        Mirrors WeisfeilerLehman.parse_input (grakel_replace/weisfeiler_lehman.py:140-328)."""
        assert cells.dtype == torch.uint8 and cells.is_cuda and cells.is_contiguous()
        stride = record_stride(n_max)
        n = cells.numel() // stride
        if n == 0:
            raise ValueError("parsed input is empty")          # weisfeiler_lehman.py:205-206
        if not (0 <= h < _lib.NBW_MAX_LEVELS):
            raise TypeError("'h' must be a non-negative integer below %d. Got h%s" % (_lib.NBW_MAX_LEVELS, h))
        for attempt in range(4):
            try:
                return self._wl_fit_once(cells, n, n_max, h, oa, seed + attempt * 0x9E3779B1, build_features)
            except _lib.HashCollision:
                continue
        raise RuntimeError("WL hashing collided for 4 different seeds")

    def _wl_fit_once(self, cells, n, n_max, h, oa, seed, build_features) -> WLFit:
        lib, ctx, st = self.lib, self.ctx, self.stream
        n_keys = n * n_max
        ids = self.empty((h + 1, n_keys), torch.int32)
        # all levels in one call, one sync (nbw_wl_fit); the dictionaries are views into two level-major buffers
        uk = self.empty((h + 1, n_keys), torch.int64)
        uc = self.empty((h + 1, n_keys), torch.int64)
        dd = (C.c_int64 * (h + 1))()
        self._check(lib.nbw_wl_fit(ctx, self._p(cells), n, n_max, h, C.c_uint64(seed & 0xFFFFFFFFFFFFFFFF), self._p(ids),
                                   self._p(uk), self._p(uc), dd, st))
        dims = [int(v) for v in dd]
        ukeys = [uk[l, :dims[l]] for l in range(h + 1)]
        uchks = [uc[l, :dims[l]] for l in range(h + 1)]
        label_off = [0]
        for d in dims:
            label_off.append(label_off[-1] + d)
        # ancestor map: the level l-1 label every level-l label refines
        parents = torch.zeros((max(label_off[-1], 1),), dtype=torch.int32, device=self.device)
        for l in range(1, h + 1):
            if dims[l] > 0:
                self._check(lib.nbw_label_parents(ctx, self._p(ids[l - 1]), self._p(ids[l]), n_keys,
                                                  self._p(parents[label_off[l]:]), st))
        thermo_base = max_count = None
        widths = list(dims)
        if oa:
            total = label_off[-1]
            thermo_base = self.empty((max(total, 1),), torch.int32)
            max_count = self.empty((max(total, 1),), torch.uint8)
            lo = (C.c_int64 * (h + 2))(*label_off)
            wd = (C.c_int64 * (h + 1))()
            self._check(lib.nbw_oa_thermo_layout(ctx, self._p(ids), n, n_max, h + 1, lo, self._p(thermo_base),
                                                 self._p(max_count), wd, st))
            widths = [int(w) for w in wd]
        col_off = [0]
        for w in widths:
            col_off.append(col_off[-1] + _round_up(max(w, 1), 128))
        op_codes = ukeys[0].cpu().numpy().astype(np.int64)
        fit = WLFit(n_cells=n, n_max=n_max, h=h, oa=bool(oa), seed=seed, ids=ids, dims=dims,
                    uniq_keys=ukeys, uniq_chks=uchks, label_off=label_off, widths=widths,
                    col_off=col_off, phi=None, self_sim=None, thermo_base=thermo_base,
                    max_count=max_count, op_codes=op_codes, parents=parents)
        if build_features:
            self.build_features(fit)
        return fit

    def wl_fit_append(self, prev: WLFit, cells: torch.Tensor, build_features: bool = True) -> Optional[WLFit]:
        """WL fit of a batch whose first prev.n_cells cells are the batch `prev` was fitted on, with the
        dictionary grown APPEND-ONLY: every label of the previous fit keeps its id (hence its feature column and
        its place in the pool operators), labels that the new cells introduce get the next ids of their level.
        Keys are canonical (a label is hashed through the key of its previous label, wl_relabel.cu), so the
        fresh dictionary of the whole batch is matched against the previous one key by key on the host (a few
        thousand u64) and the node ids are renumbered on the device (nbw_remap_ids).  Returns None when the
        hashing collided for the previous seed (the caller refits from scratch)."""
        n_max, h = prev.n_max, prev.h
        n = cells.numel() // record_stride(n_max)
        try:
            fresh = self._wl_fit_once(cells, n, n_max, h, prev.oa, prev.seed, build_features=False)
        except _lib.HashCollision:
            return None
        n_keys = n * n_max
        ids = self.empty((h + 1, n_keys), torch.int32)
        dims, ukeys, uchks = [], [], []
        host_keys = getattr(prev, "_host_keys", None) or [k.cpu().numpy().view(np.uint64) for k in prev.uniq_keys]
        new_host = []
        for l in range(h + 1):
            old = host_keys[l]
            order = np.argsort(old, kind="stable")
            old_sorted = old[order]
            new = fresh.uniq_keys[l].cpu().numpy().view(np.uint64)
            pos = np.searchsorted(old_sorted, new)
            pos_c = np.minimum(pos, max(len(old) - 1, 0))
            found = (pos < len(old)) & (old_sorted[pos_c] == new) if len(old) else np.zeros(len(new), dtype=bool)
            mapping = np.empty(len(new), dtype=np.uint32)
            mapping[found] = order[pos_c[found]]
            extra = np.nonzero(~found)[0]
            mapping[extra] = len(old) + np.arange(len(extra), dtype=np.uint32)
            if int(found.sum()) != len(old):
                return None                 # a previous label vanished: the old cells are not a prefix of this batch
            map_dev = torch.from_numpy(mapping.view(np.int32)).to(self.device)
            self._check(self.lib.nbw_remap_ids(self.ctx, self._p(fresh.ids[l]), n_keys, self._p(map_dev), self._p(ids[l]),
                                               self.stream))
            extra_dev = torch.from_numpy(extra.astype(np.int64)).to(self.device)
            ukeys.append(torch.cat([prev.uniq_keys[l], fresh.uniq_keys[l][extra_dev]]))
            uchks.append(torch.cat([prev.uniq_chks[l], fresh.uniq_chks[l][extra_dev]]))
            new_host.append(np.concatenate([old, new[extra]]))
            dims.append(len(old) + len(extra))
        label_off = [0]
        for d in dims:
            label_off.append(label_off[-1] + d)
        parents = torch.zeros((max(label_off[-1], 1),), dtype=torch.int32, device=self.device)
        for l in range(1, h + 1):
            if dims[l] > 0:
                self._check(self.lib.nbw_label_parents(self.ctx, self._p(ids[l - 1]), self._p(ids[l]), n_keys,
                                                       self._p(parents[label_off[l]:]), self.stream))
        col_off = [0]
        for w in dims:
            col_off.append(col_off[-1] + _round_up(max(w, 1), 128))
        fit = WLFit(n_cells=n, n_max=n_max, h=h, oa=prev.oa, seed=prev.seed, ids=ids, dims=dims, uniq_keys=ukeys,
                    uniq_chks=uchks, label_off=label_off, widths=list(dims), col_off=col_off, phi=None, self_sim=None,
                    op_codes=new_host[0].astype(np.int64), parents=parents)
        fit._host_keys = new_host
        if build_features:
            self.build_features(fit)
        return fit

    def build_features(self, fit: WLFit):
        """Dense int8 histogram rows (VertexHistogram.parse_input, vertex_histogram.py:98-209)."""
        n, h = fit.n_cells, fit.h
        ld = fit.col_off[-1]
        fit.phi = self.empty((n, ld), torch.int8)
        fit.self_sim = self.empty((n,), torch.int32)
        co = (C.c_int64 * (h + 2))(*fit.col_off)
        lo = (C.c_int64 * (h + 2))(*fit.label_off)
        self._check(self.lib.nbw_hist_dense_i8(self.ctx, self._p(fit.ids), n, fit.n_max, h + 1, co, lo,
                                               self._p(fit.thermo_base), int(fit.oa), self._p(fit.phi), ld,
                                               self._p(fit.self_sim), self.stream))

    # ------------------------------------------------------------------ integer Gram
    def gram(self, A: torch.Tensor, B: torch.Tensor, k0: int, k1: int, impl: int = 0,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """int32 [m, n] = A[:, k0:k1] @ B[:, k0:k1].T  (exact)."""
        assert A.dtype == torch.int8 and B.dtype == torch.int8
        m, n = A.shape[0], B.shape[0]
        if out is None:
            out = self.empty((m, _round_up(n, 4)), torch.int32)
        self._check(self.lib.nbw_gram_i8(self.ctx, self._p(A), m, A.stride(0), self._p(B), n, B.stride(0),
                                         k0, k1, self._p(out), out.stride(0), impl, self.stream))
        return out[:, :n]

    GRAM_OUT = {"s32": (0, torch.int32), "u8": (1, torch.uint8), "s16": (2, torch.int16), "f32norm": (3, torch.float32)}

    def gram_ex(self, A: torch.Tensor, B: torch.Tensor, k0: int, k1: int, mode: str = "s32", upper_only: bool = False,
                row_offset: int = 0, scale_a: Optional[torch.Tensor] = None, scale_b: Optional[torch.Tensor] = None,
                out: Optional[torch.Tensor] = None, overflow: Optional[torch.Tensor] = None) -> torch.Tensor:
        """nbw_gram_i8_ex: A[:, k0:k1] @ B[:, k0:k1].T with the output form chosen per launch (s32 / u8 / s16 with
        saturation / cosine-normalised fp32) and, for the Gram of a batch with itself, the symmetric tile schedule
        (tiles strictly below the global diagonal are neither computed nor written)."""
        code, dtype = self.GRAM_OUT[mode]
        m, n = A.shape[0], B.shape[0]
        if out is None:
            out = self.empty((m, _round_up(n, 16)), dtype)
        assert out.dtype == dtype and out.shape[0] >= m and out.stride(1) == 1
        self._check(self.lib.nbw_gram_i8_ex(self.ctx, self._p(A), m, A.stride(0), self._p(B), n, B.stride(0), k0, k1,
                                            self._p(out), out.stride(0), code, int(upper_only), int(row_offset),
                                            self._p(scale_a), self._p(scale_b), self._p(overflow), self.stream))
        return out[:m, :n]

    def sym_mirror(self, Cm: torch.Tensor):
        """Fill the lower triangle of a square matrix from its upper triangle (nbw_sym_mirror)."""
        n = Cm.shape[0]
        self._check(self.lib.nbw_sym_mirror(self.ctx, self._p(Cm), n, Cm.stride(0), Cm.element_size(), self.stream))

    def gram_stress(self, fit: WLFit, D: int, block: int, rank: int = 0, world: int = 1, mode: str = "u8",
                    symmetric: bool = True):
        """BASELINE configs[4]: the N x N Gram of a fitted batch, materialised in HBM as `mode` elements.  Row blocks
        of `block` rows are dealt round-robin to the ranks; with `symmetric` only the tiles on or above the
        diagonal are computed and (one GPU) the lower triangle is filled by nbw_sym_mirror afterwards.  Returns
        timings (CUDA events), the result-independent checks and a description for the bench record."""
        import time
        N = fit.n_cells
        code, dtype = self.GRAM_OUT[mode]
        if world > 1 and symmetric:
            block = min(block, 4096)        # smaller blocks deal out more evenly (and 4096 is the fastest on one GPU too)
        from .parallel import deal_row_blocks
        mine = deal_row_blocks(N, block, world, symmetric)[rank]
        ld = _round_up(N, 16)
        rows_mine = sum(min(block, N - r0) for r0 in mine)
        out = self.empty((max(rows_mine, 1), ld), dtype)          # this rank's row blocks, stacked
        overflow = torch.zeros((1,), dtype=torch.int32, device=self.device)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        # warm-up: one small block (module load, tensor-map encode, launch attributes)
        self.gram_ex(fit.phi[:min(512, N)], fit.phi[:min(512, N)], 0, D, mode=mode)
        torch.cuda.synchronize()
        t_begin = time.time()
        ev[0].record()
        pos, tiles_done, tiles_all, placed = 0, 0, 0, []
        for r0 in mine:
            rows = min(block, N - r0)
            self.gram_ex(fit.phi[r0:r0 + rows], fit.phi, 0, D, mode=mode, upper_only=symmetric, row_offset=r0,
                         out=out[pos:pos + rows], overflow=overflow)
            placed.append((r0, rows, pos))
            pos += rows
            for mt in range((rows + 255) // 256):
                nt_all = (N + 255) // 256
                first = 0
                if symmetric:       # tiles with nt * 256 + 255 < r0 + mt * 256 are skipped
                    first = min(nt_all, max(0, (r0 + mt * 256 - 255 + 255) // 256))
                tiles_all += nt_all
                tiles_done += nt_all - first
        if symmetric and world == 1:
            self.sym_mirror(out[:N])
        ev[1].record()
        torch.cuda.synchronize()
        t_end = time.time()
        ms = ev[0].elapsed_time(ev[1])
        # checks that do not need the whole oracle: diagonal == self-similarity; mirrored entries equal;
        # SIMT cross-check of a corner; no saturation
        checks = {"no_saturation": bool(int(overflow.item()) == 0)}
        ok_diag = True
        for r0, rows, p in placed:
            d = out[p:p + rows][torch.arange(rows, device=self.device), torch.arange(r0, r0 + rows, device=self.device)]
            ok_diag &= bool(torch.equal(d.to(torch.int32), fit.self_sim[r0:r0 + rows]))
        checks["diag_equals_self_similarity"] = ok_diag
        if placed:
            r0, rows, p = placed[0]
            c = min(rows, 2048)
            ref = self.gram(fit.phi[r0:r0 + c], fit.phi[N - c:], 0, D, impl=1)          # SIMT dp4a, far corner
            checks["simt_crosscheck_corner"] = bool(torch.equal(ref, out[p:p + c, N - c:N].to(torch.int32)))
            if symmetric and world == 1 and N > c:
                low = out[N - c:N, r0:r0 + c].to(torch.int32)                            # filled by the mirror pass
                checks["mirrored_corner_equals_transpose"] = bool(torch.equal(low, ref.t().contiguous()))
        return {"ms": ms, "checks": checks, "t_begin": t_begin, "t_end": t_end,
                "computed_fraction": tiles_done / max(tiles_all, 1), "dims": list(fit.dims),
                "schedule": (("symmetric: 256x256 tiles strictly below the diagonal skipped, " +
                              ("lower triangle filled by a mirror pass" if world == 1 else
                               "row blocks of %d dealt to the ranks by tile count (largest first, least loaded rank)" % block))
                             if symmetric else "all tiles"),
                "row_block": block,
                "output": "%s [%d x %d] materialised in HBM (%.1f GB on this rank)" % (mode, rows_mine, N, rows_mine * ld * out.element_size() / 1e9),
                "output_bytes": rows_mine * N * out.element_size() * world if world > 1 else N * N * out.element_size(),
                "kernel": "gram_i8_pair_kernel<%s>" % mode, "result": out}

    # ------------------------------------------------------------------ fp64 linear algebra
    def gram_accumulate(self, K: torch.Tensor, w: float, beta: float, out: torch.Tensor):
        m, n = K.shape
        self._check(self.lib.nbw_gram_accumulate(self.ctx, self._p(K), K.stride(0), m, n, float(w), float(beta),
                                                 self._p(out), out.stride(0), self.stream))

    def axpby(self, X: torch.Tensor, a: float, b: float, Y: torch.Tensor):
        m, n = X.shape
        self._check(self.lib.nbw_axpby_f64(self.ctx, self._p(X), X.stride(0), m, n, float(a), float(b),
                                           self._p(Y), Y.stride(0), self.stream))

    def gram_normalize(self, K: torch.Tensor) -> torch.Tensor:
        n = K.shape[0]
        s = self.empty((n,), torch.float64)
        self._check(self.lib.nbw_gram_normalize(self.ctx, self._p(K), K.stride(0), n, self._p(s), self.stream))
        return s

    def gemm(self, ta: bool, tb: bool, m: int, n: int, k: int, alpha: float, A: torch.Tensor, lda: int,
             B: torch.Tensor, ldb: int, beta: float, Cm: torch.Tensor, ldc: int):
        self._check(self.lib.nbw_gemm_f64(self.ctx, int(ta), int(tb), m, n, k, float(alpha), self._p(A), lda,
                                          self._p(B), ldb, float(beta), self._p(Cm), ldc, self.stream))

    def gp_factor(self, K: torch.Tensor, lik: float, y: torch.Tensor, want_inverse: bool = True):
        """One attempt of compute_pd_inverse (bayesopt/utils.py:127-140) at jitter `lik`.
        Returns (L, Linv, alpha, logdet, yKy, trKinv, |alpha|^2); raises NotPositiveDefinite.
        want_inverse=False: likelihood terms only (Linv = alpha = None, the last two stats 0)."""
        n = K.shape[0]
        L = self.empty((n, n), torch.float64)
        Linv = self.empty((n, n), torch.float64) if want_inverse else None
        alpha = self.empty((n,), torch.float64) if want_inverse else None
        stats = (C.c_double * 4)()
        info = C.c_int(0)
        self._check(self.lib.nbw_gp_factor(self.ctx, self._p(K), K.stride(0), n, float(lik), self._p(y),
                                           self._p(L), n, self._p(Linv), n, self._p(alpha), stats,
                                           C.byref(info), self.stream))
        return L, Linv, alpha, float(stats[0]), float(stats[1]), float(stats[2]), float(stats[3])

    def level_grams_i32(self, fit: WLFit, hmax: int) -> torch.Tensor:
        """int32 [hmax + 1, n, ld]: the Gram of every single WL level (vertex_histogram.py:243 per level)."""
        n = fit.n_cells
        ld = _round_up(n, 4)
        G = self.empty((hmax + 1, n, ld), torch.int32)
        if n <= 512 and not os.environ.get("NBW_LEVEL_GRAMS_TC"):
            # small training sets: all levels in one SIMT launch (nbw_level_grams_i32)
            co = torch.tensor(fit.col_off[:hmax + 2], dtype=torch.int32).to(self.device, non_blocking=True)
            self._check(self.lib.nbw_level_grams_i32(self.ctx, self._p(fit.phi), n, fit.phi.stride(0), self._p(co), hmax + 1,
                                                     self._p(G), ld, n * ld, self.stream))
            return G
        for l in range(hmax + 1):
            self.gram(fit.phi, fit.phi, fit.col_off[l], fit.col_off[l + 1], out=G[l])
        return G

    def gp_grid(self, G: torch.Tensor, n: int, n_levels: Sequence[int], lik: float, y: torch.Tensor):
        """nbw_gp_grid: every candidate depth of the grid search factored in ONE launch.  Returns
        [(logdet, yKy, info)] per candidate; info != 0 marks a candidate whose factorisation broke down."""
        m = len(n_levels)
        nl = (C.c_int * m)(*[int(v) for v in n_levels])
        stats = (C.c_double * (2 * m))()
        info = (C.c_int * m)()
        self._check(self.lib.nbw_gp_grid(self.ctx, self._p(G), G.stride(1), G.stride(0), n, nl, m, float(lik), self._p(y),
                                         stats, info, self.stream))
        return [(float(stats[2 * c]), float(stats[2 * c + 1]), int(info[c])) for c in range(m)]

    def pd_inverse(self, K: torch.Tensor, jitter: float, y: torch.Tensor, want_inverse: bool = True):
        """compute_pd_inverse with its jitter retry (x10, x100) — bayesopt/utils.py:120-140."""
        last = None
        for fail in range(3):
            try:
                return self.gp_factor(K, jitter * (10 ** fail), y, want_inverse)
            except _lib.NotPositiveDefinite as e:
                last = e
        raise RuntimeError("Gram matrix not positive definite despite of jitter") from last

    def kernel_weight_grad(self, Linv: torch.Tensor, alpha: torch.Tensor, N: torch.Tensor, s: torch.Tensor,
                           raw_grams: Sequence[torch.Tensor]) -> List[float]:
        """d nlml'/d w_i for K = normalise(sum_i w_i R_i) + lik I (nbw_kernel_weight_grad)."""
        n = N.shape[0]
        Kinv = self.empty((n, n), torch.float64)
        self.gemm(True, False, n, n, n, 1.0, Linv, n, Linv, n, 0.0, Kinv, n)
        ptrs = (C.c_void_p * len(raw_grams))(*[C.c_void_p(R.data_ptr()) for R in raw_grams])
        grad = (C.c_double * len(raw_grams))()
        self._check(self.lib.nbw_kernel_weight_grad(self.ctx, self._p(Kinv), n, self._p(alpha), self._p(N),
                                                    N.stride(0), self._p(s), n, ptrs, raw_grams[0].stride(0),
                                                    len(raw_grams), grad, self.stream))
        return [float(g) for g in grad]

    def scale_features(self, phi: torch.Tensor, k0: int, k1: int, s: torch.Tensor, sqrt_w: float,
                       out: torch.Tensor, ldo: int):
        n = phi.shape[0]
        self._check(self.lib.nbw_scale_features(self.ctx, self._p(phi), phi.stride(0), n, k0, k1, self._p(s),
                                                float(sqrt_w), self._p(out), ldo, self.stream))

    def combine_columns(self, B: torch.Tensor, idx: torch.Tensor, out: torch.Tensor, ldo: int):
        """out[:, j] = sum_k +-B[:, idx[j, k]] (nbw_combine_columns; ~c subtracts column c, 0x7FFFFFFF skips)."""
        J, K = idx.shape
        self._check(self.lib.nbw_combine_columns(self.ctx, self._p(B), B.shape[0], B.stride(0), self._p(idx), J, K,
                                                 self._p(out), ldo, self.stream))
        return out

    def prefix_features(self, B: torch.Tensor, fit: "WLFit", h: int, out: Optional[torch.Tensor] = None,
                        ldp: Optional[int] = None) -> torch.Tensor:
        """Bp [n, T+1]: features summed over ancestor chains (nbw_prefix_features); the extra
        last column stays zero.  `out` (with leading dimension ldp) places the T columns of this kernel
        inside the side-by-side matrix of a sum of kernels."""
        n = B.shape[0]
        T = fit.label_off[h + 1]
        if out is None:
            out = torch.zeros((n, T + 1), dtype=torch.float64, device=self.device)
            ldp = T + 1
        co = (C.c_int64 * (h + 2))(*fit.col_off[:h + 2])
        lo = (C.c_int64 * (h + 2))(*fit.label_off[:h + 2])
        self._check(self.lib.nbw_prefix_features(self.ctx, self._p(B), B.stride(0), n, h + 1, co, lo,
                                                 self._p(fit.parents), C.c_void_p(out.data_ptr()), ldp, self.stream))
        return out

    def posterior_cov_finish(self, Kss: torch.Tensor, ldk: int, Q: torch.Tensor, m: int, lik: float,
                             y_std: float) -> torch.Tensor:
        cov = self.empty((m, m), torch.float64)
        self._check(self.lib.nbw_posterior_cov_finish(self.ctx, self._p(Kss), ldk, self._p(Q), Q.stride(0), m,
                                                      float(lik), float(y_std), self._p(cov), m, self.stream))
        return cov

    # ------------------------------------------------------------------ lookup tables / model
    def build_tables(self, fit: WLFit, h: int):
        """Per-level open-addressing tables for pool-side lookups (levels 1..h) and the
        level-0 op-code table."""
        tables, masks = [None], [0]
        for l in range(1, h + 1):
            d = fit.dims[l]
            cap = _pow2_at_least(max(2 * d, 16))
            t = self.empty((cap * 16,), torch.uint8)
            self._check(self.lib.nbw_model_build_table(self.ctx, self._p(fit.uniq_keys[l]),
                                                       self._p(fit.uniq_chks[l]), d, self._p(t), cap,
                                                       self.stream))
            tables.append(t)
            masks.append(cap - 1)
        l0 = np.full(256, 0xFFFF, dtype=np.uint16)
        for rank, code in enumerate(fit.op_codes):
            l0[int(code)] = rank
        level0 = torch.from_numpy(l0.view(np.int16)).to(self.device)
        return level0, tables, masks

    def build_packed_tables(self, fit: WLFit, last_hashed: int):
        """Packed-path dictionary forms for the hashed levels 1..last_hashed (nbw_model.id_hash / bucket_*):
        per level the hashes of the previous level's ids and the tag buckets of this level's keys.  Buckets
        hold four entries; n_buckets >= 4 D keeps an overflow (five keys in one bucket) a < 1 % event, which
        is answered by doubling."""
        n = last_hashed + 1
        id_hash, tags, entries, masks = [None] * n, [None] * n, [None] * n, [0] * n
        if last_hashed < 1:
            return id_hash, tags, entries, masks
        flags = torch.zeros((n,), dtype=torch.int32, device=self.device)
        seed = C.c_uint64(fit.seed & 0xFFFFFFFFFFFFFFFF)

        def build(l, nb):
            d = fit.dims[l]
            tags[l] = self.empty((nb * 16,), torch.uint8)
            entries[l] = self.empty((nb * 64,), torch.uint8)
            masks[l] = nb - 1
            self._check(self.lib.nbw_model_build_buckets(self.ctx, self._p(fit.uniq_keys[l]), self._p(fit.uniq_chks[l]), d,
                                                         self._p(tags[l]), self._p(entries[l]), nb, self._p(flags[l:]),
                                                         self.stream))
        for l in range(1, n):
            dp = max(fit.dims[l - 1], 1)
            id_hash[l] = self.empty((dp * 16,), torch.uint8)
            self._check(self.lib.nbw_model_build_id_hash(self.ctx, self._p(fit.uniq_keys[l - 1]), fit.dims[l - 1], seed, l,
                                                         self._p(id_hash[l]), self.stream))
            build(l, _pow2_at_least(max(4 * fit.dims[l], 16)))
        for attempt in range(8):
            bad = np.nonzero(flags.cpu().numpy())[0]
            if len(bad) == 0:
                return id_hash, tags, entries, masks
            flags.zero_()
            for l in bad:
                build(int(l), 2 * (masks[int(l)] + 1))
        raise RuntimeError("packed dictionary buckets keep overflowing")

    class _Range:
        def __init__(self, lib, name):
            self.lib, self.name = lib, name.encode()

        def __enter__(self):
            self.lib.nbw_range_push(self.name)

        def __exit__(self, *exc):
            self.lib.nbw_range_pop()

    def profile_events(self, ev_start=None, ev_stop=None):
        """nbw_profile_events: torch.cuda.Event pair recorded around the next scoring kernel launches."""
        a = None if ev_start is None else C.c_void_p(ev_start.cuda_event)
        b = None if ev_stop is None else C.c_void_p(ev_stop.cuda_event)
        self._check(self.lib.nbw_profile_events(self.ctx, a, b))

    def range(self, name: str):
        """NVTX range (nbw_range_push / nbw_range_pop) around a host-layer phase."""
        return Engine._Range(self.lib, name)

    # ------------------------------------------------------------------ pool scoring
    def score_pool_topk(self, model: "_lib.Model", cells: torch.Tensor, n_cells: int, k: int, distinct: bool = True,
                        index_base: int = 0, scores: Optional[torch.Tensor] = None,
                        out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Fused scoring + local top-k (nbw_score_pool_topk): returns the device list uint8 [k * 16]."""
        if out is None:
            out = self.empty((k * 16,), torch.uint8)
        self._check(self.lib.nbw_score_pool_topk(self.ctx, C.byref(model), self._p(cells), n_cells, self._p(scores), k,
                                                 int(distinct), index_base, self._p(out), self.stream))
        return out

    def fused_topk_ok(self, model: "_lib.Model", k: int) -> bool:
        """The packed kernel serves this model and k (else: nbw_score_pool + nbw_topk*)."""
        import os
        if k > 8 or not model.H or os.environ.get("NBW_SCORE_LEGACY"):
            return False
        if model.oa and (not model.oa_exc or model.next):
            return False
        hashed1 = model.h >= 1 and (model.first_stable == 0 or model.first_stable > 1)
        return (not hashed1) or bool(model.bucket_tags[1])

    def score_pool(self, model: "_lib.Model", cells: torch.Tensor, n_cells: int,
                   scores: Optional[torch.Tensor] = None, want64: bool = False):
        if scores is None:
            scores = self.empty((n_cells,), torch.float32)
        mu = var = sc = None
        if want64:
            mu = self.empty((n_cells,), torch.float64)
            var = self.empty((n_cells,), torch.float64)
            sc = self.empty((n_cells,), torch.float64)
        self._check(self.lib.nbw_score_pool(self.ctx, C.byref(model), self._p(cells), n_cells, self._p(scores),
                                            self._p(mu), self._p(var), self._p(sc), self.stream))
        return scores, mu, var, sc

    def score_host_pool(self, model: "_lib.Model", host_cells: torch.Tensor, chunks: int = 8,
                        mapped: bool = False):
        """Score a pool that lives in HOST memory through nbw_score_pool_host: the H2D copy of
        chunk i+1, the kernel on chunk i and the D2H copy of chunk i-1's scores overlap.  Returns
        (scores on the device [M] f32, scores in pinned host memory [M] f32); the host copy is
        complete after the next topk / topk_merge or wait_host_scores() (its D2H copy runs beside
        the selection).  `mapped=True` takes
        nbw_score_pool_mapped instead (pinned memory only): the kernel reads the records over PCIe
        itself and `chunks` is the number of kernel pieces."""
        assert not host_cells.is_cuda and host_cells.dtype == torch.uint8 and host_cells.dim() == 2
        host_cells = host_cells.contiguous()
        M = host_cells.shape[0]
        st = self._host_state = getattr(self, "_host_state", None) or {}
        if st.get("M") != M:
            st.update(M=M, scores=self.empty((M,), torch.float32))
        # a fresh pinned block per call: the caller owns the returned HOST scores (the device copy is a
        # work buffer, valid until the next host-pool call).  The D2H copy runs on the library's side
        # stream, which torch's caching host allocator does not know: the engine keeps the block (and the
        # records it reads) referenced until the copy has been awaited, so it cannot be recycled mid-flight.
        if st.get("pending") is not None:
            self.wait_host_scores()
        host_scores = torch.empty((M,), dtype=torch.float32, pin_memory=True)
        st["pending"] = (host_scores, host_cells)
        fn = self.lib.nbw_score_pool_mapped if mapped else self.lib.nbw_score_pool_host
        self._check(fn(self.ctx, C.byref(model), C.c_void_p(host_cells.data_ptr()), M,
                       C.c_void_p(host_scores.data_ptr()), self._p(st["scores"]), max(1, min(chunks, M)), self.stream))
        return st["scores"], host_scores

    def propose_host(self, model: "_lib.Model", host_cells: torch.Tensor, k: int, distinct: bool = True,
                     index_base: int = 0, chunks: int = 8, mapped: int = 0, out: Optional[torch.Tensor] = None):
        """nbw_propose_host: score a HOST pool (staged or mapped pipeline) with the top-k fused into the
        scoring kernels.  Returns (device list uint8 [k * 16], host scores [M] f32 pinned); the host scores
        are complete after cands_read() / topk_merge() / wait_host_scores()."""
        assert not host_cells.is_cuda and host_cells.dtype == torch.uint8 and host_cells.dim() == 2
        host_cells = host_cells.contiguous()
        M = host_cells.shape[0]
        st = self._host_state = getattr(self, "_host_state", None) or {}
        if st.get("M") != M:
            st.update(M=M, scores=self.empty((M,), torch.float32))
        if st.get("pending") is not None:
            self.wait_host_scores()
        host_scores = torch.empty((M,), dtype=torch.float32, pin_memory=True)
        st["pending"] = (host_scores, host_cells)
        if out is None:
            out = self.empty((k * 16,), torch.uint8)
        self._check(self.lib.nbw_propose_host(self.ctx, C.byref(model), C.c_void_p(host_cells.data_ptr()), M,
                                              C.c_void_p(host_scores.data_ptr()), self._p(st["scores"]), int(mapped),
                                              max(1, min(chunks, M)), k, int(distinct), index_base, self._p(out),
                                              self.stream))
        return out, host_scores

    def cands_read(self, cands: torch.Tensor, k: int):
        """Device top-k list -> (values, indices) on the host (nbw_cands_read; syncs)."""
        vals = (C.c_float * k)()
        idx = (C.c_int64 * k)()
        found = C.c_int(0)
        self._check(self.lib.nbw_cands_read(self.ctx, self._p(cands), k, vals, idx, C.byref(found), self.stream))
        f = found.value
        self.host_scores_done()
        return np.array(vals[:f], dtype=np.float32), np.array(idx[:f], dtype=np.int64)

    def wait_host_scores(self):
        self._check(self.lib.nbw_wait_host_scores(self.ctx))
        if getattr(self, "_host_state", None):
            self._host_state["pending"] = None

    def host_scores_done(self):
        """Called after an entry point that awaited the host scores itself (nbw_topk / nbw_topk_merge)."""
        if getattr(self, "_host_state", None):
            self._host_state["pending"] = None

    def pool_features(self, model: "_lib.Model", cells: torch.Tensor, n_cells: int):
        ld = _round_up(model.n_features, 16)
        phi = self.empty((n_cells, ld), torch.int8)
        ss = self.empty((n_cells,), torch.int32)
        self._check(self.lib.nbw_pool_features(self.ctx, C.byref(model), self._p(cells), n_cells, self._p(phi),
                                               ld, self._p(ss), self.stream))
        return phi, ss

    def topk(self, scores: torch.Tensor, k: int, distinct: bool = True, index_base: int = 0):
        n = scores.numel()
        vals = (C.c_float * k)()
        idx = (C.c_int64 * k)()
        found = C.c_int(0)
        self._check(self.lib.nbw_topk(self.ctx, self._p(scores), n, k, int(distinct), index_base, vals, idx,
                                      C.byref(found), self.stream))
        f = found.value
        self.host_scores_done()
        return np.array(vals[:f], dtype=np.float32), np.array(idx[:f], dtype=np.int64)

    def topk_device(self, scores: torch.Tensor, k: int, distinct: bool = True, index_base: int = 0,
                    out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Local top-k left on the device: uint8 [k * 16] ({f32 value, i32 pad, i64 index} entries)."""
        if out is None:
            out = self.empty((k * 16,), torch.uint8)
        self._check(self.lib.nbw_topk_device(self.ctx, self._p(scores), scores.numel(), k, int(distinct),
                                             index_base, self._p(out), self.stream))
        return out

    def topk_merge_device(self, cands: torch.Tensor, k: int, distinct: bool = True,
                          out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """nbw_topk_merge_device: merge device lists into one device list uint8 [k * 16]; no sync."""
        if out is None:
            out = self.empty((k * 16,), torch.uint8)
        self._check(self.lib.nbw_topk_merge_device(self.ctx, self._p(cands), cands.numel() // 16, k, int(distinct),
                                                   self._p(out), self.stream))
        return out

    def topk_merge(self, cands: torch.Tensor, k: int, distinct: bool = True):
        count = cands.numel() // 16
        vals = (C.c_float * k)()
        idx = (C.c_int64 * k)()
        found = C.c_int(0)
        self._check(self.lib.nbw_topk_merge(self.ctx, self._p(cands), count, k, int(distinct), vals, idx,
                                            C.byref(found), self.stream))
        f = found.value
        self.host_scores_done()
        return np.array(vals[:f], dtype=np.float32), np.array(idx[:f], dtype=np.int64)

    def generate_cells(self, family: int, seed: int, first_index: int, n_cells: int) -> torch.Tensor:
        stride = 48 if family == 2 else 16
        cells = self.empty((n_cells, stride), torch.uint8)
        self._check(self.lib.nbw_generate_cells(self.ctx, family, C.c_uint64(seed), first_index, n_cells,
                                                self._p(cells), self.stream))
        return cells

    def mutate_cells(self, family: int, parents: torch.Tensor, seed: int, first_index: int, n_children: int) -> torch.Tensor:
        """BANANAS-style mutation pool on the device (nbw_mutate_cells); parents are device records."""
        cells = self.empty((n_children, 16), torch.uint8)
        self._check(self.lib.nbw_mutate_cells(self.ctx, family, self._p(parents), parents.shape[0], C.c_uint64(seed),
                                              first_index, n_children, self._p(cells), self.stream))
        return cells

    def pool_dedup(self, cells: torch.Tensor, n_max: int, labeled: bool = True, n_protected: int = 0,
                   seed: int = DEFAULT_SEED):
        """nbw_pool_dedup: keep-first de-duplication of device records by WL isomorphism signature.
        Returns (keep mask uint8 [n] on the device, number of classes, number kept beyond the protected prefix)."""
        n = cells.shape[0]
        keep = self.empty((n,), torch.uint8)
        nc, nk = C.c_int64(0), C.c_int64(0)
        for attempt in range(4):
            try:
                self._check(self.lib.nbw_pool_dedup(self.ctx, self._p(cells), n, n_max, int(labeled), int(n_protected),
                                                    C.c_uint64((seed + attempt * 0x9E3779B1) & 0xFFFFFFFFFFFFFFFF),
                                                    self._p(keep), C.byref(nc), C.byref(nk), self.stream))
                return keep, int(nc.value), int(nk.value)
            except _lib.HashCollision:
                continue
        raise RuntimeError("cell signatures collided for 4 different seeds")

    def generate_encoded(self, family: int, seed: int, first_index: int, n_cells: int):
        """(device records, device encodings uint8 [n, 8 | 16]) of random NB201 / DARTS cells
        (nbw_generate_encoded): the same cells as generate_cells plus what a later mutation needs."""
        stride, eb = (48, 16) if family == 2 else (16, 8)
        cells = self.empty((n_cells, stride), torch.uint8)
        enc = self.empty((n_cells, eb), torch.uint8)
        self._check(self.lib.nbw_generate_encoded(self.ctx, family, C.c_uint64(seed), first_index, n_cells,
                                                  self._p(cells), self._p(enc), self.stream))
        return cells, enc

    def mutate_encoded(self, family: int, parent_enc: torch.Tensor, seed: int, first_index: int, n_children: int,
                       edits: int = 1):
        """BANANAS-style mutation of NB201 / DARTS architectures on their encodings
        (nbw_mutate_encoded).  Returns (device records, device child encodings)."""
        stride, eb = (48, 16) if family == 2 else (16, 8)
        assert parent_enc.is_cuda and parent_enc.dtype == torch.uint8 and parent_enc.shape[1] == eb
        parent_enc = parent_enc.contiguous()
        cells = self.empty((n_children, stride), torch.uint8)
        enc = self.empty((n_children, eb), torch.uint8)
        self._check(self.lib.nbw_mutate_encoded(self.ctx, family, self._p(parent_enc), parent_enc.shape[0],
                                                C.c_uint64(seed), first_index, n_children, int(edits),
                                                self._p(cells), self._p(enc), self.stream))
        return cells, enc
