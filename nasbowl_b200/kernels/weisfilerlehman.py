"""Weisfeiler-Lehman subtree kernel on the GPU, with the reference's interface
(kernels/weisfilerlehman.py:15-152).

`fit_transform` returns, like the reference, the UN-normalised integer Gram
K_raw = sum_{l<=h} Phi_l Phi_l^T (grakel_replace/weisfeiler_lehman.py:316-327,353-362) as a
CPU float64 tensor; normalisation is the job of SumKernel (combine_kernels.py:54-56).
Relabelling, dictionary build, histograms and the int8 tensor-core Gram all run through the
C-ABI; this class only packs networkx graphs into cell records and keeps the fitted state.
"""
from __future__ import annotations

import logging
from typing import List, Optional

import numpy as np
import torch

from ..cells import CellBatch, OpVocab, pack_networkx
from .base import GraphKernels

GLOBAL_VOCAB = OpVocab()


class _KernState:
    """What the reference keeps in `self.kern` (the GraKeL object): enough of it for callers
    that peek (`kern._h`, `kern.layer_weights`)."""

    def __init__(self, h, layer_weights):
        self.h = h
        self._h = h + 1
        self.layer_weights = layer_weights


class WeisfilerLehman(GraphKernels):
    def __init__(self, h: int = 0, type='subtree', se_kernel=None, layer_weights=None, node_weights=None,
                 oa=False, node_label: str = 'op_name', edge_label: tuple = 'op_name', n_jobs: int = None,
                 return_tensor: bool = True, requires_grad: bool = False, undirected: bool = False, slot: int = 0,
                 **kwargs):
        """slot (extension): which cell of a multi-cell example this kernel reads when the GP is given
        tuples of graphs / CellSlots ("normal | reduce"); 0 = the reference's behaviour."""
        super().__init__(**kwargs)
        self.slot = int(slot)
        if se_kernel is not None and oa:
            raise ValueError("Only one or none of se (successive embedding) and oa (optimal assignment) may be true!")
        assert type in ['subtree', 'sp', 'edge']
        if type != 'subtree':
            # 'sp' and 'edge' are documented as untested in the reference (weisfilerlehman.py:36-37)
            raise NotImplementedError("only the 'subtree' WL base kernel is implemented on the GPU path")
        if se_kernel is not None:
            raise NotImplementedError("successive-embedding (se_kernel) is outside the GPU hot path")
        if not isinstance(h, (int, np.integer)) or h < 0:
            raise TypeError("'h' must be a non-negative integer. Got h" + str(h))
        self.h = int(h)
        self.oa = bool(oa)
        self.node_label = node_label
        self.edge_label = edge_label
        self.layer_weights = layer_weights
        self.node_weights = node_weights          # stored, never used
        self.se = None
        self.requires_grad = requires_grad
        self.undirected = undirected
        self.type = type
        self.return_tensor = return_tensor
        self.vocab = GLOBAL_VOCAB
        self.kern = _KernState(self.h, layer_weights)
        self._gram = None
        self._train, self._train_transformed = None, None
        self._fit = None                          # engine.WLFit of the last fit (device state)
        self._fit_batch = None                    # the packed cells it was fitted on (host)
        self._ref_cache = None
        self.__name__ = 'WeisfeilerLehman'

    # ------------------------------------------------------------------ reference API
    def change_se_params(self, params: dict):
        logging.warning("SE kernel is None. change_se_params action voided.")

    def change_kernel_params(self, params: dict):
        for k, v in params.items():
            if k == 'h':
                if not isinstance(v, (int, np.integer)) or v < 0:
                    raise TypeError("'h' must be a non-negative integer. Got h" + str(v))
                self.h = int(v)
            elif k == 'layer_weights':
                self.layer_weights = v
            elif hasattr(self, k):
                setattr(self, k, v)
            else:
                logging.warning(str(k) + " is not a valid attribute name of this kernel.")
        self.kern = _KernState(self.h, self.layer_weights)

    def pack(self, gr) -> CellBatch:
        """list[nx.DiGraph] | CellBatch -> CellBatch (graph_from_networkx call sites,
        weisfilerlehman.py:134-136)."""
        if isinstance(gr, CellBatch):
            return gr.to_undirected() if self.undirected else gr
        return pack_networkx(gr, self.vocab, node_label=self.node_label, undirected=self.undirected)

    def _level_weights(self, h: int) -> Optional[np.ndarray]:
        lw = self.layer_weights
        if lw is None:
            return None
        lw = np.asarray(lw.detach().cpu() if isinstance(lw, torch.Tensor) else lw, dtype=np.float64).reshape(-1)
        if lw.shape[0] != h + 1:                  # weisfeiler_lehman.py:126-127: reset to ones
            return None
        return None if np.all(lw == 1.0) else lw

    def fit_device(self, cells: torch.Tensor, n_max: int, h: Optional[int] = None, batch: Optional[CellBatch] = None):
        """Fit on packed device records; returns the engine's WLFit (features stay on device)."""
        from ..engine import Engine
        eng = Engine.get()
        self._fit = eng.wl_fit(cells, n_max, self.h if h is None else h, self.oa)
        self._fit_batch = batch
        return self._fit

    def gram_device(self, fit, h: Optional[int] = None) -> torch.Tensor:
        """fp64 device K_raw(h) from a WLFit that covers at least depth h."""
        from ..engine import Engine
        eng = Engine.get()
        h = self.h if h is None else h
        n = fit.n_cells
        out = eng.empty((n, n), torch.float64)
        lw = self._level_weights(h)
        if lw is None:
            Ki = eng.gram(fit.phi, fit.phi, 0, fit.k_end(h))
            eng.gram_accumulate(Ki, 1.0, 0.0, out)
        else:
            for l in range(h + 1):
                Ki = eng.gram(fit.phi, fit.phi, fit.col_off[l], fit.col_off[l + 1])
                eng.gram_accumulate(Ki, float(lw[l]), 0.0 if l == 0 else 1.0, out)
        return out

    def level_grams_device(self, fit, h: Optional[int] = None) -> List[torch.Tensor]:
        """fp64 device Grams of the single WL levels 0..h (K_raw = sum_l w_l K_l)."""
        from ..engine import Engine
        eng = Engine.get()
        h = self.h if h is None else h
        n = fit.n_cells
        out = []
        for l in range(h + 1):
            Kl = eng.empty((n, n), torch.float64)
            eng.gram_accumulate(eng.gram(fit.phi, fit.phi, fit.col_off[l], fit.col_off[l + 1]), 1.0, 0.0, Kl)
            out.append(Kl)
        return out

    def fit_transform(self, gr: list, rebuild_model=False, save_gram_matrix=True, layer_weights=None, **kwargs):
        if rebuild_model is False and self._gram is not None:
            return self._gram
        from ..engine import Engine
        eng = Engine.get()
        batch = self.pack(gr)
        if len(batch) == 0:
            raise ValueError('parsed input is empty')
        if rebuild_model or self._gram is None:
            self._train = gr[:] if not isinstance(gr, CellBatch) else gr
            self._train_transformed = batch
        if layer_weights is not None and layer_weights is not self.layer_weights:
            self.change_kernel_params({'layer_weights': layer_weights})
        fit = self.fit_device(eng.upload_cells(batch), batch.n_max, batch=batch)
        K = self.gram_device(fit).cpu()
        if not self.return_tensor:
            K = K.numpy()
        if save_gram_matrix:
            self._gram = K.clone() if isinstance(K, torch.Tensor) else K.copy()
            self.layer_weights = self.kern.layer_weights
        return K

    def transform(self, gr: list):
        """Embeddings of `gr` in the fitted vocabulary, one array per WL level
        (grakel_replace/weisfeiler_lehman.py:364-516 with return_embedding_only=True)."""
        phi, _ = self.feature_matrix(gr)
        fit = self._fit
        out = [phi[:, fit.col_off[l]:fit.col_off[l] + fit.widths[l]] for l in range(self.h + 1)]
        return [torch.tensor(o) for o in out] if self.return_tensor else out

    def _linear_fit(self):
        """The fitted vocabulary with the plain histogram layout (one column per label).  For an
        oa kernel the device features are thermometer-coded and capped at the training counts; the
        reference's embeddings (kern.transform(return_embedding_only=True)) are plain counts, so the
        interpretability outputs are built from this second layout of the same label ids."""
        fit = self._fit
        if fit is None or not fit.oa:
            return fit
        if getattr(self, '_lin_cache', None) is not None and self._lin_cache[0] is fit:
            return self._lin_cache[1]
        import dataclasses
        from ..engine import Engine, _round_up
        col_off = [0]
        for d in fit.dims:
            col_off.append(col_off[-1] + _round_up(max(d, 1), 128))
        lin = dataclasses.replace(fit, oa=False, widths=list(fit.dims), col_off=col_off, thermo_base=None,
                                  max_count=None, phi=None, self_sim=None)
        Engine.get().build_features(lin)
        self._lin_cache = (fit, lin)
        return lin

    def feature_matrix(self, gr, linear=False):
        """(Phi restricted to the fitted vocabulary [M, D] int8 numpy, self-similarity [M]).
        linear=True: plain per-label counts even when the kernel is oa."""
        if self._fit is None:
            raise ValueError("The kernel has not been fitted. Call fit_transform first.")
        from ..engine import Engine
        from ..pool_model import PoolModel
        eng = Engine.get()
        fit = self._linear_fit() if linear else self._fit
        batch = self.pack(gr).widen(fit.n_max)
        pm = PoolModel.from_fit(eng, fit, self.h)
        phi, ss = eng.pool_features(pm.descriptor(), eng.upload_cells(batch), len(batch))
        return phi.cpu().numpy()[:, :pm.n_features], ss.cpu().numpy()

    def embedding(self, gr=None, pad_levels: bool = False) -> np.ndarray:
        """Histogram embedding [N, sum_l D_l] float64, columns in the reference's feature order
        (feature_map order), restricted to the fitted vocabulary.  gr=None: the training cells.
        pad_levels: one all-zero spare column after each level's block — the shape of the
        reference's ordered feature matrices (vertex_histogram.py:184) that forward_t exposes."""
        ref = self._reference_labels()
        if ref is None:
            raise ValueError("The kernel has not been fitted. Call fit_transform first.")
        ref_id, _ = ref
        lin = self._linear_fit()
        if gr is None:
            phi = lin.phi.cpu().numpy()
        else:
            phi, _ = self.feature_matrix(gr, linear=True)
        total = sum(lin.dims[:self.h + 1])
        emb = np.zeros((phi.shape[0], total), dtype=np.float64)
        for l in range(self.h + 1):
            emb[:, ref_id[l]] = phi[:, lin.col_off[l]:lin.col_off[l] + lin.dims[l]]
        if pad_levels:
            ends = np.cumsum(lin.dims[:self.h + 1])
            spare = np.zeros((emb.shape[0], 1))
            emb = np.concatenate([np.concatenate([emb[:, e - d:e], spare], axis=1)
                                  for e, d in zip(ends, lin.dims[:self.h + 1])], axis=1)
        return emb

    def forward_t(self, gr2, gr1=None):
        """(K, y_, x_): the cosine-normalised LINEAR kernel between gr1 (default: the training cells)
        and gr2 as a torch graph over the leaf embeddings x_ (gr1) and y_ (gr2), for Jacobian-vector
        products (reference kernels/weisfilerlehman.py:172-217 + grakel_replace/utils.py:20-71; like
        the reference, the tensor form is linear even when the kernel is oa).  The embeddings come
        from the device; the small differentiable product is host-side torch.
        Deviation: labels of gr2 outside the training vocabulary are dropped; the reference keeps each
        level's first unknown label in the spare column and, for a cell with two or more unknown labels
        at one level, shifts the deeper levels (it truncates after concatenating) — DESIGN.md §8."""
        x_ = torch.tensor(self.embedding(None if gr1 is None else gr1, pad_levels=True)).requires_grad_()
        y_ = torch.tensor(self.embedding(gr2, pad_levels=True)).requires_grad_()
        K = y_ @ x_.t()
        K = K / torch.ger(torch.sqrt(torch.diag(y_ @ y_.t())), torch.sqrt(torch.diag(x_ @ x_.t())))
        return K, y_, x_

    # ------------------------------------------------------------------ motif strings (row f3)
    def _reference_labels(self):
        """Number the device's label classes exactly like the reference and name them.

        The reference numbers the labels of level l by the rank of their credential STRING
        `str(own) + "," + str(sorted(successor ids))` among the level's distinct credentials,
        continuing one counter over the levels (weisfeiler_lehman.py:226-229,266-274), and names
        them "root~leaf~leaf" in terms of the previous level's names (translate_label, :577-599).
        Only the partition is computed on the GPU; this assembles the strings on the host from
        one representative node per label class.  Returns (ref_id, names): per level, an array
        device id -> reference id, and a dict reference id -> motif string."""
        fit, batch = self._fit, self._fit_batch
        if fit is None or batch is None:
            return None
        if getattr(self, '_ref_cache', None) is not None and self._ref_cache[0] is fit:
            return self._ref_cache[1]
        h, n_max = self.h, fit.n_max
        ids = fit.ids[:h + 1].cpu().numpy().view(np.uint32).reshape(h + 1, fit.n_cells, n_max)
        succ = batch.succ
        names0 = {}
        codes = [int(c) for c in fit.op_codes]
        order = sorted(codes, key=lambda c: self.vocab.names[c])
        ref_of_code = {c: i for i, c in enumerate(order)}
        ref_id = [np.array([ref_of_code[c] for c in codes], dtype=np.int64)]
        names = [{ref_of_code[c]: str(self.vocab.names[c]) for c in codes}]
        counter = len(codes)
        for l in range(1, h + 1):
            D = fit.dims[l]
            flat = ids[l].reshape(-1)
            valid = np.nonzero(flat != 0xFFFFFFFF)[0]
            rep = np.full(D, -1, dtype=np.int64)
            rep[flat[valid][::-1]] = valid[::-1]              # first node carrying each label
            creds = []
            for z in range(D):
                g, v = divmod(int(rep[z]), n_max)
                own = int(ref_id[l - 1][ids[l - 1][g, v]])
                nb = sorted(int(ref_id[l - 1][ids[l - 1][g, j]]) for j in range(n_max) if (int(succ[g, v]) >> j) & 1)
                creds.append((str(own) + "," + str(nb), own, nb))
            rank = sorted(range(D), key=lambda z: creds[z][0])
            rid = np.empty(D, dtype=np.int64)
            nm = {}
            for r, z in enumerate(rank):
                rid[z] = counter + r
                _, own, nb = creds[z]
                nm[counter + r] = "~".join([names[l - 1][own]] + [names[l - 1][i] for i in nb])
            counter += D
            ref_id.append(rid)
            names.append(nm)
        self._ref_cache = (fit, (ref_id, names))
        return ref_id, names

    def feature_map(self, flatten=True):
        """{feature index: motif string}, layered by WL level unless `flatten`
        (reference kernels/weisfilerlehman.py:219-242)."""
        if self._gram is None and self._fit is None:
            return None
        ref = self._reference_labels()
        if ref is None:
            return None
        _, names = ref
        if not flatten:
            return {l: dict(sorted(nm.items())) for l, nm in enumerate(names)}
        res = {}
        for nm in names:
            res.update(dict(sorted(nm.items())))
        return res

    def feature_value(self, X_s):
        """WL embedding of X_s over the fitted features, columns in feature_map order, plus the
        list of motif strings (reference kernels/weisfilerlehman.py:244-266)."""
        ref = self._reference_labels()
        if ref is None:
            raise ValueError("The kernel has not been fitted. Call fit_transform first.")
        emb = self.embedding(X_s)
        return torch.tensor(emb), list(self.feature_map(flatten=True).values())

    # device state does not travel through pickle (run_darts.py:110-125 pickles the GP)
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_fit'] = None
        st['_fit_batch'] = None
        st['_ref_cache'] = None
        st['_lin_cache'] = None
        return st
