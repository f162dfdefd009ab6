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
from .graph_kernel import GraphKernels

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
                 return_tensor: bool = True, requires_grad: bool = False, undirected: bool = False, **kwargs):
        super().__init__(**kwargs)
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

    def fit_device(self, cells: torch.Tensor, n_max: int, h: Optional[int] = None):
        """Fit on packed device records; returns the engine's WLFit (features stay on device)."""
        from ..engine import Engine
        eng = Engine.get()
        self._fit = eng.wl_fit(cells, n_max, self.h if h is None else h, self.oa)
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
        fit = self.fit_device(eng.upload_cells(batch), batch.n_max)
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

    def feature_matrix(self, gr):
        """(Phi restricted to the fitted vocabulary [M, D] int8 numpy, self-similarity [M])."""
        if self._fit is None:
            raise ValueError("The kernel has not been fitted. Call fit_transform first.")
        from ..engine import Engine
        from ..pool_model import PoolModel
        eng = Engine.get()
        batch = self.pack(gr).widen(self._fit.n_max)
        pm = PoolModel.from_fit(eng, self._fit, self.h)
        phi, ss = eng.pool_features(pm.descriptor(), eng.upload_cells(batch), len(batch))
        return phi.cpu().numpy()[:, :pm.n_features], ss.cpu().numpy()

    def forward_t(self, gr2, gr1=None):
        raise NotImplementedError("kernel gradients (interpretability, gp.py:296-424) are outside the GPU hot path")

    def feature_map(self, flatten=True):
        raise NotImplementedError("feature_map (motif strings) is not part of the scoring hot path yet")

    def feature_value(self, X_s):
        raise NotImplementedError("feature_value is not part of the scoring hot path yet")

    # device state does not travel through pickle (run_darts.py:110-125 pickles the GP)
    def __getstate__(self):
        st = dict(self.__dict__)
        st['_fit'] = None
        return st
