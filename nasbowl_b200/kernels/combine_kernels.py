"""Weighted sum of graph kernels followed by cosine normalisation, on the GPU
(reference kernels/combine_kernels.py:7-125)."""
from __future__ import annotations

import torch

from ..cells import CellBatch
from .base import GraphKernels


class CombineKernel:
    def __init__(self, combined_by='sum', *kernels, **kwargs):
        self.has_graph_kernels = False
        self.has_vector_kernels = False
        for k in kernels:
            if isinstance(k, GraphKernels):
                self.has_graph_kernels = True
            else:
                self.has_vector_kernels = True
        self.kernels = kernels
        self._gram = None
        self.gr, self.x = None, None
        assert combined_by in ['sum', 'multiply']
        if combined_by != 'sum':
            raise NotImplementedError("only the sum combination is implemented on the GPU path")
        if self.has_vector_kernels:
            raise NotImplementedError("vectorial (Stationary) kernels are outside the GPU hot path")
        self.combined_by = combined_by

    def combine_device(self, weights, grams, normalize=True):
        """K = sum_i weights_i * K_raw,i, then K /= sqrt(diag ⊗ diag) — all on device, fp64.
        grams: list of fp64 device matrices.  Returns (K, inv_sqrt_diag or None)."""
        from ..engine import Engine
        eng = Engine.get()
        n = grams[0].shape[0]
        K = eng.empty((n, n), torch.float64)
        for i, Ki in enumerate(grams):
            eng.axpby(Ki, float(weights[i]), 0.0 if i == 0 else 1.0, K)
        s = eng.gram_normalize(K) if normalize else None
        return K, s

    def fit_transform(self, weights, gr1: list, x1=None, feature_lengthscale=None, normalize=True,
                      rebuild_model=True, save_gram_matrix=True, **kwargs):
        if x1 is not None:
            raise NotImplementedError("vectorial features are outside the GPU hot path")
        from ..engine import Engine
        eng = Engine.get()
        grams = []
        layer_weights = kwargs.get('layer_weights', None)
        packed = {}
        for k in self.kernels:
            # every kernel converts the graphs itself (its own node_label / undirected view, like the
            # per-kernel graph_from_networkx of kernels/weisfilerlehman.py:127-136); equal views are shared
            key = (getattr(k, 'node_label', None), bool(getattr(k, 'undirected', False)))
            if key not in packed:
                b = k.pack(gr1)
                packed[key] = (b, eng.upload_cells(b))
            batch, cells = packed[key]
            if layer_weights is not None and layer_weights is not k.layer_weights:
                k.change_kernel_params({'layer_weights': layer_weights})
            if rebuild_model or k._fit is None:
                if rebuild_model:
                    k._train = gr1[:] if not isinstance(gr1, CellBatch) else gr1
                k.fit_device(cells, batch.n_max, batch=batch)
            grams.append(k.gram_device(k._fit))
        K, _ = self.combine_device(weights, grams, normalize)
        K = K.cpu()
        if save_gram_matrix:
            self._gram = K.clone()
        self.x = None
        return K

    def transform(self, weights, gr: list, x=None, feature_lengthscale=None):
        raise NotImplementedError("CombineKernel.transform is unused by the reference's GP (gp.py:268 "
                                  "always calls fit_transform on train ∪ test)")


class SumKernel(CombineKernel):
    def __init__(self, *kernels, **kwargs):
        super().__init__('sum', *kernels, **kwargs)

    def forward_t(self, weights, gr2: list, x2=None, gr1: list = None, x1=None, feature_lengthscale=None):
        raise NotImplementedError("kernel gradients (interpretability) are outside the GPU hot path")
