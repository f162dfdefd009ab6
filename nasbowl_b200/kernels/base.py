"""Common parent of the graph kernels of this package.

`GraphGP` tells graph kernels from vectorial ones with ``isinstance(k, GraphKernels)`` (reference
bayesopt/gp.py:42-44), so the class name is part of the interface; everything numerical lives in
the subclasses and, ultimately, in libnbw."""
from __future__ import annotations

import torch


class GraphKernels:
    """Marker base class + the one helper callers use on it (cosine normalisation of a Gram)."""

    __name__ = "GraphKernelBase"
    n_hyperparameters = 0
    rbf_lengthscale = False

    def __init__(self, **_unused):
        self.kern = None

    @staticmethod
    def normalize_gram(K: torch.Tensor) -> torch.Tensor:
        """K_ij / sqrt(K_ii K_jj) for a Gram already on the host (small matrices only; the GPU
        path normalises on the device with nbw_gram_normalize)."""
        d = torch.diagonal(K).sqrt()
        return K / (d[:, None] * d[None, :])

    # the three entry points a concrete kernel provides
    def fit_transform(self, gr, rebuild_model=False, save_gram_matrix=False, **kwargs):
        raise NotImplementedError(type(self).__name__ + ".fit_transform")

    def transform(self, gr):
        raise NotImplementedError(type(self).__name__ + ".transform")

    def forward_t(self, gr2, gr1=None):
        raise NotImplementedError("kernel gradients are not implemented for " + type(self).__name__)
