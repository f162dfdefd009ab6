"""Base class of graph kernels (reference kernels/graph_kernel.py:5-27)."""
import torch


class GraphKernels:
    def __init__(self, **kwargs):
        self.n_hyperparameters = 0
        self.rbf_lengthscale = False
        self.kern = None
        self.__name__ = 'GraphKernelBase'

    def normalize_gram(self, K: torch.Tensor):
        K_diag = torch.sqrt(torch.diag(K))
        return K / torch.outer(K_diag, K_diag)

    def fit_transform(self, gr: list, rebuild_model=False, save_gram_matrix=False, **kwargs):
        raise NotImplementedError

    def transform(self, gr: list, ):
        raise NotImplementedError

    def forward_t(self, gr2, gr1: list = None):
        raise NotImplementedError("The kernel gradient is not implemented for the graph kernel called!")
