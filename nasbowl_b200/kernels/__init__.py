"""Graph kernels with the reference's class names (kernels/__init__.py of xingchenwan/nasbowl)."""
from .base import GraphKernels
from .combine_kernels import CombineKernel, SumKernel
from .weisfilerlehman import WeisfilerLehman

__all__ = ["GraphKernels", "WeisfilerLehman", "CombineKernel", "SumKernel"]
