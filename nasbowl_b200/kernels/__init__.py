from .graph_kernel import GraphKernels  # noqa: F401
from .weisfilerlehman import WeisfilerLehman  # noqa: F401
from .combine_kernels import CombineKernel, SumKernel  # noqa: F401
