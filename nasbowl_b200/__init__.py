"""nasbowl_b200 — B200-native implementation of NAS-BOWL's surrogate-scoring hot path.

Drop-in for the reference's `kernels.WeisfilerLehman`, `kernels.SumKernel`,
`bayesopt.GraphGP`, `bayesopt.GraphExpectedImprovement`, `bayesopt.GraphUpperConfidentBound`
(same names, arguments and error behaviour); all arithmetic runs in hand-written sm_100a
CUDA kernels behind the C-ABI of include/nbw.h.  Importing the package is light; the first
compute call loads libnbw.so and raises if it is missing — there is no CPU fallback.
"""
from .cells import CellBatch, OpVocab, pack_networkx  # noqa: F401

__version__ = "0.1.0"
