"""grakel.kernels.Kernel stand-in: the sklearn-style base the vendored kernels subclass."""
from sklearn.base import BaseEstimator, TransformerMixin


class Kernel(BaseEstimator, TransformerMixin):
    _graph_format = "dictionary"
    _method_calling = 0

    def __init__(self, n_jobs=None, normalize=False, verbose=False):
        self.n_jobs = n_jobs
        self.normalize = normalize
        self.verbose = verbose
        self._initialized = dict(n_jobs=False)

    def initialize(self):
        if not self._initialized["n_jobs"]:
            self._initialized["n_jobs"] = True

    def fit(self, X, y=None, **kwargs):
        self._is_transformed = False
        self._method_calling = 1
        self.initialize()
        if X is None:
            raise ValueError("`fit` input cannot be None")
        self.X = self.parse_input(X, **kwargs)
        return self


class ShortestPathAttr(Kernel):
    pass


class ShortestPath(Kernel):
    pass


class RandomWalkLabeled(Kernel):
    pass


class RandomWalk(Kernel):
    pass


class MultiscaleLaplacian(Kernel):
    pass
