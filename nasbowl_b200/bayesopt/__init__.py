from .acquisitions import GraphExpectedImprovement, GraphUpperConfidentBound  # noqa: F401
from .gp import GraphGP  # noqa: F401
