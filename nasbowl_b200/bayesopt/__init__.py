"""GP surrogate and acquisition functions with the reference's class names."""
from . import acquisitions, gp

GraphGP = gp.GraphGP
GraphExpectedImprovement = acquisitions.GraphExpectedImprovement
GraphUpperConfidentBound = acquisitions.GraphUpperConfidentBound

__all__ = ["GraphGP", "GraphExpectedImprovement", "GraphUpperConfidentBound"]
