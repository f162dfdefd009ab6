"""Attribute-only stand-in: bayesopt/gp.py:14-24 defines an unused ExactGP subclass at import."""
import types

models = types.SimpleNamespace(ExactGP=object)
likelihoods = types.SimpleNamespace(GaussianLikelihood=object)
means = types.SimpleNamespace(ConstantMean=object)
kernels = types.SimpleNamespace(ScaleKernel=object)
distributions = types.SimpleNamespace(MultivariateNormal=object)
