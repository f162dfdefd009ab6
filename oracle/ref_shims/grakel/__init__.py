"""Container-only stand-in for grakel==0.1b7 (reference README.md:41), see ../README.md."""
from . import graph, kernels, utils  # noqa: F401
