"""TEST INFRASTRUCTURE — loads the *unmodified* reference from /root/reference under import shims.

Only usable in the build container (the reference tree does not travel to the GPU box).
Used by oracle/make_golden.py to generate tests/golden/*.npz and by the optional
`tests/test_oracle_vs_reference.py` differential tests on fresh seeds (skipped when the tree is absent).
"""
import collections
import collections.abc
import os
import sys
import warnings

REFERENCE_ROOT = os.environ.get("NASBOWL_REFERENCE", "/root/reference")
_SHIMS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_shims")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "grakel_replace"))


def load_reference():
    """Return a namespace with the reference's hot-path classes (imports are cached)."""
    if not reference_available():
        raise RuntimeError("reference tree not found at " + REFERENCE_ROOT)
    # py3.10+ removed the alias the vendored grakel code imports (vertex_histogram.py:4)
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    for p in (_SHIMS, REFERENCE_ROOT):
        if p not in sys.path:
            sys.path.insert(0, p)
    warnings.filterwarnings("ignore")
    import types
    import kernels as ref_kernels                      # noqa: E402
    import bayesopt as ref_bayesopt                    # noqa: E402
    from bayesopt import generate_test_graphs as gtg   # noqa: E402
    ns = types.SimpleNamespace(
        kernels=ref_kernels, bayesopt=ref_bayesopt, gtg=gtg,
        WeisfilerLehman=ref_kernels.WeisfilerLehman, SumKernel=ref_kernels.SumKernel,
        GraphGP=ref_bayesopt.GraphGP,
        GraphExpectedImprovement=ref_bayesopt.GraphExpectedImprovement,
        GraphUpperConfidentBound=ref_bayesopt.GraphUpperConfidentBound,
    )
    return ns


def load_reference_functions(rel_path: str, names, namespace=None):
    """Execute only the named top-level functions of one reference source file, unmodified, in
    `namespace` — for modules whose own imports drag in things this container lacks (e.g.
    darts/darts_objectives.py imports the NAS-Bench APIs and the CNN trainer).  Returns the namespace."""
    import ast
    path = os.path.join(REFERENCE_ROOT, rel_path)
    src = open(path).read()
    tree = ast.parse(src)
    ns = {} if namespace is None else namespace
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in names:
            code = compile(ast.Module(body=[node], type_ignores=[]), path, "exec")
            exec(code, ns)
    missing = [n for n in names if n not in ns]
    if missing:
        raise RuntimeError("not found in %s: %s" % (rel_path, missing))
    return ns
