"""ORACLE support: import the reference's own ``discretekg`` module in the build container.

/root/reference is mounted only in the build container (never on the GPU box), and botorch is
absent, so the module is imported with ``oracle/_stubs`` providing the botorch names.  Used by
``oracle/make_golden.py`` and by ``tests/test_oracle_vs_reference.py`` (skipped when the
reference tree is not present).
"""
import importlib
import os
import sys

REFERENCE_SRC = "/root/reference/src"
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_stubs")


def reference_available() -> bool:
    return os.path.isfile(
        os.path.join(REFERENCE_SRC, "decoupledbo/modules/acquisition/discretekg.py")
    )


def load_reference_discretekg():
    """Returns the reference module ``decoupledbo.modules.acquisition.discretekg``."""
    if not reference_available():
        raise RuntimeError("reference tree not mounted")
    try:
        import botorch  # noqa: F401  (the real one, if it ever becomes available)
    except ImportError:
        if _STUBS not in sys.path:
            sys.path.insert(0, _STUBS)
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if repo_root not in sys.path:
        sys.path.insert(0, repo_root)
    if REFERENCE_SRC not in sys.path:
        sys.path.append(REFERENCE_SRC)
    return importlib.import_module("decoupledbo.modules.acquisition.discretekg")


def wrap_model_for_reference(oracle_model):
    """OracleModelList -> (stub) botorch ModelListGP the reference code accepts."""
    load_reference_discretekg()
    from botorch.models import ModelListGP, SingleTaskGP

    return ModelListGP(*[SingleTaskGP(o) for o in oracle_model.models])
