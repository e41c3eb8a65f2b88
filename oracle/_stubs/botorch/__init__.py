"""ORACLE support (test infrastructure): a minimal stand-in for the *names* the reference imports
from botorch (``discretekg.py:12-20``), backed by ``oracle.gp`` for the posterior.

It exists so that, in the build container (where /root/reference is mounted but botorch/gpytorch
are not installable), the reference's OWN functions can be imported and executed to (a) pin the
oracle restatement and (b) generate ``tests/golden/`` vectors (``oracle/make_golden.py``).
Nothing in the product imports it.
"""
