"""ORACLE: CPU float64 restatement of the reference hot path. TEST INFRASTRUCTURE ONLY.

Importable only from ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs.  The product (``decoupled-kg_b200/``) must never import, call, link or execute anything
in this directory; it fails loudly if its CUDA library is missing instead of falling back here.
"""
