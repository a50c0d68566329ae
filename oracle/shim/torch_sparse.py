"""Shim: name only (type annotation at gcbf/nn/gnn.py:6)."""
class SparseTensor:  # noqa: E302
    pass
