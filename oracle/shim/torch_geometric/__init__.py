"""TEST INFRASTRUCTURE ONLY -- pure-torch stand-in for the torch_geometric==2.3.0 surface that
MIT-REALM/gcbf-pytorch imports.  torch_geometric / torch_cluster / torch_scatter
are not installed in this image and cannot be installed (no network), so this package is a
*restatement* of their published semantics at exactly the call sites the reference uses
(gcbf/nn/gnn.py:4-9, gcbf/algo/gcbf.py:7-8, gcbf/env/*.py).  "parity unpinned": the reference ships no
golden vectors for these third-party routines; the only in-repo pin is the pretrained checkpoints
(semantic known-answer, tests/test_oracle_cpu.py).

Nothing in the product package imports this.  It exists so the reference's own `gcbf` package can be
imported unchanged (oracle/ref_loader.py) to produce tests/golden/*.pt and to validate oracle/gcbf_oracle.py.
"""
__version__ = "2.3.0+shim"
from . import data, utils, nn, transforms  # noqa: F401
