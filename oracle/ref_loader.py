"""TEST INFRASTRUCTURE ONLY.  Puts oracle/shim (torch_geometric & friends stand-ins) and the read-only
reference checkout on sys.path so the reference's own `gcbf` package imports unchanged.  Usable only in
the build container (/root/reference does not exist on the GPU box); used by oracle/make_golden.py and
by the CPU tests that validate oracle/gcbf_oracle.py against the real reference code."""
import os
import sys

REFERENCE_ROOT = os.environ.get('GCBF_REFERENCE_ROOT', '/root/reference')
SHIM_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'shim')


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'gcbf'))


def load_reference():
    """Import and return the reference's `gcbf` package (on the shim).  Must not be called in a process
    that already imported the product's `gcbf` package (same top-level name)."""
    if not reference_available():
        raise RuntimeError(f'reference checkout not found at {REFERENCE_ROOT}')
    if 'gcbf' in sys.modules and not sys.modules['gcbf'].__file__.startswith(REFERENCE_ROOT):
        raise RuntimeError('a different `gcbf` package is already imported in this process')
    for p in (REFERENCE_ROOT, SHIM_ROOT):
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    import gcbf  # noqa: F401
    return gcbf
