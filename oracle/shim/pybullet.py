"""Shim: import-time name only (demo/render paths are never exercised by the oracle)."""
DIRECT = GUI = GEOM_BOX = 0
def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise NotImplementedError(f'pybullet.{name} is not available (oracle shim)')
