class Axes:
    pass
def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise NotImplementedError(f'matplotlib.pyplot.{name} is not available (oracle shim)')
