def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    raise NotImplementedError(f'seaborn.{name} is not available (oracle shim)')
