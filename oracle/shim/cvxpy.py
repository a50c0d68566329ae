"""Shim: names only (isinstance checks at gcbf/env/simple_car.py:80, dubins_car.py:111, simple_drone.py:104)."""
class Expression:  # noqa: E302
    pass
class Variable(Expression):  # noqa: E302
    def __init__(self, *a, **k):
        raise NotImplementedError('cvxpy is not available (oracle shim)')
