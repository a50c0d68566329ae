from . import radius_graph  # noqa: F401
from .radius_graph import RadiusGraph  # noqa: F401
