from . import attention  # noqa: F401
from .attention import AttentionalAggregation  # noqa: F401
