"""Training / evaluation loop with the reference's entry point name (gcbf/trainer)."""
from .trainer import Trainer

__all__ = ['Trainer']
