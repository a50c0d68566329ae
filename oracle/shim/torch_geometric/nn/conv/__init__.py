from . import message_passing  # noqa: F401
