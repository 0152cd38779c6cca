from .signal import Signal  # noqa: F401
