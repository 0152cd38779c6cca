from .enums import SpectrumMethod, SpectrumScaling, Window  # noqa: F401
