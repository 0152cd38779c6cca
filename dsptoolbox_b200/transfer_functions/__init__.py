"""Transfer-function estimators for stochastic signals (next row f3 of SURVEY.md section 8f):
`compute_transfer_function` H1 / H2 / H3 with coherence, evaluated as 2x2 blocks of ONE Welch
cross-spectral matrix of the stacked [input | output] channels on the device."""

from .enums import TransferFunctionType  # noqa: F401
from .transfer_functions import compute_transfer_function  # noqa: F401

__all__ = ["TransferFunctionType", "compute_transfer_function"]
