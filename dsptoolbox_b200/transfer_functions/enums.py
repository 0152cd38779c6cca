"""reference transfer_functions/enums.py:4-15."""

from enum import Enum, auto


class TransferFunctionType(Enum):
    """H1 = Gxy/Gxx (noise in the output), H2 = Gyy/Gyx (noise in the input),
    H3 = Gxy/|Gxy| * sqrt(Gyy/Gxx) (noise in both)."""

    H1 = auto()
    H2 = auto()
    H3 = auto()
