"""Steering-vector formulations (reference beamforming/enums.py:4-18; Sarradj 2012, formulations 1-4)."""

from enum import Enum, auto


class SteeringVectorType(Enum):
    Classic = auto()
    Inverse = auto()
    TruePower = auto()
    TrueLocation = auto()
