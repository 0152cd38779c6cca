"""Beamforming front-end of the accelerated path (mirrors reference beamforming/__init__.py)."""

from .beamforming import (  # noqa: F401
    Grid,
    Regular2DGrid,
    Regular3DGrid,
    LineGrid,
    MicArray,
    SteeringVector,
    BeamformerDASFrequency,
    BeamformerMVDR,
    BeamformerFunctional,
    BeamformerOrthogonal,
    BeamformerCleanSC,
    classic_steering,
    inverse_steering,
    true_power_steering,
    true_location_steering,
    simpson_weights,
    MonopoleSource,
    mix_sources_on_array,
)
from .enums import SteeringVectorType  # noqa: F401

__all__ = ["Grid", "Regular2DGrid", "Regular3DGrid", "LineGrid", "MicArray", "SteeringVector",
           "BeamformerDASFrequency", "BeamformerMVDR", "BeamformerFunctional", "BeamformerOrthogonal", "BeamformerCleanSC", "SteeringVectorType",
           "MonopoleSource", "mix_sources_on_array"]
