"""Enums of the accelerated path with the reference's member names and scaling rules.

Mirrors reference standard/enums.py: SpectrumMethod (:7-19), SpectrumScaling (:21-229, only the
rules the Welch path uses: rfft norm :53-75, amplitude-vs-power :77-92, physical units :109-123,
scaling factor :183-229) and Window (:341-437).  Written from the behaviour, not the text.
"""

from __future__ import annotations

from enum import Enum, auto

import numpy as np
from scipy.signal.windows import get_window as _get_window


class SpectrumMethod(Enum):
    WelchPeriodogram = auto()
    FFT = auto()


class SpectrumScaling(Enum):
    AmplitudeSpectrum = auto()
    AmplitudeSpectralDensity = auto()
    PowerSpectrum = auto()
    PowerSpectralDensity = auto()
    FFTBackward = auto()
    FFTForward = auto()
    FFTOrthogonal = auto()

    # -- classification -------------------------------------------------------------
    def fft_norm(self) -> str:
        return {SpectrumScaling.FFTForward: "forward", SpectrumScaling.FFTOrthogonal: "ortho"}.get(self, "backward")

    def is_amplitude_scaling(self) -> bool:
        return self not in (SpectrumScaling.PowerSpectrum, SpectrumScaling.PowerSpectralDensity)

    def has_physical_units(self) -> bool:
        return self in _PHYSICAL

    def is_spectral_density(self) -> bool:
        return self in (SpectrumScaling.AmplitudeSpectralDensity, SpectrumScaling.PowerSpectralDensity)

    def outputs_complex_spectrum(self, method: SpectrumMethod) -> bool:
        return method != SpectrumMethod.WelchPeriodogram and self.is_amplitude_scaling()

    # -- numbers ----------------------------------------------------------------------
    def fft_scale(self, window_length_samples: int) -> float:
        """Multiplier that turns an un-normalised rfft into one with `fft_norm()`."""
        norm = self.fft_norm()
        if norm == "forward":
            return 1.0 / window_length_samples
        if norm == "ortho":
            return 1.0 / np.sqrt(window_length_samples)
        return 1.0

    def conversion_factor(self, output: "SpectrumScaling", length_time_data_samples: int, sampling_rate_hz: int, window):
        """Factor from this scaling to `output` (reference enums.py:131-181): linear / linear or squared / squared
        when both are of the same kind, else the one to multiply SQUARED data with."""
        fin = self.get_scaling_factor(length_time_data_samples, sampling_rate_hz, window)
        fout = output.get_scaling_factor(length_time_data_samples, sampling_rate_hz, window)
        if self.is_amplitude_scaling() == output.is_amplitude_scaling():
            return fout / fin
        if self.is_amplitude_scaling():
            fin = fin ** 2.0
        else:
            fout = fout ** 2.0
        return fout / fin

    def get_scaling_factor(self, length_time_data_samples: int, sampling_rate_hz: int, window):
        """One-sided forward-transform factor; linear for amplitude scalings, squared for power."""
        n = length_time_data_samples
        if self == SpectrumScaling.FFTBackward:
            return np.atleast_1d(1.0)
        if self == SpectrumScaling.FFTForward:
            return np.atleast_1d(1.0 / n)
        if self == SpectrumScaling.FFTOrthogonal:
            return np.atleast_1d(n ** -0.5)
        if self.is_spectral_density():
            energy = n if window is None else np.sum(np.asarray(window) ** 2, axis=0, keepdims=True)
            linear = (2.0 / energy / sampling_rate_hz) ** 0.5
        else:
            gain = n if window is None else np.sum(np.asarray(window), axis=0, keepdims=True)
            linear = 2.0 ** 0.5 / gain
        linear = np.atleast_1d(linear)
        return linear if self.is_amplitude_scaling() else linear ** 2.0


_PHYSICAL = (SpectrumScaling.AmplitudeSpectrum, SpectrumScaling.AmplitudeSpectralDensity,
             SpectrumScaling.PowerSpectrum, SpectrumScaling.PowerSpectralDensity)

_WINDOW_SCIPY_NAMES = {"KaiserBesselDerived": "kaiser_bessel_derived", "GeneralCosine": "general_cosine",
                       "GeneralGaussian": "general_gaussian", "GeneralHamming": "general_hamming"}
_PARAMETRIC = ("Kaiser", "KaiserBesselDerived", "Gaussian", "GeneralCosine", "GeneralGaussian",
               "GeneralHamming", "Dpss", "Chebwin")


class Window(Enum):
    """Window families of scipy.signal.windows.get_window."""

    Boxcar = auto()
    Triang = auto()
    Blackman = auto()
    Hamming = auto()
    Hann = auto()
    Bartlett = auto()
    Flattop = auto()
    Parzen = auto()
    Bohman = auto()
    Blackmanharris = auto()
    Nuttall = auto()
    Barthann = auto()
    Cosine = auto()
    Exponential = auto()
    Tukey = auto()
    Taylor = auto()
    Lanczos = auto()
    Kaiser = auto()
    KaiserBesselDerived = auto()
    Gaussian = auto()
    GeneralCosine = auto()
    GeneralGaussian = auto()
    GeneralHamming = auto()
    Dpss = auto()
    Chebwin = auto()

    @property
    def extra_parameter(self):
        return self.__dict__.get("_extra")

    def with_extra_parameter(self, extra_parameter):
        self._extra = extra_parameter
        return self

    def needs_extra_parameter(self) -> bool:
        return self.name in _PARAMETRIC

    def to_scipy_format(self):
        name = _WINDOW_SCIPY_NAMES.get(self.name, self.name.lower())
        if not self.needs_extra_parameter():
            return name
        if self == Window.GeneralGaussian:
            return (name, self.extra_parameter[0], self.extra_parameter[1])
        return (name, self.extra_parameter)

    def __call__(self, n_values: int, symmetric: bool):
        return _get_window(self.to_scipy_format(), n_values, not symmetric)
