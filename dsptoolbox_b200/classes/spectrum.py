"""Minimal `Spectrum` data holder: what `compute_transfer_function` returns (reference
classes/spectrum.py:30-56, 139-260, 871-885).  Only the fields and their validation are mirrored;
interpolation, smoothing, resampling and plotting are outside the accelerated path."""

from __future__ import annotations

import numpy as np


class Spectrum:
    def __init__(self, frequency_vector_hz, spectral_data):
        self.frequency_vector_hz = frequency_vector_hz
        self.spectral_data = spectral_data

    @staticmethod
    def from_signal(sig, complex: bool = False) -> "Spectrum":
        """From `sig.get_spectrum()` with the signal's own parameters (reference spectrum.py:56-88): the complex
        spectrum, or its magnitude (square root of the power for the power scalings)."""
        if complex:
            assert sig.spectrum_scaling.outputs_complex_spectrum(sig.spectrum_method), (
                "Method or scaling do not deliver a complex spectrum")
        f, sp = sig.get_spectrum()
        if complex:
            assert np.iscomplexobj(sp), "Spectrum of signal is not complex"
            return Spectrum(f, sp)
        return Spectrum(f, np.abs(sp) if sig.spectrum_scaling.is_amplitude_scaling() else np.abs(sp) ** 0.5)

    @property
    def frequency_vector_hz(self):
        return self.__frequency_vector_hz

    @frequency_vector_hz.setter
    def frequency_vector_hz(self, new_freqs):
        f = np.atleast_1d(np.asarray(new_freqs).squeeze())
        assert not np.iscomplexobj(f), "Frequency vector cannot be complex"
        assert f.ndim == 1, "Frequency vector must have only one dimension"
        assert np.all(f >= 0), "Frequency vector must contain only non-negative values"
        assert np.all(np.diff(f) > 0), "Frequency vector must be strictly ascending"
        self.__frequency_vector_hz = f.astype(np.float64)

    @property
    def spectral_data(self):
        return self.__spectral_data

    @spectral_data.setter
    def spectral_data(self, new_data):
        d = np.atleast_2d(np.asarray(new_data))
        assert d.ndim == 2, "Spectral data must be broadcastable to two dimensions"
        if d.shape[0] != len(self.frequency_vector_hz) and d.shape[1] == len(self.frequency_vector_hz):
            d = d.T
        assert d.shape[0] == len(self.frequency_vector_hz), "Frequency vector and spectral data lengths do not match"
        if not np.iscomplexobj(d):
            assert np.all(d >= 0), "Magnitude spectra cannot contain negative values"
        self.__spectral_data = d

    @property
    def number_of_channels(self) -> int:
        return self.spectral_data.shape[1]

    @property
    def number_frequency_bins(self) -> int:
        return self.spectral_data.shape[0]

    def set_coherence(self, coherence):
        """reference spectrum.py:871-885."""
        assert coherence.shape == self.spectral_data.shape, "Length of signals and given coherence do not match"
        assert not np.iscomplexobj(coherence), "Coherence cannot be complex"
        self.coherence = coherence
