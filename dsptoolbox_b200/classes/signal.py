"""`Signal`: the reference's time-series container, restricted to what the CSM path needs.

API-compatible with reference classes/signal.py for: the constructor (:57-115), `time_data`
(:215-301), `sampling_rate_hz` (:321-348), `set_spectrum_parameters` (:497-588),
`spectrum_scaling` / `spectrum_method` (:590-666), `get_csm` (:948-1007) and
`number_of_channels` / `len()`.  Plotting, file I/O, filtering, STFT etc. are out of scope
 and raise NotImplementedError.

`get_csm()` returns fresh host arrays like the reference; `get_csm_device()` is the additional
zero-copy accessor the beamformers use so the matrix never crosses PCIe between the stages.
"""

from __future__ import annotations

from warnings import warn

import numpy as np

from .. import engine
from ..standard import _spectral_methods as _sm
from ..standard.enums import SpectrumMethod, SpectrumScaling, Window
from ._multichannel_data import MultichannelData


class Signal(MultichannelData):
    def __init__(self, path: str | None = None, time_data=None, sampling_rate_hz: int | None = None,
                 constrain_amplitude: bool = False, activate_cache: bool = False):
        self.constrain_amplitude = constrain_amplitude
        self.activate_cache = activate_cache
        if path is not None:
            assert time_data is None, "Constructor cannot take a path and a vector at the same time"
            assert sampling_rate_hz is None, "Constructor cannot take a path and a sampling rate at the same time"
            raise NotImplementedError("Audio file reading is outside the accelerated path; pass time_data")
        assert time_data is not None, "Either a path to an audio file or a time vector has to be passed"
        assert sampling_rate_hz is not None, "A sampling rate should be passed!"
        self.sampling_rate_hz = sampling_rate_hz
        self.time_data = time_data
        self.set_spectrum_parameters()
        self.set_spectrogram_parameters()

    @staticmethod
    def from_time_data(time_data, sampling_rate_hz: int, constrain_amplitude: bool = True):
        return Signal(None, time_data, sampling_rate_hz, constrain_amplitude)

    # ---- state ---------------------------------------------------------------------------
    def _invalidate(self):
        self._csm_dirty = True
        self._csm_dev = None
        self._spectrogram_dirty = True
        self._spectrogram = None

    def _get_data(self):
        return self._time_data

    # ---- time data -------------------------------------------------------------------------
    @property
    def time_data(self):
        return self._time_data

    @time_data.setter
    def time_data(self, new_time_data):
        if hasattr(new_time_data, "is_cuda"):
            # a torch tensor (device-resident or pinned-host recording) is held as it is -- an extension of the
            # reference, whose arrays are always numpy: same (time samples, channels) rule, same auto-transpose
            td = new_time_data
            assert td.ndim <= 2, (f"{td.ndim} are too many dimensions for time data. Dimensions should"
                                  " be [time samples, channels]")
            if td.ndim < 2:
                td = td[..., None]
            if td.shape[1] > td.shape[0]:
                td = td.T
            assert not td.is_complex(), "complex torch recordings are not supported"
            self._amplitude_scale_factor = 1.0
            if self.constrain_amplitude:
                peak = float(td.abs().max())
                if peak > 1.0:
                    td = td / peak
                    warn("Signal was over 0 dBFS, normalizing to 0 dBFS peak level was triggered")
                    self._amplitude_scale_factor = 1.0 / peak
            self._time_data = td
            self.time_data_imaginary = None
            self._invalidate()
            return
        td = np.atleast_2d(new_time_data).squeeze()
        assert td.ndim <= 2, (f"{td.ndim} are too many dimensions for time data. Dimensions should"
                              " be [time samples, channels]")
        if td.ndim < 2:
            td = td[..., None]
        if td.shape[1] > td.shape[0]:      # more columns than rows: rows are channels
            td = td.T
        imag = None
        if np.iscomplexobj(td):
            imag, td = np.imag(td), np.real(td)
        self._amplitude_scale_factor = 1.0
        if self.constrain_amplitude:
            peak = np.max(np.abs(td))
            if imag is not None:
                peak = max(peak, np.max(np.abs(imag)))
            if peak > 1.0:
                td /= peak
                if imag is not None:
                    imag /= peak
                warn("Signal was over 0 dBFS, normalizing to 0 dBFS peak level was triggered")
                self._amplitude_scale_factor = 1.0 / peak
        self._time_data = td
        self.time_data_imaginary = imag
        self._invalidate()

    @property
    def amplitude_scale_factor(self) -> float:
        return self._amplitude_scale_factor

    @property
    def sampling_rate_hz(self) -> int:
        return self._sampling_rate_hz

    @sampling_rate_hz.setter
    def sampling_rate_hz(self, new_sampling_rate_hz):
        assert type(new_sampling_rate_hz) is int, "Sampling rate can only be an integer"
        self._sampling_rate_hz = new_sampling_rate_hz
        self._invalidate()

    @property
    def length_samples(self) -> int:
        return self._time_data.shape[0]

    @property
    def length_seconds(self) -> float:
        return len(self) / self.sampling_rate_hz

    # ---- spectrum parameters -----------------------------------------------------------------
    def set_spectrum_parameters(self, method: SpectrumMethod = SpectrumMethod.WelchPeriodogram,
                                smoothing: int = 0, pad_to_fast_length: bool = True,
                                window_length_samples: int = 1024, window_type: Window = Window.Hann,
                                overlap_percent: float = 50, detrend: bool = True, average: str = "mean",
                                scaling: SpectrumScaling = SpectrumScaling.FFTBackward):
        new = dict(method=method, smoothing=smoothing, pad_to_fast_length=pad_to_fast_length,
                   window_length_samples=window_length_samples, window_type=window_type,
                   overlap_percent=overlap_percent, detrend=detrend, average=average, scaling=scaling)
        old = getattr(self, "_spectrum_parameters", None)
        if old is None or any(old[k] != new[k] for k in old):
            self._spectrum_parameters = new
            self._invalidate()
        return self

    @property
    def spectrum_scaling(self) -> SpectrumScaling:
        return self._spectrum_parameters["scaling"]

    @spectrum_scaling.setter
    def spectrum_scaling(self, new_scaling: SpectrumScaling):
        assert isinstance(new_scaling, SpectrumScaling)
        self._spectrum_parameters["scaling"] = new_scaling
        self._invalidate()

    @property
    def spectrum_method(self) -> SpectrumMethod:
        return self._spectrum_parameters["method"]

    @spectrum_method.setter
    def spectrum_method(self, new_method: SpectrumMethod):
        assert isinstance(new_method, SpectrumMethod)
        self._spectrum_parameters["method"] = new_method
        self._invalidate()

    # ---- cross-spectral matrix -----------------------------------------------------------------
    def _compute_csm_device(self, frame_range=None, reduce_fn=None):
        p = self._spectrum_parameters
        if self.spectrum_method != SpectrumMethod.WelchPeriodogram:
            # reference signal.py:989-1002: the spectrum is taken with FFTBackward scaling, `_csm_fft` applies the
            # requested one
            assert frame_range is None and reduce_fn is None, "the single-FFT cross-spectral matrix is not sharded"
            f, sp = _sm.spectrum_fft_device(self.time_data, self.sampling_rate_hz, SpectrumScaling.FFTBackward,
                                            p["pad_to_fast_length"], p["smoothing"], None)
            return f, _sm.csm_fft_device(sp, p["scaling"], None, self.sampling_rate_hz)
        return _sm.csm_welch_device(self.time_data, self.sampling_rate_hz, p["window_length_samples"],
                                    p["window_type"], p["overlap_percent"], p["detrend"], p["average"],
                                    p["scaling"], frame_range=frame_range, reduce_fn=reduce_fn)

    def get_csm_device(self, force_computation: bool = False):
        """(f, csm) with csm a complex128 CUDA tensor (frequency, channels, channels).

        With `activate_cache` the tensor is kept on the device between calls under the same
        invalidation rules as the reference's host cache (signal.py:163-171, 973-1007); the
        returned tensor is then shared, not copied.
        """
        assert self.number_of_channels > 1, (
            "Cross spectral matrix can only be computed when at least two channels are available")
        if self._csm_dev is None or force_computation or self._csm_dirty:
            f, csm = self._compute_csm_device()
            if not self.activate_cache:
                return f, csm
            self._csm_dev = (f, csm)
            self._csm_dirty = False
        return self._csm_dev[0].copy(), self._csm_dev[1]

    def get_csm(self, force_computation: bool = False):
        """Cross-spectral matrix as fresh host arrays (f float64 [nb], csm complex128 [nb, M, M])."""
        f, csm = self.get_csm_device(force_computation)
        return f, engine.to_host(csm)

    # ---- explicitly out of scope -------------------------------------------------------------------
    def get_spectrum(self, force_computation: bool = False):
        """(f, spectrum): the Welch auto-spectrum of every channel, float64 (nb, channels), or -- SpectrumMethod.FFT --
        the spectrum of the whole recording (complex128 for the amplitude scalings), reference signal.py:861-946
        (SURVEY.md section 8f.3).  Fractional-octave smoothing is outside the accelerated path."""
        p = self._spectrum_parameters
        if self.spectrum_method != SpectrumMethod.WelchPeriodogram:
            f, sp = _sm.spectrum_fft_device(self.time_data, self.sampling_rate_hz, p["scaling"], p["pad_to_fast_length"],
                                            p["smoothing"], None)
            return f, sp.cpu().numpy()
        spec = _sm.welch_autospectra_device(self.time_data, self.sampling_rate_hz, p["window_length_samples"],
                                            p["window_type"], p["overlap_percent"], p["detrend"], p["average"],
                                            p["scaling"])
        freqs = np.fft.rfftfreq(p["window_length_samples"], 1 / self.sampling_rate_hz)
        return freqs, spec.cpu().numpy()

    # ---- spectrogram (next row f4) ---------------------------------------------------------------
    def set_spectrogram_parameters(self, window_length_samples: int = 1024, window_type: Window = Window.Hann,
                                   overlap_percent: float = 50.0, fft_length_samples: int | None = None,
                                   detrend: bool = False, padding: bool = True,
                                   scaling: SpectrumScaling = SpectrumScaling.FFTBackward):
        """Parameters of the STFT, reference signal.py:706-773 (same names and defaults)."""
        new = dict(window_length_samples=window_length_samples, window_type=window_type,
                   overlap_percent=overlap_percent, fft_length_samples=fft_length_samples, detrend=detrend,
                   padding=padding, scaling=scaling)
        old = getattr(self, "_spectrogram_parameters", None)
        if old is None or any(old[k] != new[k] for k in old):
            self._spectrogram_parameters = new
            self._spectrogram_dirty = True
            self._spectrogram = None
        return self

    def get_spectrogram(self, force_computation: bool = False):
        """(t_s, f_hz, spectrogram (frequency, time, channel)) -- reference signal.py:1009-1056 on the
        framing / window / rFFT kernels of the Welch stage.  Host arrays, complex128 (float64 for the power
        scalings); with `activate_cache` the result is kept until the data or the parameters change."""
        p = self._spectrogram_parameters
        if getattr(self, "_spectrogram", None) is None or force_computation or self._spectrogram_dirty:
            res = _sm._stft(self.time_data, self.sampling_rate_hz, p["window_length_samples"], p["window_type"],
                            p["overlap_percent"], p["fft_length_samples"], p["detrend"], p["padding"], p["scaling"])
            self._spectrogram_dirty = False
            if not self.activate_cache:
                return res
            self._spectrogram = res
        return tuple(a.copy() for a in self._spectrogram)
