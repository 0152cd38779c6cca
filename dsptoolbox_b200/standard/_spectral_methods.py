"""Welch cross-spectra on the GPU behind the reference's private entry points.

`_csm_welch` and `_welch` keep the signatures, return types and error behaviour of
reference standard/_spectral_methods.py:285-371 and :10-173; the arithmetic runs in
libb2dsp (K0/K1/K2/K2e, see engine.py).  Instead of one Welch call per channel pair, every
channel is transformed once and each bin's cross-spectral matrix is one Hermitian rank-K
update on the tensor cores.

`_csm_fft` (:374-443), the single-FFT spectrum behind `Signal.get_spectrum` (classes/signal.py:903-946) and `_stft`
(:176-282, any FFT length) run on the large / arbitrary-length FFT of csrc/big_fft.cu.

Not supported on the GPU path (raises NotImplementedError, there is no CPU fallback):
average="median" (reference :153-162), Welch window lengths above 16384, fractional-octave smoothing.
"""

from __future__ import annotations

from warnings import warn

import numpy as np
from scipy.signal import check_COLA

from .. import engine
from .enums import SpectrumScaling, Window


def _validate(window_length_samples, overlap_percent, average):
    # same checks and messages as the reference (_spectral_methods.py:89-100)
    valid_window_sizes = [int(2 ** p) for p in range(3, 19)]
    assert window_length_samples in valid_window_sizes, (
        "Window length should be a power of 2 between [8, 262_144] or [2**3, 2**18]")
    assert overlap_percent >= 0 and overlap_percent < 100, "overlap_percent should be between 0 and 100"
    assert average in ("mean", "median"), f"{average} is not valid. Use either mean or median"
    if average == "median":
        raise NotImplementedError(
            "average='median' needs every per-frame cross-spectrum and is not part of the B200 path")


def _window_and_step(window_type, window_length_samples, overlap_percent):
    """Periodic scipy window, hop size and the COLA warning of the reference (:103-114)."""
    spec = window_type.to_scipy_format() if isinstance(window_type, Window) else window_type
    from scipy.signal.windows import get_window

    window = get_window(spec, window_length_samples, fftbins=True)
    overlap_samples = int(overlap_percent / 100 * window_length_samples)
    step = window_length_samples - overlap_samples
    if not check_COLA(window, nperseg=len(window), noverlap=overlap_samples):
        warn("Selected window type and overlap do not meet the constant overlap and add constraint! "
             "Results might be distorted")
    return window, step


def _scaling_terms(scaling: SpectrumScaling, window_length_samples: int, fs_hz: int, window):
    """(fft_scale, factor, halve_edges, amplitude_sqrt) -- reference :141-148 and :164-171."""
    factor = 1.0
    halve = scaling.has_physical_units()
    if halve:
        factor = float(np.squeeze(scaling.get_scaling_factor(window_length_samples, fs_hz, window)))
    return scaling.fft_scale(window_length_samples), factor, halve, scaling.is_amplitude_scaling()


def csm_welch_device(time_data, sampling_rate_hz, window_length_samples, window_type, overlap_percent,
                     detrend, average, scaling, frame_range=None, reduce_fn=None, total_length=None,
                     row_offset=0):
    """Welch CSM left on the device: (f [numpy], csm [CUDA complex128 tensor (nb, M, M)]).

    `frame_range` restricts this process to a shard of the Welch frames and `reduce_fn` (e.g. an
    NCCL all-reduce, see dist.py) is applied to the partial sums before the non-linear
    finalisation -- the frame-sharded multi-GPU form of the same computation.  A rank may pass
    only its rows of the recording: `time_data` then holds rows [row_offset, row_offset + len) of a
    recording of `total_length` samples.
    """
    _validate(window_length_samples, overlap_percent, average)
    if not hasattr(time_data, "shape"):
        time_data = np.asarray(time_data)
    # the reference indexes time_data[:, channel] (_spectral_methods.py:351-356): anything but (time, channel) fails there
    assert len(time_data.shape) == 2, "time data must have the shape (time samples, channels)"
    window, step = _window_and_step(window_type, window_length_samples, overlap_percent)
    fft_scale, factor, halve, amp = _scaling_terms(scaling, window_length_samples, sampling_rate_hz, window)
    S, n_frames = engine.csm_partial_sums(
        time_data, int(total_length if total_length is not None else time_data.shape[0]),
        window_length_samples, step, window, detrend, fft_scale, frame_range=frame_range,
        row_offset=row_offset, reduce_fn=reduce_fn)
    csm = engine.csm_finalize(S, int(time_data.shape[1]), n_frames, factor, halve, amp)
    f = np.fft.rfftfreq(window_length_samples, 1 / sampling_rate_hz)
    return f, csm


def welch_autospectra_device(time_data, sampling_rate_hz, window_length_samples, window_type, overlap_percent,
                             detrend, average, scaling):
    """Welch auto-spectra of every channel, left on the device: float64 CUDA tensor (nb, M).
    Same frames / window / FFT kernels as the CSM, but only sum_k |X|^2 is accumulated."""
    _validate(window_length_samples, overlap_percent, average)
    window, step = _window_and_step(window_type, window_length_samples, overlap_percent)
    fft_scale, factor, halve, amp = _scaling_terms(scaling, window_length_samples, sampling_rate_hz, window)
    P, n_frames = engine.csm_partial_sums(time_data, int(time_data.shape[0]), window_length_samples, step, window,
                                          detrend, fft_scale, diag_only=True)
    return engine.auto_finalize(P, n_frames, factor, halve, amp)


def _csm_welch(time_data, sampling_rate_hz: int, window_length_samples: int, window_type,
               overlap_percent, detrend: bool, average: str, scaling: SpectrumScaling):
    """Cross-spectral matrix (frequency, channels, channels) with Welch's method.

    Drop-in for reference _spectral_methods.py:285-371: returns host arrays
    (f float64 [nb], csm complex128 [nb, M, M]).
    """
    f, csm = csm_welch_device(time_data, sampling_rate_hz, window_length_samples, window_type,
                              overlap_percent, detrend, average, scaling)
    return f, engine.to_host(csm)


def _welch(x, y, fs_hz: int, window_type, window_length_samples: int, overlap_percent: float,
           detrend: bool, average: str, scaling: SpectrumScaling):
    """Auto (y=None, real) or cross (complex) spectral density of 1-D or (time, channel) data.

    Drop-in for reference _spectral_methods.py:10-173, evaluated as entries of a small CSM:
    csd[:, c] = mean_k conj(X_c) Y_c is element [y_c, x_c] of the CSM of the stacked channels.
    """
    x = np.asarray(x).squeeze()
    auto = y is None
    if not auto:
        y = np.asarray(y).squeeze()
        assert x.shape == y.shape, "Shapes of data do not match"
    assert x.ndim <= 2, f"{x.shape} are too many dimensions. Use flat arrays or 2D-Arrays instead"
    flat = x.ndim == 1
    xs = x[:, None] if flat else x
    ch = xs.shape[1]
    if auto:
        out = welch_autospectra_device(xs, fs_hz, window_length_samples, window_type, overlap_percent,
                                       detrend, average, scaling).cpu().numpy()
    else:
        stacked = np.concatenate([xs, y[:, None] if flat else y], axis=1)
        _, csm = csm_welch_device(stacked, fs_hz, window_length_samples, window_type, overlap_percent,
                                  detrend, average, scaling)
        idx = np.arange(ch)
        out = csm[:, idx + ch, idx].cpu().numpy()
    return out[:, 0] if flat else out


# --------------------------------------------------------------------------- next row f4: STFT
def stft_device(time_data, fs_hz: int, window_length_samples: int, window_type, overlap_percent: float,
                fft_length_samples, detrend: bool, padding: bool, scaling: SpectrumScaling):
    """STFT left on the device: (time_s, freqs_hz [numpy], stft CUDA tensor (frequency, time, channel)),
    complex128, or float64 for the power scalings -- reference `_stft`, _spectral_methods.py:176-282.
    Framing, window, mean removal and the real FFT are the kernels of the Welch stage; the scaling
    epilogue of :266-277 runs as a few element-wise device operations."""
    import torch

    valid_window_sizes = [int(2 ** p) for p in range(4, 17)]
    assert window_length_samples in valid_window_sizes, (
        "Window length should be a power of 2 between [16, 65536] or [2**4, 2**16]")
    assert overlap_percent >= 0 and overlap_percent < 100, "overlap_percent should be between 0 and 100"
    n_fft = window_length_samples if fft_length_samples is None else int(fft_length_samples)
    assert n_fft >= 2, "the FFT length must be at least 2"
    spec = window_type.to_scipy_format() if isinstance(window_type, Window) else window_type
    from scipy.signal.windows import get_window

    window = get_window(spec, window_length_samples, fftbins=True)
    overlap_samples = int(overlap_percent / 100 * window_length_samples + 0.5)      # rounded, unlike _welch (:248)
    step = window_length_samples - overlap_samples
    if not check_COLA(window, nperseg=len(window), noverlap=overlap_samples):
        warn("Selected window type and overlap do not meet the constant overlap and add constraint! "
             "Results might be distorted")
    if not hasattr(time_data, "shape"):
        time_data = np.asarray(time_data)
    if len(time_data.shape) == 1:          # one channel given flat: (time,) -> (time, 1)
        time_data = time_data[:, None]
    assert len(time_data.shape) == 2, "time data must have the shape (time samples, channels)"
    if padding:
        if isinstance(time_data, torch.Tensor):
            time_data = torch.nn.functional.pad(time_data, (0, 0, overlap_samples, overlap_samples))
        else:
            time_data = np.pad(np.asarray(time_data), ((overlap_samples, overlap_samples), (0, 0)))
    length = int(time_data.shape[0])
    norm_scale = {"backward": 1.0, "forward": 1.0 / n_fft, "ortho": n_fft ** -0.5}[scaling.fft_norm()]
    if n_fft == window_length_samples and window_length_samples <= 16384:
        # the shared-memory kernels of the Welch stage
        stft = engine.frame_spectra(time_data, length, window_length_samples, step, window, detrend, norm_scale)
    else:
        # another FFT length (numpy crops or zero-pads the frames, :263) or a window beyond the shared-memory kernel:
        # framing from the channel-major copy, large / arbitrary-length FFT (Bluestein unless a power of two)
        stft = engine.frame_spectra_any(time_data, length, window_length_samples, step, window, detrend, n_fft, norm_scale)
    if scaling.has_physical_units():
        stft[0] /= 2 ** 0.5
        if n_fft % 2 == 0:
            stft[-1] /= 2 ** 0.5                   # (:268)
        factor = float(np.squeeze(scaling.get_scaling_factor(n_fft, fs_hz, window)))
        if not scaling.is_amplitude_scaling():
            stft = stft.real ** 2 + stft.imag ** 2
        stft = stft * factor
    time_s = np.linspace(0, length / fs_hz, stft.shape[1])
    return time_s, np.fft.rfftfreq(window_length_samples, 1 / fs_hz), stft


def _stft(x, fs_hz: int, window_length_samples: int, window_type, overlap_percent: float,
          fft_length_samples, detrend: bool, padding: bool, scaling: SpectrumScaling):
    """Drop-in for reference _spectral_methods.py:176-282: host arrays (time_s, freqs_hz, stft)."""
    t, f, st = stft_device(x, fs_hz, window_length_samples, window_type, overlap_percent, fft_length_samples,
                           detrend, padding, scaling)
    return t, f, engine.to_host(st)


# --------------------------------------------------------------------------- next row f3: single-FFT spectrum and CSM
def spectrum_fft_device(time_data, sampling_rate_hz: int, scaling: SpectrumScaling, pad_to_fast_length: bool = True,
                        smoothing: float = 0.0, window=None):
    """The FFT branch of `Signal.get_spectrum` (reference classes/signal.py:903-946) left on the device:
    (freqs [numpy], spectrum CUDA tensor (nb, channels)) -- complex128 for the amplitude scalings, float64 power for
    PowerSpectrum / PowerSpectralDensity.  One FFT over the whole recording of length next_fast_len(L) (or L)."""
    import torch
    from scipy.fft import next_fast_len

    if smoothing != 0:
        raise NotImplementedError("fractional-octave smoothing of the spectrum is outside the accelerated path")
    if not hasattr(time_data, "shape"):
        time_data = np.asarray(time_data)
    if len(time_data.shape) == 1:
        time_data = time_data[:, None]
    L = int(time_data.shape[0])
    n = int(next_fast_len(L, True)) if pad_to_fast_length else L
    x = time_data if isinstance(time_data, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(time_data))
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(torch.float64)
    x = x.to(torch.device("cuda", torch.cuda.current_device()), non_blocking=True)
    norm_scale = {"backward": 1.0, "forward": 1.0 / n, "ortho": n ** -0.5}[scaling.fft_norm()]
    sp = engine.rfft_rows(x, n, norm_scale)
    if scaling.has_physical_units():
        # `_scale_spectrum` (helpers/spectrum_utilities.py:268-330): DC (and Nyquist for an even length) / sqrt(2),
        # |.|^2 for the power scalings, times the scaling factor
        sp[0] /= 2 ** 0.5
        if n % 2 == 0:
            sp[-1] /= 2 ** 0.5
        factor = torch.from_numpy(np.atleast_1d(scaling.get_scaling_factor(n, sampling_rate_hz, window)).astype(np.float64)).to(sp.device)
        if not scaling.is_amplitude_scaling():
            sp = sp.real ** 2 + sp.imag ** 2
        sp = sp * factor
    return np.fft.rfftfreq(n, 1 / sampling_rate_hz), sp


def csm_fft_device(spectrum_dev, scaling: SpectrumScaling, window, sampling_rate_hz: int):
    """`_csm_fft` (reference _spectral_methods.py:374-443) on a device spectrum (nb, channels) with FFTBackward
    normalisation -> complex128 CUDA tensor (nb, channels, channels)."""
    nb = int(spectrum_dev.shape[0])
    if scaling == SpectrumScaling.FFTBackward:
        return engine.csm_rank1(spectrum_dev, 1.0, False, False)
    # the reference passes `spectrum.shape[0] // 2 + 1` as the time length of the conversion (:437-439)
    factor = np.atleast_1d(SpectrumScaling.FFTBackward.conversion_factor(scaling, nb // 2 + 1, sampling_rate_hz, window))
    if factor.size != 1:
        raise NotImplementedError("per-channel time windows are outside the accelerated path")
    return engine.csm_rank1(spectrum_dev, float(factor.ravel()[0]), True, scaling.is_amplitude_scaling())


def _csm_fft(spectrum, scaling: SpectrumScaling, window, sampling_rate_hz: int):
    """Drop-in for reference _spectral_methods.py:374-443: complex128 host array (frequency, channels, channels)."""
    import torch

    if window is not None:
        assert np.asarray(window).shape[1] == np.asarray(spectrum).shape[1], "Number of channels does not match"
    sp = torch.from_numpy(np.ascontiguousarray(spectrum, dtype=np.complex128)).to(torch.device("cuda", torch.cuda.current_device()))
    return csm_fft_device(sp, scaling, window, sampling_rate_hz).cpu().numpy()
