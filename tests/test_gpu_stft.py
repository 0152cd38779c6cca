"""Next row f4 parity: Signal.get_spectrogram / _stft on the GPU (the framing, window and real-FFT kernels of the
Welch stage without the averaging) against the golden vectors of the real reference and against the oracle."""

import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine  # noqa: E402
from dsptoolbox_b200.standard._spectral_methods import _stft, stft_device  # noqa: E402
from oracle import restatement as R  # noqa: E402

# fp64 FFT: far inside any sensible contract; the DC bin of de-meaned frames is exact 0 here and ~1e-14 of the
# maximum in the reference (rounding noise), hence an absolute bound relative to the largest value
REL_TO_MAX = 1e-11


@pytest.fixture(autouse=True)
def _strict():
    engine.config.csm_precision = "fp64"
    yield
    engine.config.csm_precision = "fp64"


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "stft_small.npz"))
    for i in range(int(g["n_cases"])):
        wl, ov, nfft, det, pad = g[f"params_{i}"]
        win, sc = (str(v) for v in g[f"names_{i}"])
        yield g, i, int(wl), win, float(ov), (None if nfft < 0 else int(nfft)), bool(det), bool(pad), sc


def test_golden_spectrograms(golden_dir):
    seen = 0
    for g, i, wl, win, ov, nfft, det, pad, sc in _cases(golden_dir):
        s = dsp.Signal(None, g["x"].copy(), int(g["fs"]))
        s.set_spectrogram_parameters(window_length_samples=wl, window_type=getattr(dsp.Window, win), overlap_percent=ov,
                                     fft_length_samples=nfft, detrend=det, padding=pad,
                                     scaling=getattr(dsp.SpectrumScaling, sc))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, f, st = s.get_spectrogram()           # nfft != window length runs on the any-length FFT (big_fft.cu)
        ref = g[f"stft_{i}"]
        assert st.shape == ref.shape and st.dtype == ref.dtype, (i, st.shape, ref.shape, st.dtype)
        assert np.max(np.abs(st - ref)) <= REL_TO_MAX * np.max(np.abs(ref)), i
        assert np.array_equal(f, g[f"f_{i}"]) and np.allclose(t, g[f"t_{i}"], rtol=0, atol=1e-15)
        seen += 1
    assert seen == 9


@pytest.mark.parametrize("L,M,wl,ov", [(5000, 1, 1024, 50.0), (12345, 5, 4096, 75.0), (40000, 3, 16384, 50.0),
                                       (777, 2, 16, 25.0), (20000, 17, 256, 90.0)])
def test_against_oracle_ragged(L, M, wl, ov):
    """Ragged lengths (zero-padded last frames), one channel, the largest supported window, 90 % overlap (many
    frames per CTA group) -- against the oracle, complex and power scalings, fp64 and fp32 input."""
    rng = np.random.default_rng(L)
    x = rng.standard_normal((L, M)) + 0.2
    for sc, det, pad in [("FFTBackward", False, True), ("PowerSpectralDensity", True, False)]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, f, st = _stft(x, 48000, wl, dsp.Window.Hann, ov, None, det, pad, getattr(dsp.SpectrumScaling, sc))
            to, fo, so = R.stft(x, 48000, wl, "hann", ov, None, det, pad, sc)
        assert st.shape == so.shape and st.dtype == so.dtype
        assert np.max(np.abs(st - so)) <= REL_TO_MAX * np.max(np.abs(so))
        assert np.array_equal(f, fo) and np.allclose(t, to, rtol=0, atol=1e-15)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        _, _, s32 = _stft(x.astype(np.float32), 48000, wl, dsp.Window.Hann, ov, None, False, True, dsp.SpectrumScaling.FFTBackward)
        _, _, o32 = R.stft(x.astype(np.float32), 48000, wl, "hann", ov, None, False, True, "FFTBackward")
    assert np.max(np.abs(s32 - o32)) <= REL_TO_MAX * np.max(np.abs(o32))


def test_device_tensor_input_cache_and_errors():
    rng = np.random.default_rng(3)
    x = rng.standard_normal((6000, 4))
    # a CUDA tensor as input stays on the device (padding included)
    t, f, st = stft_device(torch.from_numpy(x).cuda(), 8000, 512, dsp.Window.Hann, 50.0, None, False, True,
                           dsp.SpectrumScaling.FFTBackward)
    _, _, so = R.stft(x, 8000, 512, "hann", 50.0, None, False, True, "FFTBackward")
    assert st.is_cuda and np.max(np.abs(st.cpu().numpy() - so)) <= REL_TO_MAX * np.max(np.abs(so))
    # cache semantics of signal.py:1031-1056
    s = dsp.Signal(None, x, 8000, activate_cache=True)
    s.set_spectrogram_parameters(window_length_samples=256)
    a = s.get_spectrogram()
    n0 = engine.launches
    b = s.get_spectrogram()
    assert engine.launches == n0 and np.array_equal(a[2], b[2]) and a[2] is not b[2]
    s.set_spectrogram_parameters(window_length_samples=128)
    assert s.get_spectrogram()[2].shape[0] == 65 and engine.launches > n0
    # the reference's argument checks (bare asserts) and what this path does not do
    with pytest.raises(AssertionError):
        _stft(x, 8000, 100, dsp.Window.Hann, 50.0, None, False, True, dsp.SpectrumScaling.FFTBackward)
    with pytest.raises(AssertionError):
        _stft(x, 8000, 256, dsp.Window.Hann, 100.0, None, False, True, dsp.SpectrumScaling.FFTBackward)


@pytest.mark.parametrize("L,M,wl,nfft,ov", [(200000, 2, 32768, None, 50.0), (150000, 3, 65536, None, 75.0),
                                            (9000, 2, 512, 2048, 50.0), (9000, 3, 1024, 512, 25.0),
                                            (70000, 2, 16384, 32768, 50.0)])
def test_long_windows_and_fft_lengths(L, M, wl, nfft, ov):
    """Window lengths above the shared-memory kernel's 16384 (up to the reference's 2^16 exploration sizes) and FFT
    lengths that differ from the window (zero-padded or cropped frames, _spectral_methods.py:1180-1215) run on the
    global-memory Stockham path -- against the oracle."""
    rng = np.random.default_rng(wl + (nfft or 0))
    x = rng.standard_normal((L, M)) + 0.1
    for sc, det in [("FFTBackward", False), ("AmplitudeSpectrum", True)]:
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, f, st = _stft(x, 48000, wl, dsp.Window.Hann, ov, nfft, det, True, getattr(dsp.SpectrumScaling, sc))
            to, fo, so = R.stft(x, 48000, wl, "hann", ov, nfft, det, True, sc)
        assert st.shape == so.shape and st.dtype == so.dtype
        assert np.max(np.abs(st - so)) <= REL_TO_MAX * np.max(np.abs(so))
        assert np.array_equal(f, fo) and np.allclose(t, to, rtol=0, atol=1e-15)
