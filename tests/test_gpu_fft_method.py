"""Next row f3 parity: SpectrumMethod.FFT -- `Signal.get_spectrum` (one real FFT over the whole recording of any
length: power of two -> Stockham passes, else Bluestein) and `Signal.get_csm` -> `_csm_fft` (rank-one matrix per bin)
on the GPU against the golden vectors of the real reference and against the oracle."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine  # noqa: E402
from dsptoolbox_b200.standard._spectral_methods import _csm_fft, spectrum_fft_device  # noqa: E402
from oracle import restatement as R  # noqa: E402

# fp64 transforms; Bluestein goes through three FFTs of a longer length and a chirp, the bound is relative to the
# largest bin like the Welch parity bound
REL_TO_MAX = 1e-11


def _cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "fft_method.npz"))
    for i in range(int(g["n_cases"])):
        yield g, i, str(g[f"names_{i}"][0]), bool(g[f"fast_{i}"][0])


def test_golden_spectra_and_csm(golden_dir):
    seen = 0
    for g, i, sc, fast in _cases(golden_dir):
        s = dsp.Signal(None, g[f"x_{i}"].copy(), int(g["fs"]))
        s.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT, pad_to_fast_length=fast,
                                  scaling=getattr(dsp.SpectrumScaling, sc))
        f, sp = s.get_spectrum()
        ref = g[f"sp_{i}"]
        assert sp.shape == ref.shape and sp.dtype == ref.dtype, (i, sp.shape, ref.shape, sp.dtype, ref.dtype)
        assert np.max(np.abs(sp - ref)) <= REL_TO_MAX * np.max(np.abs(ref)), (i, sc)
        assert np.allclose(f, g[f"f_{i}"], rtol=0, atol=1e-9)
        fc, csm = s.get_csm()
        cref = g[f"csm_{i}"]
        assert csm.shape == cref.shape and csm.dtype == cref.dtype
        assert np.max(np.abs(csm - cref)) <= REL_TO_MAX * np.max(np.abs(cref)), (i, sc)
        assert np.allclose(fc, g[f"f_{i}"], rtol=0, atol=1e-9)
        seen += 1
    assert seen == 8


@pytest.mark.parametrize("L,M,fast", [(1 << 20, 3, True), (1000003, 2, False), (480000, 4, True), (65537, 2, False),
                                      (17, 2, False), (2, 2, False), (3 * (1 << 18), 2, True), (100000, 64, True)])
def test_lengths_against_oracle(L, M, fast):
    """Recording lengths from 2 samples to a million: powers of two (radix-16 Stockham passes), 5-smooth lengths and
    primes (Bluestein), many channels -- against the oracle's numpy rfft, fp64 and fp32 input."""
    rng = np.random.default_rng(L)
    x = rng.standard_normal((L, M)) * (1 + np.arange(M)) + 0.5
    for sc in ("FFTBackward", "PowerSpectralDensity", "AmplitudeSpectrum"):
        f, sp = spectrum_fft_device(x, 48000, getattr(dsp.SpectrumScaling, sc), fast)
        fo, so = R.spectrum_fft(x, 48000, sc, fast)
        sp = sp.cpu().numpy()
        assert sp.shape == so.shape and sp.dtype == so.dtype
        assert np.max(np.abs(sp - so)) <= REL_TO_MAX * np.max(np.abs(so)), (sc, np.max(np.abs(sp - so)) / np.max(np.abs(so)))
        assert np.allclose(f, fo, rtol=0, atol=1e-9)
    x32 = x.astype(np.float32)
    _, s32 = spectrum_fft_device(x32, 48000, dsp.SpectrumScaling.FFTBackward, fast)
    _, o32 = R.spectrum_fft(x32, 48000, "FFTBackward", fast)
    assert np.max(np.abs(s32.cpu().numpy() - o32)) <= REL_TO_MAX * np.max(np.abs(o32))


@pytest.mark.parametrize("sc", ["FFTBackward", "FFTForward", "FFTOrthogonal", "AmplitudeSpectrum",
                                "AmplitudeSpectralDensity", "PowerSpectrum", "PowerSpectralDensity"])
def test_csm_fft_function(sc):
    """`_csm_fft` (reference _spectral_methods.py:374-443) as a function: host spectrum in, host matrix out."""
    rng = np.random.default_rng(5)
    sp = rng.standard_normal((1025, 7)) + 1j * rng.standard_normal((1025, 7))
    # real DC / Nyquist bins of both signs: negative real products sit on the branch cut of the square root the
    # amplitude scalings take, and the sign of the zero imaginary part (fixed by HOW the reference builds the matrix)
    # decides between +i and -i
    sp[0] = rng.standard_normal(7)
    sp[-1] = rng.standard_normal(7)
    csm = _csm_fft(sp, getattr(dsp.SpectrumScaling, sc), None, 44100)
    ref = R.csm_fft(sp, sc, 44100)
    assert csm.shape == ref.shape == (1025, 7, 7) and csm.dtype == np.complex128
    assert np.max(np.abs(csm - ref)) <= 1e-14 * np.max(np.abs(ref))
    d = np.einsum("fii->fi", csm)
    assert np.all(d.imag == 0.0)                           # the diagonal is written as |X|^2, exactly real


def test_device_input_spectrum_class_and_errors():
    rng = np.random.default_rng(9)
    x = rng.standard_normal((30000, 3))
    # CUDA tensor input stays on the device
    f, sp = spectrum_fft_device(torch.from_numpy(x).cuda(), 16000, dsp.SpectrumScaling.FFTBackward, True)
    _, so = R.spectrum_fft(x, 16000, "FFTBackward", True)
    assert sp.is_cuda and np.max(np.abs(sp.cpu().numpy() - so)) <= REL_TO_MAX * np.max(np.abs(so))
    # Spectrum.from_signal (reference classes/spectrum.py:56-88)
    s = dsp.Signal(None, x, 16000)
    s.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT, scaling=dsp.SpectrumScaling.AmplitudeSpectrum)
    spec = dsp.Spectrum.from_signal(s, complex=True)
    _, so = R.spectrum_fft(x, 16000, "AmplitudeSpectrum", True)
    assert np.iscomplexobj(spec.spectral_data) and np.max(np.abs(spec.spectral_data - so)) <= REL_TO_MAX * np.max(np.abs(so))
    mag = dsp.Spectrum.from_signal(s, complex=False)
    assert np.max(np.abs(mag.spectral_data - np.abs(so))) <= REL_TO_MAX * np.max(np.abs(so))
    s.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT, scaling=dsp.SpectrumScaling.PowerSpectrum)
    with pytest.raises(AssertionError):
        dsp.Spectrum.from_signal(s, complex=True)
    _, po = R.spectrum_fft(x, 16000, "PowerSpectrum", True)
    mag = dsp.Spectrum.from_signal(s)
    assert np.max(np.abs(mag.spectral_data - po ** 0.5)) <= REL_TO_MAX * np.max(po ** 0.5)
    # smoothing is not on this path; one channel has no cross-spectral matrix
    s.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT, smoothing=3)
    with pytest.raises(NotImplementedError):
        s.get_spectrum()
    one = dsp.Signal(None, x[:, :1], 16000)
    one.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT)
    with pytest.raises(AssertionError):
        one.get_csm()
    n0 = engine.launches
    one.get_spectrum()
    assert engine.launches > n0
