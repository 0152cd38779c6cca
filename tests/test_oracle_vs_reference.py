"""Oracle vs the real reference on fresh random inputs -- only where /root/reference exists
(the build container).  Skipped on the GPU box; the golden vectors carry the pin there."""

import contextlib
import io

import numpy as np
import pytest

from dsptoolbox_b200 import synth
from oracle import ref_loader, restatement as R

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def test_csm_and_map_random():
    dsp = ref_loader.load()
    rng = np.random.default_rng(123)
    pos = dict(x=rng.uniform(-0.3, 0.3, 6), y=rng.uniform(-0.3, 0.3, 6), z=np.zeros(6))
    mic = synth.positions_to_xyz(pos)
    fs = 12000
    x = synth.point_source_recording(mic, 5000, fs, source_xyz=(0.05, 0.1, 0.7), noise_sigma=0.4, seed=9)
    s = dsp.Signal(None, x.copy(), fs)
    s.set_spectrum_parameters(window_length_samples=128, overlap_percent=60,
                              scaling=dsp.SpectrumScaling.AmplitudeSpectrum)
    f, c = s.get_csm()
    fo, co = R.csm_welch(x, fs, 128, "hann", 60, True, "amplitudespectrum")
    assert np.max(np.abs(c - co)) / np.max(np.abs(c)) < 1e-13
    grid = dsp.beamforming.Grid(dict(x=rng.uniform(-0.5, 0.5, 40), y=rng.uniform(-0.5, 0.5, 40), z=np.full(40, 0.7)))
    bf = dsp.beamforming.BeamformerDASFrequency(s, dsp.beamforming.MicArray(pos), grid,
                                                dsp.beamforming.SteeringVector(dsp.beamforming.SteeringVectorType.Inverse))
    with contextlib.redirect_stdout(io.StringIO()):
        m = bf.get_beamformer_map(2500.0, 2, True)
    mo = R.das_map(fo, co, mic, grid.coordinates, 2500.0, 2, True, 343.0, "inverse")
    assert np.max(np.abs(m - mo)) / np.max(m) < 1e-12


def test_steering_vectors():
    dsp = ref_loader.load()
    rng = np.random.default_rng(5)
    pos = dict(x=rng.uniform(-0.3, 0.3, 7), y=rng.uniform(-0.3, 0.3, 7), z=rng.uniform(-0.1, 0.1, 7))
    ma = dsp.beamforming.MicArray(pos)
    g = dsp.beamforming.Regular2DGrid(np.arange(-0.5, 0.5, 0.25), np.arange(-0.5, 0.5, 0.2), ["x", "y"], 1.0)
    k = np.array([1000.0, 1200.0]) * 2 * np.pi / 343
    for name in ("Classic", "Inverse", "TruePower", "TrueLocation"):
        h = dsp.beamforming.SteeringVector(getattr(dsp.beamforming.SteeringVectorType, name)).get_vector(k, g, ma)
        ho = R.steering_vector(k, g.coordinates, ma.coordinates, name)
        assert np.allclose(h, ho, rtol=1e-13, atol=0)


def test_stft_random():
    """Next row f4: oracle `stft` vs the real `_stft` (through Signal.get_spectrogram) on fresh random input."""
    import warnings
    dsp = ref_loader.load()
    rng = np.random.default_rng(123)
    x = rng.standard_normal((2500, 2)) + 0.3
    for wl, win, ov, nfft, det, pad, sc in [(128, "Hann", 60.0, None, True, True, "PowerSpectrum"),
                                             (256, "Hamming", 33.0, None, False, False, "FFTBackward"),
                                             (64, "Hann", 50.0, 128, True, True, "AmplitudeSpectralDensity")]:
        s = dsp.Signal(None, x.copy(), 16000)
        s.set_spectrogram_parameters(window_length_samples=wl, window_type=getattr(dsp.Window, win), overlap_percent=ov,
                                     fft_length_samples=nfft, detrend=det, padding=pad,
                                     scaling=getattr(dsp.SpectrumScaling, sc))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, f, st = s.get_spectrogram()
        to, fo, so = R.stft(x, 16000, wl, win.lower(), ov, nfft, det, pad, sc)
        assert so.shape == st.shape and np.max(np.abs(so - st)) <= 1e-13 * np.max(np.abs(st))
        assert np.allclose(to, t, rtol=0, atol=1e-15) and np.array_equal(fo, f)
