"""Next row f3 parity: compute_transfer_function (H1 / H2 / H3 + coherence) on the GPU -- one Welch CSM of the
stacked [input | output] channels -- against the golden vectors of the real reference and the oracle."""

import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine  # noqa: E402
from dsptoolbox_b200.transfer_functions import TransferFunctionType, compute_transfer_function  # noqa: E402
from oracle import restatement as R  # noqa: E402

RTOL = 1e-7     # element-wise; strict fp64 stage 1 gives ~1e-12


@pytest.fixture(autouse=True)
def _strict():
    engine.config.csm_precision = "fp64"
    yield
    engine.config.csm_precision = "fp64"


def _close(got, ref, detrend):
    # with de-meaned frames bin 0 of every spectrum is rounding noise in the reference (exact 0 here), and the
    # estimators divide noise by noise there: compared from bin 1 on, like the CSM contract (parity.py)
    s = slice(1, None) if detrend else slice(None)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    err = np.abs(got - ref)
    # At DC and Nyquist every cross-spectrum is exactly real.  Where it is negative, the amplitude scalings take
    # the complex sqrt ON the branch cut and the sign of the result follows the sign of a zero imaginary part
    # (+0 / -0 left over from conj() and the frame sum) -- data-dependent in the reference itself.  Those two
    # bins are therefore compared up to conjugation.
    for edge in (0, -1):
        err[edge] = np.minimum(err[edge], np.abs(got[edge] - np.conj(ref[edge])))
    assert np.all(err[s] <= RTOL * np.abs(ref[s]) + 1e-12 * np.max(np.abs(ref[s])))


def test_golden_transfer_functions(golden_dir):
    g = np.load(os.path.join(golden_dir, "transfer_functions.npz"))
    for i in range(int(g["n_cases"])):
        shared, wl, ov, det = g[f"params_{i}"]
        mode, sc = (str(v) for v in g[f"names_{i}"])
        sin = dsp.Signal(None, (g["x_in"][:, :1] if shared else g["x_in"]).copy(), int(g["fs"]))
        sin.set_spectrum_parameters(overlap_percent=float(ov), detrend=bool(det), scaling=getattr(dsp.SpectrumScaling, sc))
        sout = dsp.Signal(None, (g["y_shared"] if shared else g["y_out"]).copy(), int(g["fs"]))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = compute_transfer_function(sout, sin, int(wl), getattr(TransferFunctionType, mode))
        assert isinstance(spec, dsp.Spectrum) and np.array_equal(spec.frequency_vector_hz, g[f"f_{i}"])
        _close(spec.spectral_data, g[f"tf_{i}"], bool(det))
        _close(spec.coherence, g[f"coh_{i}"], bool(det))
        assert spec.number_of_channels == 3
        # the DC bin of de-meaned frames holds no energy: defined as 0 here (0 / 0), never NaN
        assert np.all(np.isfinite(spec.spectral_data)) and np.all(np.isfinite(spec.coherence))
        if det:
            assert np.all(spec.spectral_data[0] == 0) and np.all(spec.coherence[0] == 0)


@pytest.mark.parametrize("mode", ["H1", "H2", "H3"])
def test_against_oracle_and_physics(mode):
    """A known FIR between input and output: the estimate follows the oracle, recovers the filter's frequency
    response where the coherence is high, and the coherence stays in [0, 1]."""
    rng = np.random.default_rng(5)
    fs, L, wl = 48000, 60000, 1024
    x = rng.standard_normal((L, 2))
    h = np.array([0.5, -0.3, 0.2, 0.1])
    y = np.stack([np.convolve(x[:, c], h)[:L] for c in range(2)], 1) + 0.01 * rng.standard_normal((L, 2))
    sin, sout = dsp.Signal(None, x, fs), dsp.Signal(None, y, fs)
    sin.set_spectrum_parameters(detrend=False, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
    spec = compute_transfer_function(sout, sin, wl, getattr(TransferFunctionType, mode))
    f, tf, coh = R.transfer_function(y, x, fs, wl, "hann", 50, False, "PowerSpectralDensity", mode)
    _close(spec.spectral_data, tf, False)
    _close(spec.coherence, coh, False)
    H = np.fft.rfft(h, wl)
    assert np.max(np.abs(spec.spectral_data[:, 0] - H)) < 0.05 * np.max(np.abs(H))
    assert np.all(spec.coherence <= 1 + 1e-9) and np.all(spec.coherence >= 0) and np.median(spec.coherence) > 0.99


def test_argument_checks():
    x = np.random.default_rng(0).standard_normal((4000, 2))
    a, b = dsp.Signal(None, x, 8000), dsp.Signal(None, x[:3000], 8000)
    with pytest.raises(AssertionError):
        compute_transfer_function(b, a, 256)                                    # lengths differ
    with pytest.raises(AssertionError):
        compute_transfer_function(dsp.Signal(None, x, 16000), a, 256)           # sampling rates differ
    with pytest.raises(AssertionError):
        compute_transfer_function(dsp.Signal(None, np.c_[x, x[:, :1]], 8000), a, 256)   # 3 outputs, 2 inputs
    with pytest.raises(ValueError):
        compute_transfer_function(a, a, 256, "H1")
