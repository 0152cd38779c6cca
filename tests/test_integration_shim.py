"""INTEGRATION.md section 2 (`dsptoolbox_b200/integration.py`) executed against the real reference package:
`install()` must re-bind `_csm_welch` where `classes/signal.py:39` looks it up and replace
`BeamformerDASFrequency.get_beamformer_map`, so that the reference's OWN `Signal` / beamformer objects run on
the B200 path.  CPU: the engine is stood in for by the oracle (wiring, argument translation, side effects).
GPU: the real kernels, compared with what the unpatched reference computes."""

import contextlib
import io

import numpy as np
import pytest

from dsptoolbox_b200 import integration, synth
from oracle import ref_loader, restatement as R

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference package not present")


class _FakeTensor:
    def __init__(self, a):
        self.a = a

    def cpu(self):
        return self

    def numpy(self):
        return self.a


def _setup(ref, seed=3):
    rng = np.random.default_rng(seed)
    pos = dict(x=rng.uniform(-0.3, 0.3, 6), y=rng.uniform(-0.3, 0.3, 6), z=np.zeros(6))
    mic = synth.positions_to_xyz(pos)
    x = synth.point_source_recording(mic, 6000, 12000, source_xyz=(0.05, 0.1, 0.7), noise_sigma=0.3, seed=seed)
    sig = ref.Signal(None, x.copy(), 12000)
    sig.set_spectrum_parameters(window_length_samples=256, overlap_percent=50, window_type=ref.Window.Hamming,
                                scaling=ref.SpectrumScaling.PowerSpectralDensity)
    grid = ref.beamforming.Regular2DGrid(np.linspace(-0.4, 0.4, 7), np.linspace(-0.4, 0.4, 5), ("x", "y"), 0.7)
    bf = ref.beamforming.BeamformerDASFrequency(sig, ref.beamforming.MicArray(pos), grid,
                                                ref.beamforming.SteeringVector(ref.beamforming.SteeringVectorType.TruePower))
    return x, mic, sig, grid, bf


def test_install_rebinds_and_translates_arguments(monkeypatch):
    ref = ref_loader.load()
    x, mic, sig, grid, bf = _setup(ref)
    with contextlib.redirect_stdout(io.StringIO()):
        want_f, want_csm = sig.get_csm()
        want_map = bf.get_beamformer_map(2500.0, 3, True)
    seen = {}

    def fake_csm_welch_device(td, fs, wl, window, overlap, detrend, average, scaling, **kw):
        seen["csm"] = (wl, window, overlap, detrend, average, scaling.name)
        f, c = R.csm_welch(td, fs, wl, window, overlap, detrend, scaling.name)
        return f, _FakeTensor(c)

    def fake_prepare(csm, bin0, nbins, remove_diag, mode):
        seen["prepare"] = (bin0, nbins, remove_diag)
        return csm

    def fake_map(op, M, freqs, weights, c, mic_xyz, center, grid_xyz, steering, clamp, mode, per_bin=False):
        seen["map"] = (M, steering, clamp)
        form = {0: "classic", 1: "inverse", 2: "truepower", 3: "truelocation"}[steering]
        full_f = np.fft.rfftfreq(256, 1 / 12000)
        P = R.das_map(full_f, op.a, mic_xyz, grid_xyz, 2500.0, 3, seen["prepare"][2], c, form, return_per_bin=True)
        return _FakeTensor(P @ weights), None

    monkeypatch.setattr(integration.b2sm, "csm_welch_device", fake_csm_welch_device)
    monkeypatch.setattr(integration.engine, "resolve_map_mode", lambda mode=None: 5)
    monkeypatch.setattr(integration.engine, "das_prepare", fake_prepare)
    monkeypatch.setattr(integration.engine, "das_map", fake_map)
    try:
        integration.install(ref)
        assert ref.classes.signal._csm_welch is integration.csm_welch_b200          # the signal.py:39 binding
        assert ref.standard._spectral_methods._csm_welch is integration.csm_welch_b200
        f, csm = sig.get_csm(force_computation=True)                                  # reference class, patched backend
        assert seen["csm"] == (256, "hamming", 50, True, "mean", "PowerSpectralDensity")
        assert np.allclose(f, want_f) and np.max(np.abs(csm - want_csm)) <= 1e-12 * np.abs(want_csm).max()
        with contextlib.redirect_stdout(io.StringIO()) as out:
            m = bf.get_beamformer_map(2500.0, 3, True)
        assert "...csm..." in out.getvalue() and "...Apply..." in out.getvalue()
        assert seen["map"] == (6, 2, True) and m.shape == want_map.shape == (7, 5)
        assert np.max(np.abs(m - want_map)) <= 1e-10 * want_map.max()
        assert np.allclose(bf.f_range_hz, [f[seen["prepare"][0]], f[seen["prepare"][0] + seen["prepare"][1] - 1]])
        assert bf.center_frequency_hz == 2500.0 and bf.octave_fraction == 3 and np.array_equal(bf.map, m)
    finally:
        integration.uninstall()
    assert ref.classes.signal._csm_welch is not integration.csm_welch_b200
    assert ref.beamforming.beamforming.BeamformerDASFrequency.get_beamformer_map is not None


@pytest.mark.gpu
def test_reference_objects_run_on_the_gpu():
    """The reference's own Signal / BeamformerDASFrequency, patched, against the unpatched reference."""
    from parity import assert_csm_close, assert_map_close

    ref = ref_loader.load()
    x, mic, sig, grid, bf = _setup(ref, seed=8)
    with contextlib.redirect_stdout(io.StringIO()):
        want_f, want_csm = sig.get_csm()
        want_map = bf.get_beamformer_map(2500.0, 3, True)
    try:
        integration.install(ref)
        with contextlib.redirect_stdout(io.StringIO()):
            f, csm = sig.get_csm(force_computation=True)
            m = bf.get_beamformer_map(2500.0, 3, True)
    finally:
        integration.uninstall()
    assert np.allclose(f, want_f)
    assert_csm_close(csm, want_csm)
    assert_map_close(m, want_map)
