"""Parity at BASELINE.json configs[1] and [2]: 64 microphones x 10 s @ 48 kHz, Welch nfft 4096 / hann / 50 %
(CSM, compared in FULL with the oracle -- it fits the CPU at this size) and the 640-bin band 500 Hz - 8 kHz of
the 200 x 200 x 20 grid (a 40 000-point sub-sample: 313 tiles > 148 CTAs, i.e. bin-range work units; seeded
points recomputed by the oracle).  This is the `<64, *>` instantiation of the map kernel at its real shape."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine, synth  # noqa: E402
from oracle import restatement as R  # noqa: E402
from parity import assert_csm_close  # noqa: E402

FS, WL, M, L = 48_000, 4096, 64, 10 * 48_000
FC, FRAC = 2000.0, 0.25


@pytest.fixture(scope="module")
def setup():
    pos = synth.disc_array(M, 0.75, seed=7)
    mic = synth.positions_to_xyz(pos)
    x = synth.point_source_recording(mic, L, FS, noise_sigma=0.1, seed=1)
    engine.config.csm_precision = "fp64"
    sig = dsp.Signal(None, x, FS, activate_cache=True).set_spectrum_parameters(window_length_samples=WL)
    return pos, mic, x, sig


def test_cfg2_csm_full_against_oracle(setup):
    pos, mic, x, sig = setup
    f, c = sig.get_csm()
    assert c.shape == (WL // 2 + 1, M, M)
    fo, co = R.csm_welch(x, FS, WL, "hann", 50, True, "fftbackward")
    assert np.array_equal(f, fo)
    e = assert_csm_close(c, co)
    assert e["max_rel"] <= 1e-9          # strict mode


@pytest.mark.parametrize("mode", ["f16x3", "mixed", "tf32x3", "simt_f32"])
def test_cfg3_map_subsample_against_oracle(setup, mode):
    pos, mic, x, sig = setup
    if mode != "simt_f32" and not engine.tc_map_available():
        pytest.skip("tensor-core map kernel not built")
    f, c = sig.get_csm_device()
    id1, id2 = R.band_bins(f, FC, FRAC)
    assert (id1, id2) == (43, 683)       # 503.9 - 7992.2 Hz, SURVEY.md section 8a
    rng = np.random.default_rng(2)
    lines = (np.linspace(-1, 1, 200), np.linspace(-1, 1, 200), np.linspace(0.5, 1.5, 20))
    n = 40_000 if mode != "simt_f32" else 4_000
    idx = rng.choice(800_000, size=n, replace=False)
    ix, iy, iz = np.unravel_index(idx, (200, 200, 20))
    gxyz = np.stack([lines[0][ix], lines[1][iy], lines[2][iz]], axis=1)
    grid = dsp.beamforming.Grid(dict(x=gxyz[:, 0], y=gxyz[:, 1], z=gxyz[:, 2]))
    engine.config.map_mode = mode
    try:
        import contextlib
        import io

        bf = dsp.beamforming.BeamformerDASFrequency(sig, dsp.beamforming.MicArray(pos), grid, dsp.beamforming.SteeringVector())
        with contextlib.redirect_stdout(io.StringIO()):
            m = bf.get_beamformer_map(FC, FRAC, True)
    finally:
        engine.config.map_mode = "auto"
    pick = np.unique(np.concatenate([rng.choice(n, size=48, replace=False), [int(np.argmax(m))]]))
    ref = R.das_map(f, c.cpu().numpy(), mic, gxyz[pick], FC, FRAC, True, 343.0, "truelocation", chunk_points=16)
    mx = m.max()
    assert np.max(np.abs(m[pick] - ref)) <= 1e-4 * mx
    sel = ref >= mx * 1e-2
    assert sel.any() and np.max(np.abs(10 * np.log10(m[pick][sel] / ref[sel]))) <= 0.01
    assert np.all(m >= 0) and np.all(np.isfinite(m))
