"""The oracle (numpy restatement) against the golden vectors produced by the REAL reference
(oracle/make_golden.py).  CPU only."""

import os

import numpy as np
import pytest

from dsptoolbox_b200 import synth
from oracle import restatement as R

SCALINGS = ["AmplitudeSpectrum", "AmplitudeSpectralDensity", "PowerSpectrum", "PowerSpectralDensity",
            "FFTBackward", "FFTForward", "FFTOrthogonal"]


def rel_to_max(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


@pytest.mark.parametrize("scaling", SCALINGS)
def test_csm_all_scalings(golden_dir, scaling):
    g = np.load(os.path.join(golden_dir, "csm_small_scalings.npz"))
    f, c = R.csm_welch(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, scaling)
    assert np.array_equal(f, g["f"])
    assert rel_to_max(c, g["csm_" + scaling]) < 1e-13
    _, cp = R.csm_welch_pairwise(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, scaling)
    assert rel_to_max(cp, g["csm_" + scaling]) < 1e-13


VARIANTS = [("ov75_hamming", 64, 75, True, "hamming"), ("ov0_boxcar_nodetrend", 256, 0, False, "boxcar"),
            ("ov33_blackman_Lmultiple", 128, 33, True, "blackman"), ("ov50_kaiser8_wl8", 8, 50, True, ("kaiser", 8.0)),
            ("ov90_flattop_short", 512, 90, True, "flattop")]


@pytest.mark.parametrize("tag,wl,ov,detrend,window", VARIANTS)
def test_csm_variants(golden_dir, tag, wl, ov, detrend, window):
    g = np.load(os.path.join(golden_dir, "csm_variants.npz"))
    _, c = R.csm_welch(g["x_" + tag], int(g["fs"]), wl, window, ov, detrend, "PowerSpectralDensity")
    assert rel_to_max(c, g["csm_" + tag]) < 1e-13


MAP_CASES = [(2000.0, 3, True, "third_rd"), (2000.0, 3, False, "third_full"), (3000.0, 0, True, "single_rd"),
             (1500.0, 1, True, "octave_rd"), (2500.0, 12, False, "twelfth_full")]


@pytest.mark.parametrize("form", ["Classic", "Inverse", "TruePower", "TrueLocation"])
@pytest.mark.parametrize("fc,frac,rd,tag", MAP_CASES)
def test_das_small(golden_dir, form, fc, frac, rd, tag):
    g = np.load(os.path.join(golden_dir, "das_small.npz"))
    mic = np.stack([g["mic_x"], g["mic_y"], g["mic_z"]], axis=1)
    a, b = np.meshgrid(g["line1"], g["line2"], indexing="ij")
    grid = np.stack([a.ravel(), b.ravel(), np.full(a.size, float(g["value3"]))], axis=1)
    f, c = R.csm_welch(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, "fftbackward")
    ref = g[f"map_{form}_{tag}"]
    m = R.das_map(f, c, mic, grid, fc, frac, rd, 343.0, form).reshape(ref.shape)
    assert rel_to_max(m, ref) < 1e-12
    ml = R.das_map_loop(f, c, mic, grid, fc, frac, rd, 343.0, form, grid_subset=np.arange(0, 63, 7))
    assert rel_to_max(ml, ref.ravel()[::7]) < 1e-12
    id1, id2 = R.band_bins(f, fc, frac)
    assert np.allclose(g[f"frange_{form}_{tag}"], [f[id1], f[id2 - 1]])


def test_config1(golden_dir):
    """BASELINE.json configs[0]: array.xml geometry, 1 s @ 48 kHz, nfft 1024, 50x50 grid, 4 kHz third octave."""
    g = np.load(os.path.join(golden_dir, "cfg1.npz"))
    arr = np.load(os.path.join(golden_dir, "array_xml_positions.npz"))
    mic = synth.positions_to_xyz(arr)
    fs = int(g["fs"])
    x = synth.point_source_recording(mic, fs, fs, noise_sigma=float(g["noise_sigma"]), seed=int(g["seed"]))
    fp = np.array([x.sum(), np.abs(x).sum(), x[12345, 7], x[-1, -1]])
    assert np.allclose(fp, g["x_fingerprint"], rtol=1e-12), "synthetic input is not reproducible on this numpy"
    f, c = R.csm_welch(x, fs, int(g["wl"]), "hann", 50, True, "fftbackward")
    idx = g["csm_idx"]
    got = c[idx[:, 0], idx[:, 1], idx[:, 2]]
    assert np.max(np.abs(got - g["csm_vals"]) / np.abs(g["csm_vals"])) < 1e-9
    assert rel_to_max(c[g["csm_bins"]], g["csm_full_bins"]) < 1e-13
    l = np.linspace(-0.5, 0.5, 50)
    a, b = np.meshgrid(l, l, indexing="ij")
    grid = np.stack([a.ravel(), b.ravel(), np.ones(a.size)], axis=1)
    m = R.das_map(f, c, mic, grid, 4000.0, 3, True).reshape(50, 50)
    assert rel_to_max(m, g["map"]) < 1e-12
    assert np.unravel_index(np.argmax(m), m.shape) == (29, 15)       # grid node nearest (0.1, -0.2)


def test_simpson_weights_match_scipy():
    from scipy.integrate import simpson

    rng = np.random.default_rng(0)
    for n in (2, 3, 4, 7, 20, 21):
        y = rng.standard_normal((5, n))
        assert np.allclose(y @ R.simpson_weights(n, 0.3), simpson(y, dx=0.3, axis=1), rtol=1e-13)


@pytest.mark.parametrize("form", ["Classic", "TrueLocation"])
def test_mvdr_functional_next_rows(golden_dir, form):
    """oracle restatements of BeamformerMVDR / BeamformerFunctional vs the reference."""
    g = np.load(os.path.join(golden_dir, "das_small.npz"))
    ge = np.load(os.path.join(golden_dir, "beamformers_next.npz"))
    mic = np.stack([g["mic_x"], g["mic_y"], g["mic_z"]], axis=1)
    a, b = np.meshgrid(g["line1"], g["line2"], indexing="ij")
    grid = np.stack([a.ravel(), b.ravel(), np.full(a.size, float(g["value3"]))], axis=1)
    f, c = R.csm_welch(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, "fftbackward")
    for fc, frac, tag in [(2000.0, 3, "third"), (3000.0, 0, "single")]:
        assert rel_to_max(R.mvdr_map(f, c, mic, grid, fc, frac, 343.0, form), ge[f"mvdr_{form}_{tag}"].ravel()) < 1e-9
        assert rel_to_max(R.functional_map(f, c, mic, grid, fc, frac, 10, 343.0, form), ge[f"functional_{form}_{tag}"].ravel()) < 1e-9
        assert rel_to_max(R.functional_map(f, c, mic, grid, fc, frac, 3, 343.0, form), ge[f"functional3_{form}_{tag}"].ravel()) < 1e-9


@pytest.mark.parametrize("form", ["Classic", "TrueLocation"])
def test_orthogonal_next_row(golden_dir, form):
    g = np.load(os.path.join(golden_dir, "das_small.npz"))
    ge = np.load(os.path.join(golden_dir, "beamformers_next.npz"))
    mic = np.stack([g["mic_x"], g["mic_y"], g["mic_z"]], axis=1)
    a, b = np.meshgrid(g["line1"], g["line2"], indexing="ij")
    grid = np.stack([a.ravel(), b.ravel(), np.full(a.size, float(g["value3"]))], axis=1)
    f, c = R.csm_welch(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, "fftbackward")
    for fc, frac, tag in [(2000.0, 3, "third"), (3000.0, 0, "single")]:
        assert rel_to_max(R.orthogonal_map(f, c, mic, grid, fc, frac, None, 343.0, form), ge[f"orthogonal_{form}_{tag}"].ravel()) < 1e-9
        assert rel_to_max(R.orthogonal_map(f, c, mic, grid, fc, frac, 3, 343.0, form), ge[f"orthogonal3_{form}_{tag}"].ravel()) < 1e-9


@pytest.mark.parametrize("form", ["Classic", "TrueLocation"])
def test_cleansc_next_row(golden_dir, form):
    g = np.load(os.path.join(golden_dir, "das_small.npz"))
    ge = np.load(os.path.join(golden_dir, "beamformers_next.npz"))
    mic = np.stack([g["mic_x"], g["mic_y"], g["mic_z"]], axis=1)
    a, b = np.meshgrid(g["line1"], g["line2"], indexing="ij")
    grid = np.stack([a.ravel(), b.ravel(), np.full(a.size, float(g["value3"]))], axis=1)
    f, c = R.csm_welch(g["x"], int(g["fs"]), int(g["wl"]), "hann", 50, True, "fftbackward")
    for fc, frac, tag in [(2000.0, 3, "third"), (3000.0, 0, "single")]:
        assert rel_to_max(R.clean_sc_map(f, c, mic, grid, fc, frac, None, 0.5, False, 343.0, form), ge[f"cleansc_{form}_{tag}"].ravel()) < 1e-9
        assert rel_to_max(R.clean_sc_map(f, c, mic, grid, fc, frac, 5, 0.8, True, 343.0, form), ge[f"cleansc_rd_{form}_{tag}"].ravel()) < 1e-9


@pytest.mark.parametrize("scaling", SCALINGS)
def test_get_spectrum_welch(golden_dir, scaling):
    g = np.load(os.path.join(golden_dir, "csm_small_scalings.npz"))
    x = g["x"]
    so = np.stack([R.welch_pair(x[:, i], None, int(g["fs"]), "hann", int(g["wl"]), 50, True, scaling) for i in range(x.shape[1])], axis=1)
    assert rel_to_max(so, g["spectrum_" + scaling]) < 1e-13


def _stft_cases(golden_dir):
    g = np.load(os.path.join(golden_dir, "stft_small.npz"))
    for i in range(int(g["n_cases"])):
        wl, ov, nfft, det, pad = g[f"params_{i}"]
        win, sc = (str(v) for v in g[f"names_{i}"])
        yield g, i, int(wl), win.lower(), float(ov), (None if nfft < 0 else int(nfft)), bool(det), bool(pad), sc


def test_stft_next_row(golden_dir):
    """Next row f4: the oracle's `stft` against the reference's Signal.get_spectrogram for every scaling family,
    padding / detrend on and off, rounded overlaps, and FFT lengths that zero-pad or crop the frames."""
    n = 0
    for g, i, wl, win, ov, nfft, det, pad, sc in _stft_cases(golden_dir):
        t, f, st = R.stft(g["x"], int(g["fs"]), wl, win, ov, nfft, det, pad, sc)
        ref = g[f"stft_{i}"]
        assert st.shape == ref.shape and st.dtype == ref.dtype
        assert np.max(np.abs(st - ref)) <= 1e-13 * np.max(np.abs(ref))
        assert np.array_equal(f, g[f"f_{i}"]) and np.allclose(t, g[f"t_{i}"], rtol=0, atol=1e-15)
        n += 1
    assert n == 9


def test_transfer_functions_next_row(golden_dir):
    """Next row f3: the oracle's H1 / H2 / H3 + coherence against the reference's compute_transfer_function."""
    g = np.load(os.path.join(golden_dir, "transfer_functions.npz"))
    for i in range(int(g["n_cases"])):
        shared, wl, ov, det = g[f"params_{i}"]
        mode, sc = (str(v) for v in g[f"names_{i}"])
        y = g["y_shared"] if shared else g["y_out"]
        x = g["x_in"][:, :1] if shared else g["x_in"]
        f, tf, coh = R.transfer_function(y, x, int(g["fs"]), int(wl), "hann", float(ov), bool(det), sc, mode)
        assert np.array_equal(f, g[f"f_{i}"])
        assert np.max(np.abs(tf - g[f"tf_{i}"])) <= 1e-12 * np.max(np.abs(g[f"tf_{i}"]))
        assert np.max(np.abs(coh - g[f"coh_{i}"])) <= 1e-12 * np.max(np.abs(g[f"coh_{i}"]))


def test_fft_method_next_row(golden_dir):
    """Next row f3: the oracle's single-FFT spectrum and `_csm_fft` restatements against the reference's
    Signal.get_spectrum / get_csm with SpectrumMethod.FFT, every scaling, padded-to-fast and exact lengths."""
    g = np.load(os.path.join(golden_dir, "fft_method.npz"))
    for i in range(int(g["n_cases"])):
        sc, fast = str(g[f"names_{i}"][0]), bool(g[f"fast_{i}"][0])
        f, sp = R.spectrum_fft(g[f"x_{i}"], int(g["fs"]), sc, fast)
        assert sp.dtype == g[f"sp_{i}"].dtype and np.allclose(f, g[f"f_{i}"], rtol=0, atol=1e-9)
        assert np.max(np.abs(sp - g[f"sp_{i}"])) <= 1e-13 * np.max(np.abs(g[f"sp_{i}"]))
        _, spb = R.spectrum_fft(g[f"x_{i}"], int(g["fs"]), "fftbackward", fast)
        csm = R.csm_fft(spb, sc, int(g["fs"]))
        assert np.max(np.abs(csm - g[f"csm_{i}"])) <= 1e-13 * np.max(np.abs(g[f"csm_{i}"]))


def test_sources_next_row(golden_dir):
    """Next row f2: the oracle's fractional-delay source simulation against the reference's MonopoleSource /
    mix_sources_on_array, including a source on a microphone (zero delay) and one 2 cm away (negative shift)."""
    g = np.load(os.path.join(golden_dir, "sources_on_array.npz"))
    fs, mic, src = int(g["fs"]), g["mic"], g["src"]
    for i in range(int(g["n_src"])):
        got = R.monopole_on_array(g[f"x_{i}"], fs, src[i], mic, float(g["c"]))
        assert np.max(np.abs(got - g[f"on_array_{i}"])) <= 1e-13 * np.max(np.abs(g[f"on_array_{i}"]))
    for tag in ("equal", "ragged", "ragged2"):
        ids = g[f"mix_{tag}_ids"]
        got = R.mix_sources_on_array([g[f"x_{i}"] for i in ids], [src[i] for i in ids], fs, mic, float(g["c"]))
        assert got.shape == g[f"mix_{tag}"].shape
        assert np.max(np.abs(got - g[f"mix_{tag}"])) <= 1e-13 * np.max(np.abs(g[f"mix_{tag}"]))
    assert R.fractional_delay_filter(0.93)[0] == -14 and R.fractional_delay_filter(140.2)[0] == 125
