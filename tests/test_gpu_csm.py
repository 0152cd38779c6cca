"""Stage 1 parity: Signal.get_csm / _csm_welch / _welch on the GPU against the golden vectors of
the real reference and against the oracle.  Tolerance: rtol 1e-5 element-wise (north star)."""

import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine, synth  # noqa: E402
from dsptoolbox_b200.standard._spectral_methods import _csm_welch, _welch  # noqa: E402
from oracle import restatement as R  # noqa: E402
from parity import assert_csm_close, csm_errors  # noqa: E402

SCALINGS = ["AmplitudeSpectrum", "AmplitudeSpectralDensity", "PowerSpectrum", "PowerSpectralDensity",
            "FFTBackward", "FFTForward", "FFTOrthogonal"]


@pytest.fixture(autouse=True)
def _strict():
    engine.config.csm_precision = "fp64"
    yield
    engine.config.csm_precision = "fp64"


@pytest.mark.parametrize("scaling", SCALINGS)
def test_golden_all_scalings(golden_dir, scaling):
    g = np.load(os.path.join(golden_dir, "csm_small_scalings.npz"))
    s = dsp.Signal(None, g["x"].copy(), int(g["fs"]))
    s.set_spectrum_parameters(window_length_samples=int(g["wl"]), scaling=getattr(dsp.SpectrumScaling, scaling))
    f, c = s.get_csm()
    assert c.dtype == np.complex128 and c.shape == g["csm_" + scaling].shape and f.dtype == np.float64
    assert np.array_equal(f, g["f"])
    e = assert_csm_close(c, g["csm_" + scaling])
    assert e["max_rel"] < 1e-9                      # strict mode is far inside the 1e-5 contract
    assert np.allclose(c, np.swapaxes(c, 1, 2).conj())          # Hermitian by construction


VARIANTS = [("ov75_hamming", 64, 75, True, dsp.Window.Hamming), ("ov0_boxcar_nodetrend", 256, 0, False, dsp.Window.Boxcar),
            ("ov33_blackman_Lmultiple", 128, 33, True, dsp.Window.Blackman),
            ("ov50_kaiser8_wl8", 8, 50, True, dsp.Window.Kaiser),
            ("ov90_flattop_short", 512, 90, True, dsp.Window.Flattop)]


@pytest.mark.parametrize("tag,wl,ov,detrend,window", VARIANTS)
def test_golden_variants(golden_dir, tag, wl, ov, detrend, window):
    g = np.load(os.path.join(golden_dir, "csm_variants.npz"))
    if window is dsp.Window.Kaiser:
        window = window.with_extra_parameter(8.0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, c = _csm_welch(g["x_" + tag], int(g["fs"]), wl, window, ov, detrend, "mean", dsp.SpectrumScaling.PowerSpectralDensity)
    assert_csm_close(c, g["csm_" + tag], skip_dc=detrend)


def test_config1_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "cfg1.npz"))
    mic = synth.positions_to_xyz(np.load(os.path.join(golden_dir, "array_xml_positions.npz")))
    fs = int(g["fs"])
    x = synth.point_source_recording(mic, fs, fs, noise_sigma=float(g["noise_sigma"]), seed=int(g["seed"]))
    s = dsp.Signal(None, x, fs)
    s.set_spectrum_parameters(window_length_samples=int(g["wl"]))
    f, c = s.get_csm()
    idx = g["csm_idx"]
    got = c[idx[:, 0], idx[:, 1], idx[:, 2]]
    assert np.max(np.abs(got - g["csm_vals"]) / np.abs(g["csm_vals"])) <= 1e-5
    bins = g["csm_bins"]
    assert np.max(np.abs(c[bins] - g["csm_full_bins"]) / np.abs(g["csm_full_bins"])) <= 1e-5
    assert np.allclose(np.diagonal(c[85]), g["csm_diag_bin85"], rtol=1e-9)


@pytest.mark.parametrize("M,L,wl,ov", [(2, 700, 8, 50), (3, 5000, 32, 40), (7, 4097, 256, 50), (64, 12000, 512, 66),
                                        (70, 9000, 1024, 50), (129, 6000, 2048, 25), (5, 70000, 16384, 50)])
def test_against_oracle_shapes(M, L, wl, ov):
    """Ragged sizes: channel counts off the 64-block, lengths off the hop, tiny and huge windows."""
    rng = np.random.default_rng(M * 1000 + wl)
    x = rng.standard_normal((L, M)) + 0.5 * rng.standard_normal((L, 1))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        f, c = _csm_welch(x, 44100, wl, dsp.Window.Hann, ov, True, "mean", dsp.SpectrumScaling.FFTBackward)
        fo, co = R.csm_welch(x, 44100, wl, "hann", ov, True, "fftbackward")
    assert np.array_equal(f, fo)
    assert_csm_close(c, co)


def test_welch_entry_point():
    rng = np.random.default_rng(8)
    x, y = rng.standard_normal(6000), rng.standard_normal(6000)
    y += 0.5 * x
    for sc in (dsp.SpectrumScaling.PowerSpectralDensity, dsp.SpectrumScaling.FFTBackward):
        cross = _welch(x, y, 8000, dsp.Window.Hann, 256, 50, True, "mean", sc)
        auto = _welch(x, None, 8000, dsp.Window.Hann, 256, 50, True, "mean", sc)
        ref_c = R.welch_pair(x, y, 8000, "hann", 256, 50, True, sc.name)
        ref_a = R.welch_pair(x, None, 8000, "hann", 256, 50, True, sc.name)
        assert np.iscomplexobj(cross) and not np.iscomplexobj(auto)
        assert np.max(np.abs(cross[1:] - ref_c[1:]) / np.abs(ref_c[1:])) < 1e-9
        assert np.max(np.abs(auto[1:] - ref_a[1:]) / np.abs(ref_a[1:])) < 1e-9
    multi = _welch(np.stack([x, y], 1), np.stack([y, x], 1), 8000, dsp.Window.Hann, 256, 50, True, "mean",
                   dsp.SpectrumScaling.PowerSpectrum)
    assert multi.shape == (129, 2)
    assert np.allclose(multi[1:, 0], np.conj(multi[1:, 1]), rtol=1e-9)
    with pytest.raises(AssertionError):
        _welch(x, y, 8000, dsp.Window.Hann, 100, 50, True, "mean", sc)             # not a power of two
    with pytest.raises(AssertionError):
        _welch(x, y, 8000, dsp.Window.Hann, 256, 100, True, "mean", sc)
    with pytest.raises(NotImplementedError):
        _welch(x, y, 8000, dsp.Window.Hann, 256, 50, True, "median", sc)
    with pytest.raises(NotImplementedError):
        _welch(x, y, 8000, dsp.Window.Hann, 2 ** 16, 50, True, "mean", sc)         # beyond the smem FFT


def test_input_layouts_and_dtypes():
    rng = np.random.default_rng(4)
    x = rng.standard_normal((9000, 6))
    ref = R.csm_welch(x, 8000, 256, "hann", 50, True, "fftbackward")[1]
    for arr in (x, np.ascontiguousarray(x.T), np.asfortranarray(x), x[::1, :]):
        c = dsp.Signal(None, arr, 8000).set_spectrum_parameters(window_length_samples=256).get_csm()[1]
        assert_csm_close(c, ref)
    c32 = dsp.Signal(None, x.astype(np.float32), 8000).set_spectrum_parameters(window_length_samples=256).get_csm()[1]
    assert_csm_close(c32, R.csm_welch(x.astype(np.float32), 8000, 256, "hann", 50, True, "fftbackward")[1])
    xd = torch.from_numpy(x).cuda()                       # already-resident recording
    from dsptoolbox_b200.standard._spectral_methods import csm_welch_device
    cd = csm_welch_device(xd, 8000, 256, dsp.Window.Hann, 50, True, "mean", dsp.SpectrumScaling.FFTBackward)[1]
    assert_csm_close(cd.cpu().numpy(), ref)


def test_pageable_recordings_go_through_the_pinned_staging_ring():
    """Large pageable numpy recordings are uploaded through `engine._Stager` (threads -> pinned slots -> async
    copies): bit-identical device data, several ring rounds, row- and channel-major storage, ragged sizes."""
    rng = np.random.default_rng(8)
    old = engine.config.h2d_stage_min_bytes, engine._Stager.SLOT_BYTES
    try:
        engine.config.h2d_stage_min_bytes = 1 << 20
        engine._Stager.SLOT_BYTES = (1 << 20) + 4096          # 8 slots of ~1 MB: the 20 MB upload laps the ring twice
        engine._stager.ring = None
        x = rng.standard_normal((262_147, 10))
        for arr in (x, np.ascontiguousarray(x.T).T):
            t, code, s_t, s_m = engine._as_device_block(arr, 3, x.shape[0] - 5, torch.device("cuda", 0))
            torch.cuda.synchronize()
            got = t.cpu().numpy()
            want = arr[3:x.shape[0] - 5]
            assert np.array_equal(got if s_m == 1 else got.T, want)
        xs = x[:40_000]
        ref = R.csm_welch(xs, 8000, 256, "hann", 50, True, "fftbackward")[1]
        c = dsp.Signal(None, xs, 8000).set_spectrum_parameters(window_length_samples=256).get_csm()[1]
        assert_csm_close(c, ref)
        # ... and the way back (engine.to_host: device -> pinned slots -> host threads -> numpy), real and complex
        for td in (torch.randn((1_300_001, 3), device="cuda", dtype=torch.float64),
                   torch.randn((700_003, 2), device="cuda", dtype=torch.complex128),
                   torch.randn((2_000_001,), device="cuda", dtype=torch.float32)):
            got = engine.to_host(td)
            assert got.dtype == td.cpu().numpy().dtype and np.array_equal(got, td.cpu().numpy())
    finally:
        engine.config.h2d_stage_min_bytes, engine._Stager.SLOT_BYTES = old
        engine._stager.ring = None


def test_frame_chunking_and_sharding_are_exact_partitions():
    """Sums over disjoint frame ranges add up to the full sum (what frame sharding relies on)."""
    rng = np.random.default_rng(6)
    x = rng.standard_normal((20000, 9))
    w = R.welch_window("hann", 256)
    full, K = engine.csm_partial_sums(x, 20000, 256, 128, w, True, 1.0)
    parts = [engine.csm_partial_sums(x, 20000, 256, 128, w, True, 1.0, frame_range=r)[0]
             for r in ((0, 50), (50, 51), (51, K))]
    assert full.shape == (129, engine.packed_size(9))       # packed lower triangle: (j, i), j >= i, at j(j+1)/2 + i
    tot = (parts[0] + parts[1] + parts[2]).cpu().numpy()
    fu = full.cpu().numpy()
    assert np.max(np.abs(tot - fu)) <= 1e-12 * np.abs(fu).max()
    # the packed layout against the oracle's un-normalised sums
    Xo = R.channel_spectra(x, 256, "hann", 50, True, "fftbackward")                    # (nb, K, M)
    So = np.matmul(np.swapaxes(Xo, 1, 2), Xo.conj())
    un = engine.unpack_lower(full, 9).cpu().numpy()
    low = np.tril(np.ones((9, 9), bool))
    assert np.max(np.abs(un[1:][:, low] - So[1:][:, low])) <= 1e-12 * np.abs(So).max()
    assert np.all(un[:, ~low] == 0)
    old = engine.config.spectra_budget_bytes
    try:
        engine.config.spectra_budget_bytes = 129 * 9 * 16 * 40          # forces several chunks of 32 frames
        ch, _ = engine.csm_partial_sums(x, 20000, 256, 128, w, True, 1.0)
    finally:
        engine.config.spectra_budget_bytes = old
    assert np.max(np.abs(ch.cpu().numpy() - fu)) <= 1e-12 * np.abs(fu).max()
    # host recordings are uploaded in frame chunks on a copy stream (overlapping the kernels): same sums
    oldc = engine.config.h2d_chunk_bytes
    try:
        engine.config.h2d_chunk_bytes = 20000 * 9 * 8 // 5
        hc, _ = engine.csm_partial_sums(x, 20000, 256, 128, w, True, 1.0)
        assert engine.config.stats["csm"]["frame_chunks"] >= 4
    finally:
        engine.config.h2d_chunk_bytes = oldc
    assert np.max(np.abs(hc.cpu().numpy() - fu)) <= 1e-12 * np.abs(fu).max()


def test_fast_precision_mode():
    """fp32 FFT + fp64 accumulation: inside rtol 1e-5 on source-dominated data, and always tiny
    relative to sqrt(C_ii C_jj)."""
    mic = synth.positions_to_xyz(synth.disc_array(16, 0.5, seed=1))
    x = synth.point_source_recording(mic, 30000, 48000, noise_sigma=0.1, seed=2)
    ref = R.csm_welch(x, 48000, 1024, "hann", 50, True, "powerspectraldensity")[1]
    engine.config.csm_precision = "fp32"
    s = dsp.Signal(None, x, 48000).set_spectrum_parameters(window_length_samples=1024, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
    c = s.get_csm()[1]
    e = csm_errors(c, ref)
    assert e["max_coherence_normalised"] < 2e-6, e
    assert e["pass_fraction"] > 0.999, e


def test_cache_semantics():
    rng = np.random.default_rng(5)
    s = dsp.Signal(None, rng.standard_normal((4000, 3)), 8000, activate_cache=True)
    f1, c1 = s.get_csm()
    assert not s._csm_dirty
    f2, c2 = s.get_csm()
    assert c1 is not c2 and np.array_equal(c1, c2)                 # fresh copies, same values
    s.set_spectrum_parameters(window_length_samples=512)
    assert s._csm_dirty
    assert s.get_csm()[1].shape[0] == 257
    s.time_data = rng.standard_normal((4000, 3))
    assert s._csm_dirty


@pytest.mark.parametrize("scaling", SCALINGS)
def test_get_spectrum_welch_golden(golden_dir, scaling):
    """Signal.get_spectrum, WelchPeriodogram branch (reference classes/signal.py:886-898)."""
    g = np.load(os.path.join(golden_dir, "csm_small_scalings.npz"))
    s = dsp.Signal(None, g["x"].copy(), int(g["fs"]))
    s.set_spectrum_parameters(window_length_samples=int(g["wl"]), scaling=getattr(dsp.SpectrumScaling, scaling))
    f, sp = s.get_spectrum()
    ref = g["spectrum_" + scaling]
    assert sp.shape == ref.shape and sp.dtype == np.float64 and np.array_equal(f, g["f"])
    assert np.max(np.abs(sp[1:] - ref[1:]) / np.abs(ref[1:])) <= 1e-9
    # the auto-spectra are the diagonal of the CSM
    c = s.get_csm()[1]
    assert np.allclose(sp[1:], np.einsum("fii->fi", c).real[1:], rtol=1e-10)
    s.spectrum_method = dsp.SpectrumMethod.FFT               # the single-FFT branch (tests/test_gpu_fft_method.py)
    assert s.get_spectrum()[1].shape[0] != sp.shape[0]
