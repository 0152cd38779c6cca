"""Generate tests/golden/*.npz by running the REAL reference (stub-loaded, unmodified).

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden

Every fixture stores the inputs (or the seed + a fingerprint of the regenerated input when
the input is too large to commit) and the reference's own outputs of
`Signal.get_csm()` (classes/signal.py:948) and
`BeamformerDASFrequency.get_beamformer_map()` (beamforming/beamforming.py:805).
While generating, the oracle restatement is checked against the reference on each case and
the script aborts if they disagree (that is the oracle's pin).
"""

from __future__ import annotations

import contextlib
import io
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import ref_loader, restatement as R  # noqa: E402
from dsptoolbox_b200 import synth  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

SCALINGS = ["AmplitudeSpectrum", "AmplitudeSpectralDensity", "PowerSpectrum",
            "PowerSpectralDensity", "FFTBackward", "FFTForward", "FFTOrthogonal"]


def rel_to_max(a, b):
    return float(np.max(np.abs(a - b)) / np.max(np.abs(b)))


def ref_csm(dsp, x, fs, wl, window_enum, overlap, detrend, scaling_name):
    s = dsp.Signal(None, x.copy(), fs)
    s.set_spectrum_parameters(
        method=dsp.SpectrumMethod.WelchPeriodogram, window_length_samples=wl,
        window_type=window_enum, overlap_percent=overlap, detrend=detrend, average="mean",
        scaling=getattr(dsp.SpectrumScaling, scaling_name))
    return s.get_csm()


def ref_map(dsp, x, fs, wl, mic_pos, grid, formulation, fc, frac, remove_diag, c=343):
    s = dsp.Signal(None, x.copy(), fs)
    s.set_spectrum_parameters(window_length_samples=wl)
    bf = dsp.beamforming.BeamformerDASFrequency(
        s, dsp.beamforming.MicArray(mic_pos), grid,
        dsp.beamforming.SteeringVector(getattr(dsp.beamforming.SteeringVectorType, formulation)), c=c)
    with contextlib.redirect_stdout(io.StringIO()):
        m = bf.get_beamformer_map(fc, frac, remove_diag)
    return m, np.asarray(bf.f_range_hz)


STFT_CASES = [  # (wl, window, overlap %, fft_length, detrend, padding, scaling)
    (256, "Hann", 50.0, None, False, True, "FFTBackward"),
    (256, "Hann", 75.0, None, True, True, "PowerSpectralDensity"),
    (128, "Hamming", 50.0, None, False, False, "AmplitudeSpectrum"),
    (512, "Blackman", 66.0, None, True, False, "PowerSpectrum"),
    (64, "Hann", 0.0, None, False, True, "FFTOrthogonal"),
    (256, "Hann", 50.0, None, True, True, "AmplitudeSpectralDensity"),
    (128, "Hann", 50.0, None, False, True, "FFTForward"),
    (128, "Hann", 50.0, 256, False, True, "PowerSpectralDensity"),     # zero-padded FFT (oracle only)
    (256, "Hann", 50.0, 128, True, False, "FFTBackward"),               # cropped frames (oracle only)
]


def stft_golden(dsp):
    """Next row f4: `_stft` (standard/_spectral_methods.py:176-282) through Signal.get_spectrogram."""
    import warnings
    rng = np.random.default_rng(77)
    fs = 8000
    x = rng.standard_normal((3000, 3)) * np.array([1.0, 0.3, 2.0]) + np.array([0.5, 0.0, -1.0])
    out = dict(x=x, fs=fs, n_cases=len(STFT_CASES))
    for i, (wl, win, ov, nfft, det, pad, sc) in enumerate(STFT_CASES):
        s = dsp.Signal(None, x.copy(), fs)
        s.set_spectrogram_parameters(window_length_samples=wl, window_type=getattr(dsp.Window, win), overlap_percent=ov,
                                     fft_length_samples=nfft, detrend=det, padding=pad,
                                     scaling=getattr(dsp.SpectrumScaling, sc))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            t, f, st = s.get_spectrogram()
        to, fo, so = R.stft(x, fs, wl, win.lower(), ov, nfft, det, pad, sc)
        err = rel_to_max(so, st)
        print(f"F{i} stft wl={wl} {win} ov={ov} n={nfft} detrend={det} pad={pad} {sc}: shape {st.shape} {st.dtype} "
              f"oracle-vs-ref {err:.2e}")
        assert so.shape == st.shape and so.dtype == st.dtype and err < 1e-13
        assert np.allclose(to, t, rtol=0, atol=1e-15) and np.allclose(fo, f, rtol=0, atol=1e-12)
        out[f"params_{i}"] = np.array([wl, ov, -1 if nfft is None else nfft, int(det), int(pad)], dtype=np.float64)
        out[f"names_{i}"] = np.array([win, sc])
        out[f"t_{i}"], out[f"f_{i}"], out[f"stft_{i}"] = t, f, st
    np.savez_compressed(os.path.join(GOLDEN, "stft_small.npz"), **out)


TF_CASES = [  # (mode, shared one-channel input, window length, overlap %, detrend, scaling)
    ("H1", False, 256, 50, True, "FFTBackward"),
    ("H2", False, 256, 50, True, "FFTBackward"),
    ("H3", False, 256, 50, True, "FFTBackward"),
    ("H1", True, 128, 75, False, "PowerSpectralDensity"),
    ("H2", True, 512, 50, True, "AmplitudeSpectrum"),
    ("H3", True, 256, 25, True, "PowerSpectrum"),
]


def tf_golden(dsp):
    """Next row f3: `compute_transfer_function` (transfer_functions/transfer_functions.py:419-536)."""
    import warnings
    from scipy.signal import lfilter
    rng = np.random.default_rng(2024)
    fs = 16000
    L, ch = 6000, 3
    xin = rng.standard_normal((L, ch))
    # outputs: each input through a short FIR plus uncorrelated noise (coherence < 1)
    firs = [np.array([1.0, 0.5, -0.25]), np.array([0.2, 0.0, 0.9, 0.1]), np.array([-0.7, 0.3])]
    yout = np.stack([lfilter(firs[c], [1.0], xin[:, c]) for c in range(ch)], 1) + 0.2 * rng.standard_normal((L, ch))
    yshared = np.stack([lfilter(firs[c], [1.0], xin[:, 0]) for c in range(ch)], 1) + 0.2 * rng.standard_normal((L, ch))
    out = dict(x_in=xin, y_out=yout, y_shared=yshared, fs=fs, n_cases=len(TF_CASES))
    for i, (mode, shared, wl, ov, det, sc) in enumerate(TF_CASES):
        sin = dsp.Signal(None, (xin[:, :1] if shared else xin).copy(), fs)
        sin.set_spectrum_parameters(overlap_percent=ov, detrend=det, scaling=getattr(dsp.SpectrumScaling, sc))
        sout = dsp.Signal(None, (yshared if shared else yout).copy(), fs)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            spec = dsp.transfer_functions.compute_transfer_function(
                sout, sin, wl, getattr(dsp.transfer_functions.TransferFunctionType, mode))
        fo, tfo, coho = R.transfer_function(yshared if shared else yout, xin[:, :1] if shared else xin, fs, wl, "hann",
                                            ov, det, sc, mode)
        e1, e2 = rel_to_max(tfo, spec.spectral_data), rel_to_max(coho, spec.coherence)
        print(f"G{i} {mode} shared={shared} wl={wl} ov={ov} detrend={det} {sc}: tf oracle-vs-ref {e1:.2e} coherence {e2:.2e}")
        assert e1 < 1e-12 and e2 < 1e-12 and np.array_equal(fo, spec.frequency_vector_hz)
        out[f"params_{i}"] = np.array([int(shared), wl, ov, int(det)], dtype=np.float64)
        out[f"names_{i}"] = np.array([mode, sc])
        out[f"tf_{i}"], out[f"coh_{i}"], out[f"f_{i}"] = spec.spectral_data, spec.coherence, spec.frequency_vector_hz
    np.savez_compressed(os.path.join(GOLDEN, "transfer_functions.npz"), **out)


FFT_CASES = [  # (samples, channels, pad_to_fast_length, scaling)
    (2001, 3, True, "FFTBackward"),
    (2001, 3, False, "FFTBackward"),              # odd, prime-ish length: Bluestein
    (1536, 2, True, "FFTForward"),
    (4096, 2, True, "FFTOrthogonal"),             # power of two
    (1999, 3, False, "AmplitudeSpectrum"),        # prime length
    (3000, 2, True, "AmplitudeSpectralDensity"),
    (2500, 3, True, "PowerSpectrum"),
    (2187, 2, False, "PowerSpectralDensity"),
]


def fft_golden(dsp):
    """Next row f3: SpectrumMethod.FFT -- `Signal.get_spectrum` (classes/signal.py:903-946), `Signal.get_csm` ->
    `_csm_fft` (:989-1002, _spectral_methods.py:374-443)."""
    rng = np.random.default_rng(31)
    fs = 12000
    out = dict(fs=fs, n_cases=len(FFT_CASES))
    for i, (L, ch, fast, sc) in enumerate(FFT_CASES):
        x = rng.standard_normal((L, ch)) * np.linspace(1.0, 2.0, ch) + 0.3
        s = dsp.Signal(None, x.copy(), fs)
        s.set_spectrum_parameters(method=dsp.SpectrumMethod.FFT, pad_to_fast_length=fast, scaling=getattr(dsp.SpectrumScaling, sc))
        f, sp = s.get_spectrum()
        fc, csm = s.get_csm()
        fo, spo = R.spectrum_fft(x, fs, sc, fast)
        _, sp_b = R.spectrum_fft(x, fs, "fftbackward", fast)
        csmo = R.csm_fft(sp_b, sc, fs)
        e1, e2 = rel_to_max(spo, sp), rel_to_max(csmo, csm)
        print(f"H{i} FFT method L={L} ch={ch} fast={fast} {sc}: n={2 * (len(f) - 1)}.. spectrum oracle-vs-ref {e1:.2e} ({sp.dtype}), csm {e2:.2e}")
        assert spo.shape == sp.shape and spo.dtype == sp.dtype and e1 < 1e-12 and np.allclose(fo, f, rtol=0, atol=1e-9)
        assert csmo.shape == csm.shape and e2 < 1e-12 and np.allclose(fc, f)
        out[f"x_{i}"], out[f"names_{i}"], out[f"fast_{i}"] = x, np.array([sc]), np.array([int(fast)])
        out[f"f_{i}"], out[f"sp_{i}"], out[f"csm_{i}"] = f, sp, csm
    np.savez_compressed(os.path.join(GOLDEN, "fft_method.npz"), **out)


def sources_golden(dsp):
    """Next row f2: `MonopoleSource.get_signals_on_array` / `mix_sources_on_array`
    (beamforming/beamforming.py:1397-1511) with `fractional_delay` (standard/latency_delay.py:159-257)."""
    import warnings

    rng = np.random.default_rng(41)
    fs = 16000
    mic = np.concatenate([rng.uniform(-0.4, 0.4, (9, 2)), np.zeros((9, 1))], axis=1)
    mic[3] = [0.02, 0.0, 0.0]                                  # 2 cm from source 2: negative integer shift
    mics = dsp.beamforming.MicArray(dict(x=mic[:, 0], y=mic[:, 1], z=mic[:, 2]))
    src = np.array([[0.3, -0.2, 1.5], [-1.0, 0.7, 2.5], [0.0, 0.0, 0.0], [0.1, 0.1, 0.8]])
    src[2] = mic[5]                                            # a source ON a microphone: zero delay for that channel
    src[3] = mic[3] + [0.02, 0.0, 0.0]
    lens = [3000, 3000, 2500, 2800]
    sig = [rng.standard_normal(n) * (1 + 0.5 * i) for i, n in enumerate(lens)]
    out = dict(fs=fs, mic=mic, src=src, c=343.0, n_src=len(src))
    for i, x in enumerate(sig):
        out[f"x_{i}"] = x
        got = dsp.beamforming.MonopoleSource(dsp.Signal(None, x.copy(), fs), src[i]).get_signals_on_array(mics, 343.0).time_data
        mine = R.monopole_on_array(x, fs, src[i], mic, 343.0)
        print(f"I{i} monopole source {i}: {got.shape}, oracle-vs-ref {rel_to_max(mine, got):.2e}")
        assert mine.shape == got.shape and rel_to_max(mine, got) < 1e-13
        out[f"on_array_{i}"] = got
    for tag, ids in [("equal", [0, 1]), ("ragged", [0, 1, 2, 3]), ("ragged2", [2, 0, 3])]:
        srcs = [dsp.beamforming.MonopoleSource(dsp.Signal(None, sig[i].copy(), fs), src[i]) for i in ids]
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            got = dsp.beamforming.mix_sources_on_array(srcs, mics, 343.0).time_data
        mine = R.mix_sources_on_array([sig[i] for i in ids], [src[i] for i in ids], fs, mic, 343.0)
        print(f"I mix {tag} {ids}: {got.shape}, oracle-vs-ref {rel_to_max(mine, got):.2e}")
        assert mine.shape == got.shape and rel_to_max(mine, got) < 1e-13
        out[f"mix_{tag}"], out[f"mix_{tag}_ids"] = got, np.array(ids)
    for d in (0.3, 7.5, 123.456, 15.0, 0.5):                  # the filter itself at a few delays
        di, h = dsp.standard._standard_backend._fractional_delay_filter(d, 30, 60)
        dm, hm = R.fractional_delay_filter(d)
        assert di == dm and np.array_equal(h, hm), d
    np.savez_compressed(os.path.join(GOLDEN, "sources_on_array.npz"), **out)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "sources":     # regenerate only the source-simulation fixture
        sources_golden(ref_loader.load())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "fft":         # regenerate only the FFT-method fixture
        fft_golden(ref_loader.load())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "stft":        # regenerate only the STFT fixture
        stft_golden(ref_loader.load())
        return
    if len(sys.argv) > 1 and sys.argv[1] == "tf":          # regenerate only the transfer-function fixture
        tf_golden(ref_loader.load())
        return
    os.makedirs(GOLDEN, exist_ok=True)
    dsp = ref_loader.load()
    W = dsp.Window

    # ---- geometry of example_data/array.xml (config 1) -------------------------------
    arr = ref_loader.read_array_xml()
    np.savez(os.path.join(GOLDEN, "array_xml_positions.npz"), **arr)

    # ---- case A: small CSM, all seven scalings ----------------------------------------
    mic = synth.positions_to_xyz(synth.disc_array(4, 0.3, seed=3))
    fs = 8000
    x = synth.point_source_recording(mic, 3000, fs, noise_sigma=0.3, seed=11)
    out = dict(x=x, fs=fs, wl=128, overlap=50, detrend=True, window="hann")
    for sc in SCALINGS:
        f, c = ref_csm(dsp, x, fs, 128, W.Hann, 50, True, sc)
        _, c1 = R.csm_welch(x, fs, 128, "hann", 50, True, sc)
        _, c2 = R.csm_welch_pairwise(x, fs, 128, "hann", 50, True, sc)
        e1, e2 = rel_to_max(c1, c), rel_to_max(c2, c)
        print(f"A {sc:26s} oracle-vs-ref rel-to-max: vectorised {e1:.2e} pairwise {e2:.2e}")
        assert e1 < 1e-13 and e2 < 1e-13
        out["csm_" + sc] = c
        # Signal.get_spectrum, Welch branch (classes/signal.py:886-898): auto-spectra of every channel
        sg = dsp.Signal(None, x.copy(), fs)
        sg.set_spectrum_parameters(window_length_samples=128, scaling=getattr(dsp.SpectrumScaling, sc))
        fsp, sp = sg.get_spectrum()
        so = np.stack([R.welch_pair(x[:, i], None, fs, "hann", 128, 50, True, sc) for i in range(x.shape[1])], axis=1)
        assert rel_to_max(so, sp) < 1e-13
        out["spectrum_" + sc] = sp
    out["f"] = f
    np.savez_compressed(os.path.join(GOLDEN, "csm_small_scalings.npz"), **out)

    # ---- case B: Welch parameter variants ----------------------------------------------
    variants = [
        dict(tag="ov75_hamming", L=2500, wl=64, overlap=75, detrend=True, window=("Hamming", "hamming")),
        dict(tag="ov0_boxcar_nodetrend", L=2048, wl=256, overlap=0, detrend=False, window=("Boxcar", "boxcar")),
        dict(tag="ov33_blackman_Lmultiple", L=43 * 86, wl=128, overlap=33, detrend=True, window=("Blackman", "blackman")),
        dict(tag="ov50_kaiser8_wl8", L=999, wl=8, overlap=50, detrend=True, window=("Kaiser", ("kaiser", 8.0))),
        dict(tag="ov90_flattop_short", L=300, wl=512, overlap=90, detrend=True, window=("Flattop", "flattop")),
    ]
    mic3 = synth.positions_to_xyz(synth.disc_array(3, 0.2, seed=5))
    outB = {}
    for v in variants:
        xv = synth.point_source_recording(mic3, v["L"], fs, noise_sigma=0.5, seed=21)
        wenum = getattr(W, v["window"][0])
        if isinstance(v["window"][1], tuple):
            wenum = wenum.with_extra_parameter(v["window"][1][1])
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            f, c = ref_csm(dsp, xv, fs, v["wl"], wenum, v["overlap"], v["detrend"], "PowerSpectralDensity")
            _, c1 = R.csm_welch(xv, fs, v["wl"], v["window"][1], v["overlap"], v["detrend"], "PowerSpectralDensity")
        e1 = rel_to_max(c1, c)
        print(f"B {v['tag']:28s} oracle-vs-ref {e1:.2e}")
        assert e1 < 1e-13
        outB["x_" + v["tag"]] = xv
        outB["csm_" + v["tag"]] = c
    np.savez_compressed(os.path.join(GOLDEN, "csm_variants.npz"), fs=fs, **outB)

    # ---- case C: small DAS maps: 4 steering formulations x diag removal x band widths ---
    rng = np.random.default_rng(17)
    micp = dict(x=rng.uniform(-0.4, 0.4, 10), y=rng.uniform(-0.4, 0.4, 10), z=rng.uniform(-0.05, 0.05, 10))
    micxyz = synth.positions_to_xyz(micp)
    fs = 16000
    xm = synth.point_source_recording(micxyz, 6000, fs, source_xyz=(0.15, -0.1, 0.8), noise_sigma=0.2, seed=31)
    l1, l2 = np.linspace(-0.4, 0.4, 9), np.linspace(-0.3, 0.3, 7)
    grid = dsp.beamforming.Regular2DGrid(l1, l2, ("x", "y"), 0.8)
    gxyz = grid.coordinates
    outC = dict(x=xm, fs=fs, wl=256, mic_x=micp["x"], mic_y=micp["y"], mic_z=micp["z"], line1=l1, line2=l2, value3=0.8)
    fref, cref = ref_csm(dsp, xm, fs, 256, W.Hann, 50, True, "FFTBackward")
    for form in ["Classic", "Inverse", "TruePower", "TrueLocation"]:
        for (fc, frac, rd, tag) in [(2000.0, 3, True, "third_rd"), (2000.0, 3, False, "third_full"),
                                    (3000.0, 0, True, "single_rd"), (1500.0, 1, True, "octave_rd"),
                                    (2500.0, 12, False, "twelfth_full")]:
            m, fr = ref_map(dsp, xm, fs, 256, micp, grid, form, fc, frac, rd)
            mo = R.das_map(fref, cref, micxyz, gxyz, fc, frac, rd, 343.0, form).reshape(m.shape)
            ml = R.das_map_loop(fref, cref, micxyz, gxyz, fc, frac, rd, 343.0, form).reshape(m.shape)
            e, el = rel_to_max(mo, m), rel_to_max(ml, m)
            print(f"C {form:13s} {tag:13s} F={R.band_bins(fref, fc, frac)} oracle-vs-ref {e:.2e} loop {el:.2e}")
            assert e < 1e-12 and el < 1e-12
            outC[f"map_{form}_{tag}"] = m
            outC[f"frange_{form}_{tag}"] = fr
    # 3-D and line grids, plain Grid
    g3 = dsp.beamforming.Regular3DGrid(np.linspace(-0.2, 0.2, 4), np.linspace(-0.2, 0.2, 3), np.linspace(0.6, 1.0, 5))
    m3, _ = ref_map(dsp, xm, fs, 256, micp, g3, "TrueLocation", 2000.0, 3, True)
    outC["map3d"] = m3
    gl = dsp.beamforming.LineGrid(np.linspace(-0.5, 0.5, 11), "y", 0.8, 0.1)
    ml_, _ = ref_map(dsp, xm, fs, 256, micp, gl, "TrueLocation", 2000.0, 3, True)
    outC["mapline"] = ml_
    outC["mapline_coords"] = gl.coordinates
    np.savez_compressed(os.path.join(GOLDEN, "das_small.npz"), **outC)

    # ---- case E: MVDR and Functional beamformers (next rows) on the case-C data ----
    outE = {}
    sE = dsp.Signal(None, xm.copy(), fs)
    sE.set_spectrum_parameters(window_length_samples=256)
    for form in ["Classic", "TrueLocation"]:
        stv = dsp.beamforming.SteeringVector(getattr(dsp.beamforming.SteeringVectorType, form))
        for (fc, frac, tag) in [(2000.0, 3, "third"), (3000.0, 0, "single")]:
            for cls, name, kw in [(dsp.beamforming.BeamformerMVDR, "mvdr", {}),
                                  (dsp.beamforming.BeamformerFunctional, "functional", dict(gamma=10)),
                                  (dsp.beamforming.BeamformerFunctional, "functional3", dict(gamma=3))]:
                bf = cls(sE, dsp.beamforming.MicArray(micp), grid, stv)
                with contextlib.redirect_stdout(io.StringIO()):
                    m = bf.get_beamformer_map(fc, frac, **kw)
                if name == "mvdr":
                    mo = R.mvdr_map(fref, cref, micxyz, gxyz, fc, frac, 343.0, form)
                else:
                    mo = R.functional_map(fref, cref, micxyz, gxyz, fc, frac, kw["gamma"], 343.0, form)
                e = rel_to_max(mo.reshape(m.shape), m)
                print(f"E {name:12s} {form:13s} {tag:7s} oracle-vs-ref {e:.2e}")
                assert e < 1e-9
                outE[f"{name}_{form}_{tag}"] = m
            for ne, nm in [(None, "orthogonal"), (3, "orthogonal3")]:
                bf = dsp.beamforming.BeamformerOrthogonal(sE, dsp.beamforming.MicArray(micp), grid, stv)
                with contextlib.redirect_stdout(io.StringIO()):
                    m = bf.get_beamformer_map(fc, frac, ne)
                mo = R.orthogonal_map(fref, cref, micxyz, gxyz, fc, frac, ne, 343.0, form)
                e = rel_to_max(mo.reshape(m.shape), m)
                print(f"E {nm:12s} {form:13s} {tag:7s} oracle-vs-ref {e:.2e}")
                assert e < 1e-9
                outE[f"{nm}_{form}_{tag}"] = m
            for kw, nm in [(dict(), "cleansc"), (dict(maximum_iterations=5, safety_factor=0.8, remove_csm_diagonal=True), "cleansc_rd")]:
                bf = dsp.beamforming.BeamformerCleanSC(sE, dsp.beamforming.MicArray(micp), grid, stv)
                with contextlib.redirect_stdout(io.StringIO()):
                    m = bf.get_beamformer_map(fc, frac, **kw)
                mo = R.clean_sc_map(fref, cref, micxyz, gxyz, fc, frac, kw.get("maximum_iterations"), kw.get("safety_factor", 0.5),
                                    kw.get("remove_csm_diagonal", False), 343.0, form)
                e = rel_to_max(mo.reshape(m.shape), m)
                print(f"E {nm:12s} {form:13s} {tag:7s} oracle-vs-ref {e:.2e}")
                assert e < 1e-9
                outE[f"{nm}_{form}_{tag}"] = m
    np.savez_compressed(os.path.join(GOLDEN, "beamformers_next.npz"), **outE)

    # ---- case D: config 1 (array.xml, 1 s @ 48 kHz, nfft 1024, 50x50 grid, 4 kHz third) ---
    micxyz = synth.positions_to_xyz(arr)
    fs = 48000
    x1 = synth.point_source_recording(micxyz, fs, fs, noise_sigma=0.1, seed=1)
    print("D running the reference at config 1 (about 15 s) ...")
    f1, c1 = ref_csm(dsp, x1, fs, 1024, W.Hann, 50, True, "FFTBackward")
    _, co = R.csm_welch(x1, fs, 1024, "hann", 50, True, "FFTBackward")
    print(f"D csm oracle-vs-ref {rel_to_max(co, c1):.2e}")
    assert rel_to_max(co, c1) < 1e-13
    l = np.linspace(-0.5, 0.5, 50)
    grid1 = dsp.beamforming.Regular2DGrid(l, l, ("x", "y"), 1.0)
    m1, fr1 = ref_map(dsp, x1, fs, 1024, arr, grid1, "TrueLocation", 4000.0, 3, True)
    mo = R.das_map(f1, c1, micxyz, grid1.coordinates, 4000.0, 3, True).reshape(m1.shape)
    print(f"D map oracle-vs-ref {rel_to_max(mo, m1):.2e}; argmax {np.unravel_index(np.argmax(m1), m1.shape)}")
    assert rel_to_max(mo, m1) < 1e-12
    rs = np.random.default_rng(5)
    nb, M = c1.shape[0], c1.shape[1]
    idx = np.stack([rs.integers(1, nb, 6000), rs.integers(0, M, 6000), rs.integers(0, M, 6000)], axis=1)
    np.savez_compressed(
        os.path.join(GOLDEN, "cfg1.npz"), fs=fs, seed=1, noise_sigma=0.1, wl=1024,
        x_fingerprint=np.array([x1.sum(), np.abs(x1).sum(), x1[12345, 7], x1[-1, -1]]),
        csm_idx=idx, csm_vals=c1[idx[:, 0], idx[:, 1], idx[:, 2]],
        csm_bins=np.array([1, 85, 512]), csm_full_bins=c1[[1, 85, 512]],
        csm_diag_bin85=np.diagonal(c1[85]).copy(), map=m1, f_range=fr1)
    stft_golden(dsp)
    tf_golden(dsp)
    fft_golden(dsp)
    sources_golden(dsp)
    print("golden vectors written to", GOLDEN)


if __name__ == "__main__":
    main()
