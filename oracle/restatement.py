"""CPU restatement (numpy/scipy, float64) of the reference's Welch-CSM + DAS-map path.

TEST INFRASTRUCTURE ONLY.  Nothing in ``dsptoolbox_b200`` may import this module; it is
used by ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl
reference`` legs of ``bench.py`` as the checker / CPU baseline, never as the product.

Parity pin: every function here is checked against the *real* reference (imported from
``/root/reference`` with its I/O/plot dependencies stubbed, see ``oracle/ref_loader.py``)
by ``oracle/make_golden.py`` (which also writes ``tests/golden/*.npz``) and, whenever the
reference tree is present, by ``tests/test_oracle_vs_reference.py``.  The committed golden
vectors are outputs of the reference itself, so the oracle is pinned on the GPU box too.

Every function cites the reference lines it restates (paths relative to
``/root/reference/dsptoolbox``).

Two flavours are provided for each stage:

* ``*_pairwise`` / ``*_loop``  -- keep the reference's own loop structure (one Welch call
  per channel pair, one ``multi_dot`` per grid point and bin).  These are the "reference
  arm" timed by ``bench.py --impl reference``.
* the vectorised forms -- the same arithmetic as batched BLAS calls; the strong CPU
  baseline and the checker used at larger sizes.
"""

from __future__ import annotations

import numpy as np
from scipy.integrate import simpson as _scipy_simpson
from scipy.signal.windows import get_window as _scipy_get_window

# --------------------------------------------------------------------------- scalings
# standard/enums.py:21-229 (SpectrumScaling).  Names are the lower-cased enum members.
AMPLITUDE_SCALINGS = (
    "amplitudespectrum",
    "amplitudespectraldensity",
    "fftbackward",
    "fftforward",
    "fftorthogonal",
)  # enums.py:77-92
PHYSICAL_SCALINGS = (
    "amplitudespectrum",
    "amplitudespectraldensity",
    "powerspectrum",
    "powerspectraldensity",
)  # enums.py:109-123


def fft_norm(scaling: str) -> str:
    """enums.py:53-75."""
    scaling = scaling.lower()
    if scaling == "fftforward":
        return "forward"
    if scaling == "fftorthogonal":
        return "ortho"
    return "backward"


def scaling_factor(scaling: str, wl: int, fs: int, window: np.ndarray) -> float:
    """enums.py:183-229 with a window always passed (as `_welch` does, :165)."""
    scaling = scaling.lower()
    if scaling == "fftbackward":
        return 1.0
    if scaling == "fftforward":
        return 1.0 / wl
    if scaling == "fftorthogonal":
        return (1.0 / wl) ** 0.5
    if scaling in ("amplitudespectraldensity", "powerspectraldensity"):
        factor = (2.0 / np.sum(window**2) / fs) ** 0.5
    else:
        factor = 2.0**0.5 / np.sum(window)
    if scaling in AMPLITUDE_SCALINGS:
        return float(factor)
    return float(factor**2.0)


# --------------------------------------------------------------------------- framing
def frame_count(wl: int, step: int, length: int) -> tuple[int, int]:
    """helpers/other.py:181-213 with zero_padding=True: (frames, padded samples)."""
    n_frames = int(np.ceil(length / step))
    pad = wl - int(length % step)
    return n_frames, pad


def framed(x: np.ndarray, wl: int, step: int) -> np.ndarray:
    """standard/_framed_signal_representation.py:9-67 for one or more channels.

    x: (L, ch) -> (wl, K, ch) float64, tail zero padded (helpers/other.py:216-259).
    """
    x = np.asarray(x, dtype=np.float64)
    L = x.shape[0]
    K, pad = frame_count(wl, step, L)
    xp = np.concatenate([x, np.zeros((pad, x.shape[1]))], axis=0)
    idx = np.arange(wl)[:, None] + step * np.arange(K)[None, :]
    return xp[idx, :]


def welch_window(window, wl: int) -> np.ndarray:
    """_spectral_methods.py:103-105 (periodic window, fftbins=True)."""
    return _scipy_get_window(window, wl, fftbins=True)


def welch_step(wl: int, overlap_percent: float) -> int:
    """_spectral_methods.py:106-107."""
    return wl - int(overlap_percent / 100 * wl)


# --------------------------------------------------------------------------- stage 1
def channel_spectra(x, wl, window, overlap_percent, detrend, scaling) -> np.ndarray:
    """Windowed, de-meaned frame spectra of every channel.

    _spectral_methods.py:120-148: frames * window, THEN minus the frame mean, then rfft
    with the scaling's norm.  Returns complex128 (nb, K, ch).
    """
    w = welch_window(window, wl)
    step = welch_step(wl, overlap_percent)
    fr = framed(x, wl, step)
    fr *= w[:, None, None]
    if detrend:
        fr -= np.mean(fr, axis=0)
    return np.fft.rfft(fr, axis=0, norm=fft_norm(scaling))


def _finish_welch(csd, scaling, wl, fs, window):
    """_spectral_methods.py:164-171 on an averaged cross-spectrum (bins on axis 0)."""
    scaling = scaling.lower()
    if scaling in PHYSICAL_SCALINGS:
        csd = csd * scaling_factor(scaling, wl, fs, window)
        csd[np.array([0, -1]), ...] /= 2
    if scaling in AMPLITUDE_SCALINGS:
        csd = np.sqrt(csd)
    return csd


def welch_pair(x, y, fs, window, wl, overlap_percent, detrend, scaling):
    """One `_welch(x, y)` call (_spectral_methods.py:10-173, average='mean').

    x, y are 1-D; y=None gives the real auto-spectrum.
    """
    w = welch_window(window, wl)
    X = channel_spectra(np.asarray(x)[:, None], wl, window, overlap_percent, detrend, scaling)[..., 0]
    if y is None:
        sp = np.abs(X) ** 2.0
    else:
        Y = channel_spectra(np.asarray(y)[:, None], wl, window, overlap_percent, detrend, scaling)[..., 0]
        sp = X.conjugate() * Y
    csd = np.mean(sp, axis=1)
    return _finish_welch(csd, scaling, wl, fs, w)


def csm_welch_pairwise(x, fs, wl=1024, window="hann", overlap_percent=50, detrend=True,
                       scaling="fftbackward", pairs=None):
    """`_csm_welch` with the reference's own loop structure (_spectral_methods.py:285-371).

    ``pairs`` (list of (ind1, ind2), ind2 >= ind1) restricts the double loop to a sub-sample
    (used to time the reference arm on a bounded sample); entries outside stay zero.
    """
    x = np.asarray(x)
    M = x.shape[1]
    nb = wl // 2 + 1
    csm = np.zeros((nb, M, M), dtype=np.complex128)
    todo = pairs if pairs is not None else [(i, j) for i in range(M) for j in range(i, M)]
    for ind1, ind2 in todo:
        csm[:, ind2, ind1] = welch_pair(
            x[:, ind1], x[:, ind2] if ind1 != ind2 else None, fs, window, wl,
            overlap_percent, detrend, scaling)
        if ind1 == ind2:
            csm[:, ind1, ind2] *= 0.5
    csm += np.swapaxes(csm, 1, 2).conjugate()
    return np.fft.rfftfreq(wl, 1 / fs), csm


def csm_welch(x, fs, wl=1024, window="hann", overlap_percent=50, detrend=True,
              scaling="fftbackward", chunk_bins: int = 128):
    """Vectorised `_csm_welch`: every channel transformed once, one zgemm per bin block.

    Same arithmetic as `csm_welch_pairwise`:
    S[f, j, i] = mean_k X_k[f, j] conj(X_k[f, i]) for j >= i, scaling/sqrt applied to the
    lower-triangle value, upper triangle = conjugate of the lower one
    (_spectral_methods.py:351-369).
    """
    x = np.asarray(x)
    M = x.shape[1]
    nb = wl // 2 + 1
    w = welch_window(window, wl)
    X = channel_spectra(x, wl, window, overlap_percent, detrend, scaling)  # (nb, K, M)
    K = X.shape[1]
    csm = np.empty((nb, M, M), dtype=np.complex128)
    for b0 in range(0, nb, chunk_bins):
        Xb = X[b0:b0 + chunk_bins]                      # (b, K, M)
        # S[f, j, i] = sum_k X[f,k,j] conj(X[f,k,i]) / K
        csm[b0:b0 + chunk_bins] = np.matmul(np.swapaxes(Xb, 1, 2), Xb.conjugate()) / K
    lower = np.tril(np.ones((M, M), dtype=bool))
    S = _finish_welch(csm, scaling, wl, fs, w)
    # mirror the (possibly sqrt-ed) lower triangle; the diagonal stays as computed
    out = np.where(lower[None], S, 0)
    strict = np.tril(np.ones((M, M), dtype=bool), -1)
    out = out + np.swapaxes(np.where(strict[None], S, 0), 1, 2).conjugate()
    return np.fft.rfftfreq(wl, 1 / fs), out


# --------------------------------------------------------------------------- next row f3: H1/H2/H3
def transfer_function(x_out, x_in, fs, wl, window="hann", overlap_percent=50, detrend=True,
                      scaling="fftbackward", mode="h2"):
    """`compute_transfer_function` (transfer_functions/transfer_functions.py:419-536) with the reference's own
    loop: per output channel the auto-spectra G_xx, G_yy and the cross-spectra G_xy = _welch(x, y) (and
    G_yx = _welch(y, x) for H2), then H1 = G_xy/G_xx, H2 = G_yy/G_yx, H3 = G_xy/|G_xy| sqrt(G_yy/G_xx) and the
    coherence |G_xy|^2/G_xx/G_yy.  A one-channel input serves all outputs.  -> (f, tf (nb, ch), coherence)."""
    x_out = np.asarray(x_out, dtype=np.float64)
    x_in = np.asarray(x_in, dtype=np.float64)
    if x_in.ndim == 1:
        x_in = x_in[:, None]
    if x_out.ndim == 1:
        x_out = x_out[:, None]
    n_out = x_out.shape[1]
    shared_input = x_in.shape[1] == 1
    nb = wl // 2 + 1
    tf = np.zeros((nb, n_out), dtype=np.complex128)
    coherence = np.zeros((nb, n_out))
    mode = mode.lower()
    args = (fs, window, wl, overlap_percent, detrend, scaling)
    for n in range(n_out):
        xi = x_in[:, 0 if shared_input else n]
        y = x_out[:, n]
        g_xx = welch_pair(xi, None, *args)
        g_yy = welch_pair(y, None, *args)
        g_xy = welch_pair(xi, y, *args)
        if mode == "h1":
            tf[:, n] = g_xy / g_xx
        elif mode == "h2":
            tf[:, n] = g_yy / welch_pair(y, xi, *args)
        elif mode == "h3":
            tf[:, n] = g_xy / np.abs(g_xy) * (g_yy / g_xx) ** 0.5
        else:
            raise ValueError("Unsupported transfer function type")
        coherence[:, n] = np.abs(g_xy) ** 2 / g_xx / g_yy
    return np.fft.rfftfreq(wl, 1 / fs), tf, coherence


# --------------------------------------------------------------------------- next row f2: source simulation in front of stage 1
def fractional_delay_filter(delay_samples, filter_order=30, side_lobe_suppression_db=60.0):
    """`_fractional_delay_filter` (standard/_standard_backend.py:430-493) with its Kaiser window
    (`_kaiser_window_fractional` :290-321, beta :257-287): windowed-sinc FIR of `filter_order + 1` taps for the
    fractional part and the integer shift that completes the delay (it may be negative for short delays)."""
    from scipy.special import iv

    delay_int = int(delay_samples)
    frac = delay_samples - delay_int
    if filter_order % 2:
        m_opt = int(frac) - (filter_order - 1) / 2
    else:
        m_opt = np.round(frac) - filter_order / 2
    sinc = np.sinc(np.arange(filter_order + 1) + m_opt - frac)
    a = abs(side_lobe_suppression_db)
    beta = 0.1102 * (a - 8.7) if a > 50 else (0.5842 * (a - 21) ** 0.4 + 0.07886 * (a - 21) if a >= 21 else 0.0)
    alpha = filter_order / 2
    pos = np.arange(filter_order + 1).astype(float) - frac
    if filter_order % 2:
        pos += 0.5
    elif frac > 0.5:
        pos += 1
    z = beta * np.sqrt(np.array(1 - ((pos - alpha) / alpha) ** 2, dtype="complex"))
    kaiser = np.real(iv(0, z)) / iv(0, beta)
    return int(delay_int + m_opt), sinc * kaiser


def fractional_delay_keep_length(x, fs, delay_seconds, order=30, side_lobe_suppression_db=60.0):
    """`fractional_delay(sig, delay, keep_length=True)` on one channel (standard/latency_delay.py:159-257):
    full convolution with the filter, shifted by the integer delay (leading zeros, or leading samples dropped
    when it is negative), cut to the input length.  Direct convolution here, overlap-add FFT there."""
    x = np.asarray(x, dtype=np.float64)
    assert delay_seconds >= 0, "Delay must be positive"
    if delay_seconds == 0:
        return x.copy()
    delay_samples = delay_seconds * fs
    assert delay_samples < x.shape[0], "Delay too large for the given signal"
    d, h = fractional_delay_filter(delay_samples, order, side_lobe_suppression_db)
    full = np.convolve(x, h)
    y = np.concatenate([np.zeros(d), full]) if d >= 0 else full[-d:]
    return y[: x.shape[0]]


def monopole_on_array(x, fs, source_xyz, mic_xyz, c=343.0):
    """`MonopoleSource.get_signals_on_array` (beamforming/beamforming.py:1427-1458): channel m is the emitted
    signal delayed by r_m / c and divided by 1 + r_m, r_m from the expansion-formula distance matrix."""
    r = distance_matrix(np.asarray(mic_xyz, dtype=np.float64), np.asarray(source_xyz, dtype=np.float64)[None, :])[:, 0]
    return np.stack([fractional_delay_keep_length(x, fs, r[m] / c) / (1.0 + r[m]) for m in range(len(r))], axis=1)


def mix_sources_on_array(signals, sources_xyz, fs, mic_xyz, c=343.0):
    """`mix_sources_on_array` (beamforming/beamforming.py:1461-1511): sum of the sources' array signals; when the
    emitted signals differ in length the running sum and the next emitted signal are trimmed to the shorter."""
    out = monopole_on_array(signals[0], fs, sources_xyz[0], mic_xyz, c)
    total = out.shape[0]
    for x, xyz in zip(signals[1:], sources_xyz[1:]):
        x = np.asarray(x, dtype=np.float64)
        if total != x.shape[0]:
            total = min(total, x.shape[0])
            out, x = out[:total], x[:total]
        out = out + monopole_on_array(x, fs, xyz, mic_xyz, c)
    return out


# --------------------------------------------------------------------------- next row f3: single-FFT spectrum and CSM
def spectrum_fft(x, fs, scaling="fftbackward", pad_to_fast_length=True):
    """FFT branch of `Signal.get_spectrum` (classes/signal.py:903-946, smoothing = 0, no time window): one rfft over
    the whole recording of length next_fast_len(L, real) or L, with the scaling's norm; physical scalings go through
    `_scale_spectrum` (helpers/spectrum_utilities.py:268-330).  -> (freqs, spectrum (nb, channels))."""
    from scipy.fft import next_fast_len

    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    L = x.shape[0]
    n = next_fast_len(L, True) if pad_to_fast_length else L
    sp = np.fft.rfft(x, axis=0, n=n, norm=fft_norm(scaling))
    sc = scaling.lower()
    if sc in PHYSICAL_SCALINGS:
        if sc in ("amplitudespectraldensity", "powerspectraldensity"):
            factor = (2 / n / fs) ** 0.5
        else:
            factor = 2 ** 0.5 / n
        if sc not in AMPLITUDE_SCALINGS:
            factor = factor ** 2.0
        sp[0] /= 2 ** 0.5
        if n % 2 == 0:
            sp[-1] /= 2 ** 0.5
        if sc not in AMPLITUDE_SCALINGS:
            sp = np.abs(sp) ** 2
        sp = sp * factor
    return np.fft.rfftfreq(n, 1 / fs), sp


def csm_fft(spectrum, scaling, fs):
    """`_csm_fft` (_spectral_methods.py:374-443) without a time window: csm[f, j, i] = X[f, j] conj(X[f, i]); for
    scalings other than FFTBackward the first and last bin are halved, the matrix is multiplied by the output
    scaling's factor evaluated -- as the reference does -- for a time length of `nb // 2 + 1`, and the amplitude
    scalings take the element-wise square root."""
    sp = np.asarray(spectrum, dtype=np.complex128)
    full = sp[:, :, None] * sp[:, None, :].conj()
    idx = np.arange(sp.shape[1])
    # built like the reference (:418-426): lower triangle with half the diagonal, plus its conjugate transpose --
    # this fixes the sign of zero imaginary parts (real DC / Nyquist bins), i.e. the branch of the square root below
    low = np.tril(full, -1)
    low[:, idx, idx] = 0.5 * full[:, idx, idx]
    csm = low + np.swapaxes(low, 1, 2).conj()
    sc = scaling.lower()
    if sc == "fftbackward":
        return csm
    csm[np.array([0, -1])] /= 2.0
    n = sp.shape[0] // 2 + 1
    if sc == "fftforward":
        lin = 1.0 / n
    elif sc == "fftorthogonal":
        lin = (1.0 / n) ** 0.5
    elif sc in ("amplitudespectraldensity", "powerspectraldensity"):
        lin = (2 / n / fs) ** 0.5
    else:
        lin = 2 ** 0.5 / n
    # conversion from FFTBackward (amplitude, factor 1): linear factor for amplitude outputs, squared for power outputs
    factor = lin if sc in AMPLITUDE_SCALINGS else lin ** 2.0
    csm = csm * factor
    if sc in AMPLITUDE_SCALINGS:
        csm = np.sqrt(csm)
    return csm


# --------------------------------------------------------------------------- next row f4: STFT
def stft_step(wl: int, overlap_percent: float) -> tuple[int, int]:
    """_spectral_methods.py:248-249: NOTE the overlap is rounded here (+0.5), unlike `_welch`."""
    overlap_samples = int(overlap_percent / 100 * wl + 0.5)
    return overlap_samples, wl - overlap_samples


def stft(x, fs, wl=1024, window="hann", overlap_percent=50.0, fft_length=None, detrend=False,
         padding=True, scaling="fftbackward"):
    """`_stft` (_spectral_methods.py:176-282) -> (time_s, freqs_hz, stft (frequency, time, channel)).

    Frames like Welch's (zero-padded tail, window THEN mean removal), optional zero padding of
    `overlap_samples` at both ends, rfft of length `fft_length` (crop or zero-pad) with the scaling's
    norm; physical scalings: DC (and Nyquist for even lengths) divided by sqrt(2), |.|^2 for the
    power scalings, times the scaling factor.  The frequency vector uses the WINDOW length (:281).
    """
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    n = wl if fft_length is None else int(fft_length)
    w = welch_window(window, wl)
    overlap_samples, step = stft_step(wl, overlap_percent)
    if padding:
        z = np.zeros((overlap_samples, x.shape[1]))
        x = np.concatenate([z, x, z], axis=0)
    fr = framed(x, wl, step)
    fr *= w[:, None, None]
    if detrend:
        fr -= np.mean(fr, axis=0)
    out = np.fft.rfft(fr, axis=0, n=n, norm=fft_norm(scaling))
    sc = scaling.lower()
    if sc in PHYSICAL_SCALINGS:
        out[0] /= 2**0.5
        if n % 2 == 0:
            out[-1] /= 2**0.5
        if sc not in AMPLITUDE_SCALINGS:
            out = np.abs(out) ** 2.0
        out = out * scaling_factor(sc, n, fs, w)
    time_s = np.linspace(0, x.shape[0] / fs, out.shape[1])
    return time_s, np.fft.rfftfreq(wl, 1 / fs), out


# --------------------------------------------------------------------------- stage 2
def distance_matrix(x: np.ndarray, y: np.ndarray) -> np.ndarray:
    """helpers/other.py:131-153 (expansion formula, NOT direct differences)."""
    x = np.atleast_2d(np.asarray(x, dtype=np.float64))
    y = np.atleast_2d(np.asarray(y, dtype=np.float64))
    return np.sqrt(np.sum(x**2, axis=1, keepdims=True) + np.sum(y.T**2, axis=0, keepdims=True)
                   - 2 * x @ y.T)


def array_center(mic_xyz: np.ndarray) -> tuple[int, np.ndarray]:
    """beamforming/beamforming.py:517-541: the mic closest to the centroid."""
    center = np.mean(mic_xyz, axis=0)
    d = distance_matrix(mic_xyz, center[None, :])[:, 0]
    ind = int(np.argmin(d))
    return ind, mic_xyz[ind]


def steering_vector(wave_number, grid_xyz, mic_xyz, formulation="truelocation") -> np.ndarray:
    """beamforming/beamforming.py:1515-1703 -> complex128 (F, M, G)."""
    k = np.atleast_1d(np.asarray(wave_number, dtype=np.float64))
    N = mic_xyz.shape[0]
    _, pc = array_center(mic_xyz)
    rt0 = distance_matrix(grid_xyz, pc[None, :])[:, 0]          # (G,)
    rti = distance_matrix(grid_xyz, mic_xyz).T                  # (M, G)
    phase = np.exp(-1j * k[:, None, None] * (rti[None] - rt0[None, None, :]))
    formulation = formulation.lower()
    if formulation == "classic":
        return 1 / N * phase
    if formulation == "inverse":
        return rti[None] / N / rt0[None, None, :] * phase
    rtj = np.sum(1 / distance_matrix(mic_xyz, grid_xyz) ** 2, axis=0)   # (G,)
    if formulation == "truepower":
        return 1 / rt0[None, None, :] / rti[None] / rtj[None, None, :] * phase
    if formulation == "truelocation":
        return 1 / rti[None] / np.sqrt(N * rtj[None, None, :]) * phase
    raise ValueError(formulation)


def fractional_octave_band(fc: float, fraction) -> np.ndarray:
    """helpers/other.py:156-178."""
    if fraction == 0:
        return np.array([fc, fc])
    return np.array([fc * 2 ** (-1 / fraction / 2), fc * 2 ** (1 / fraction / 2)])


def band_bins(f: np.ndarray, fc: float, fraction) -> tuple[int, int]:
    """beamforming.py:849-856 + helpers/other.py:9-31: half-open bin range [id1, id2)."""
    lo, hi = fractional_octave_band(fc, fraction)
    id1 = int(np.argmin(np.abs(lo - f)))
    id2 = int(np.argmin(np.abs(hi - f)))
    if id1 == id2:
        id2 += 1
    return id1, id2


def simpson_weights(n: int, dx: float) -> np.ndarray:
    """Weights w with simpson(y, dx=dx) == y @ w (scipy.integrate.simpson, beamforming.py:876)."""
    return _scipy_simpson(np.eye(n), dx=dx, axis=1)


def _prepare_csm(csm, M, remove_diag):
    csm = np.array(csm, dtype=np.complex128, copy=True)
    if remove_diag:
        csm *= M / (M - 1)                              # beamforming.py:843
        idx = np.arange(M)
        csm[:, idx, idx] = 0                            # :844-845 (all bins)
    return csm


def das_map_loop(f, csm, mic_xyz, grid_xyz, fc, fraction=3, remove_diag=True, c=343.0,
                 formulation="truelocation", grid_subset=None):
    """`BeamformerDASFrequency.get_beamformer_map` keeping the reference's loops
    (beamforming.py:841-878): dense h, one multi_dot per (grid point, bin).

    ``grid_subset`` (indices) evaluates only those grid points (bounded sample for timing).
    Returns the flat float64 map (len G or len(grid_subset)).
    """
    M = mic_xyz.shape[0]
    csm = _prepare_csm(csm, M, remove_diag)
    id1, id2 = band_bins(f, fc, fraction)
    fsel = f[id1:id2]
    csm = csm[id1:id2]
    k = fsel * np.pi * 2 / c
    g = grid_xyz if grid_subset is None else grid_xyz[np.asarray(grid_subset)]
    h = steering_vector(k, g, mic_xyz, formulation)
    h_H = np.swapaxes(h, 1, 2).conjugate()
    out = np.zeros((g.shape[0], len(fsel)))
    for gi in range(g.shape[0]):
        for fi in range(len(fsel)):
            out[gi, fi] = np.linalg.multi_dot([h_H[fi, gi, :], csm[fi], h[fi, :, gi]]).real
    if remove_diag:
        out[out < 0] = 0
    if len(fsel) > 1:
        return _scipy_simpson(out, dx=fsel[1] - fsel[0], axis=1)
    return out[:, 0]


def das_map(f, csm, mic_xyz, grid_xyz, fc, fraction=3, remove_diag=True, c=343.0,
            formulation="truelocation", chunk_points: int = 2048, return_per_bin=False):
    """Vectorised DAS map: P[g,f] = Re(h^H C h) as batched matmul + einsum, grid-chunked so
    the dense steering array never exceeds chunk_points * F * M complex128."""
    M = mic_xyz.shape[0]
    csm = _prepare_csm(csm, M, remove_diag)
    id1, id2 = band_bins(f, fc, fraction)
    fsel = f[id1:id2]
    csm = csm[id1:id2]
    k = fsel * np.pi * 2 / c
    G = grid_xyz.shape[0]
    P = np.empty((G, len(fsel)))
    for g0 in range(0, G, chunk_points):
        h = steering_vector(k, grid_xyz[g0:g0 + chunk_points], mic_xyz, formulation)  # (F,M,g)
        t = np.matmul(csm, h)                                                         # (F,M,g)
        P[g0:g0 + chunk_points] = np.einsum("fmg,fmg->gf", h.conjugate(), t).real
    if remove_diag:
        P[P < 0] = 0
    if return_per_bin:
        return P
    if len(fsel) > 1:
        return P @ simpson_weights(len(fsel), fsel[1] - fsel[0])
    return P[:, 0]


# --------------------------------------------------------------------------- "next" rows
def _band(f, csm, fc, fraction):
    id1, id2 = band_bins(f, fc, fraction)
    return f[id1:id2], np.asarray(csm)[id1:id2]


def _quadratic_form(h, A):
    """P[g, f] = Re(h[f,:,g]^H A[f] h[f,:,g])."""
    return np.einsum("fmg,fmg->gf", h.conjugate(), np.matmul(A, h)).real


def _integrate(P, fsel):
    if len(fsel) > 1:
        return _scipy_simpson(P, dx=fsel[1] - fsel[0], axis=1)
    return P[:, 0]


def mvdr_map(f, csm, mic_xyz, grid_xyz, fc, fraction=3, c=343.0, formulation="truelocation"):
    """`BeamformerMVDR.get_beamformer_map` (beamforming.py:1238-1314): 1 / Re(h^H C^-1 h) per bin."""
    fsel, C = _band(f, csm, fc, fraction)
    h = steering_vector(fsel * np.pi * 2 / c, grid_xyz, mic_xyz, formulation)
    return _integrate(1.0 / _quadratic_form(h, np.linalg.inv(C)), fsel)


def functional_map(f, csm, mic_xyz, grid_xyz, fc, fraction=3, gamma=10, c=343.0, formulation="truelocation"):
    """`BeamformerFunctional.get_beamformer_map` (beamforming.py:1139-1220): the CSM is replaced by
    u s^(1/gamma) v^H (SVD), and the beamformer output is raised to gamma around the steering norm."""
    fsel, C = _band(f, csm, fc, fraction)
    h = steering_vector(fsel * np.pi * 2 / c, grid_xyz, mic_xyz, formulation)
    u, s, vh = np.linalg.svd(C)
    Cg = np.matmul(u * (s ** (1 / gamma))[:, None, :], vh)
    P = _quadratic_form(h, Cg)
    norm = np.einsum("fmg,fmg->gf", h.conjugate(), h).real
    return _integrate((P / norm) ** gamma * norm, fsel)


def orthogonal_map(f, csm, mic_xyz, grid_xyz, fc, fraction=3, number_eigenvalues=None, c=343.0,
                   formulation="truelocation"):
    """`BeamformerOrthogonal.get_beamformer_map` (beamforming.py:1025-1124): per bin and per retained
    eigenpair (largest first) the grid point maximising |h^H v|^2 receives that value times the
    eigenvalue (assignment: a later eigenpair landing on the same point overwrites it)."""
    M = mic_xyz.shape[0]
    E = M // 2 if number_eigenvalues is None else number_eigenvalues
    fsel, C = _band(f, csm, fc, fraction)
    h = steering_vector(fsel * np.pi * 2 / c, grid_xyz, mic_xyz, formulation)
    out = np.zeros((grid_xyz.shape[0], len(fsel)))
    for fi in range(len(fsel)):
        w, v = np.linalg.eigh(C[fi])
        for e in range(E):
            prod = h[fi].conj().T @ v[:, -e - 1]                     # (G,)
            em = (prod * prod.conj()).real
            gi = int(np.argmax(em))
            out[gi, fi] = em[gi] * w[-e - 1]
    return _integrate(out, fsel)


def _clean_sc_deconvolve(dirty, csm, h, maximum_iterations, remove_diag, safety_factor):
    """beamforming/_beamforming.py:194-297 for one bin.  h: (M, G)."""
    D0, D1 = 2 * csm, csm.copy()
    m = np.array(dirty, dtype=np.float64, copy=True)
    second = np.zeros_like(m)
    for _ in range(maximum_iterations):
        gi = int(np.argmax(m))
        pmax = m[gi]
        second[gi] += pmax * safety_factor
        if np.linalg.norm(D1, ord=1) >= np.linalg.norm(D0, ord=1):
            break
        w_max = h[:, gi]
        h_ = w_max.copy()
        w2 = w_max.conjugate() * w_max
        D_ = D1 @ w_max / pmax
        for _ in range(20):
            H = h_.conjugate() * h_
            h_ = (D_ + H * w_max) / np.sqrt(1 + H @ w2)
        Gm = np.outer(h_, h_.conjugate()) * pmax
        if remove_diag:
            np.fill_diagonal(Gm, 0)
        m -= np.einsum("mg,mg->g", h.conjugate(), Gm @ h).real * safety_factor
        D0, D1 = D1, D1 - safety_factor * Gm
    return second


def clean_sc_map(f, csm, mic_xyz, grid_xyz, fc, fraction=3, maximum_iterations=None, safety_factor=0.5,
                 remove_diag=False, c=343.0, formulation="truelocation"):
    """`BeamformerCleanSC.get_beamformer_map` (beamforming.py:897-1007)."""
    M = mic_xyz.shape[0]
    maxit = 2 * M if maximum_iterations is None else maximum_iterations
    fsel, C = _band(f, csm, fc, fraction)
    C = np.array(C, dtype=np.complex128, copy=True)
    if remove_diag:
        idx = np.arange(M)
        C[:, idx, idx] = 0
    h = steering_vector(fsel * np.pi * 2 / c, grid_xyz, mic_xyz, formulation)
    dirty = _quadratic_form(h, C)
    out = np.zeros_like(dirty)
    for fi in range(len(fsel)):
        out[:, fi] = _clean_sc_deconvolve(dirty[:, fi], C[fi], h[fi], maxit, remove_diag, safety_factor).real
    return _integrate(out, fsel)
