"""K0 + K1 (gather, frame, window, detrend, real FFT) through the C ABI against numpy."""

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from dsptoolbox_b200 import _native as N  # noqa: E402
from oracle import restatement as R  # noqa: E402


def run_k0_k1(x, wl, step, window, detrend, scale, prec, transposed_storage=False):
    lib = N.lib()
    dev = torch.device("cuda")
    L, M = x.shape
    K, _ = R.frame_count(wl, step, L)
    Kp = (K + 15) // 16 * 16
    Lp = (Kp - 1) * step + wl
    Lp += (-Lp) % 4
    rdt, cdt = (torch.float64, torch.complex128) if prec == N.F64 else (torch.float32, torch.complex64)
    if transposed_storage:
        xd = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        s_t, s_m = 1, L
    else:
        xd = torch.from_numpy(np.ascontiguousarray(x)).to(dev)
        s_t, s_m = M, 1
    code = N.F64 if x.dtype == np.float64 else N.F32
    xc = torch.empty((M, Lp), dtype=rdt, device=dev)
    st = N.stream_handle()
    N.check(lib.b2_prepare_channels(N.ptr(xd), code, s_t, s_m, 0, L, M, N.ptr(xc), Lp, prec, st))
    tw = torch.empty(wl, dtype=cdt, device=dev)
    N.check(lib.b2_fft_twiddles(wl, prec, N.ptr(tw), st))
    win = torch.from_numpy(window).to(dev).to(rdt)
    nb = wl // 2 + 1
    ng = wl // 8 + 1
    assert lib.b2_welch_spectra_elems(wl, M, Kp) == Kp // 16 * ng * M * 64
    X = torch.full((Kp // 16, ng, M, 8, 4, 2), float("nan"), dtype=cdt, device=dev)
    N.check(lib.b2_welch_rfft(N.ptr(xc), Lp, M, wl, step, K, N.ptr(win), N.ptr(tw), int(detrend), scale,
                              N.ptr(X), Kp, prec, st))
    torch.cuda.synchronize()
    # device layout [Kp/16][slot group][M][8 frame pairs][4 slots][2 frames] -> (slots, M, Kp) -> the bins' slots
    slots = np.array([lib.b2_welch_bin_slot(wl, f) for f in range(nb)])
    assert len(set(slots)) == nb and slots.max() < 4 * ng
    Xn = X.cpu().numpy().transpose(1, 4, 2, 0, 3, 5).reshape(4 * ng, M, Kp)[slots]
    return xc.cpu().numpy(), Xn, K, tw.cpu().numpy()


@pytest.mark.parametrize("wl", [8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384])
@pytest.mark.parametrize("prec", [N.F64, N.F32])
def test_rfft_all_lengths(wl, prec):
    rng = np.random.default_rng(wl)
    M = 3
    step = wl // 2
    L = step * 5 + 3
    x = rng.standard_normal((L, M))
    w = R.welch_window("hann", wl)
    xc, X, K, tw = run_k0_k1(x, wl, step, w, True, 1.0, prec)
    assert np.array_equal(xc[:, :L], x.T.astype(xc.dtype)) and np.all(xc[:, L:] == 0)
    ref = R.channel_spectra(x, wl, "hann", 50, True, "fftbackward")      # (nb, K, M)
    got = np.swapaxes(X[:, :, :K], 1, 2)
    tol = 1e-12 if prec == N.F64 else 2e-5
    scale = np.abs(ref[1:]).max()
    assert np.max(np.abs(got[1:] - ref[1:])) <= tol * scale
    assert np.max(np.abs(got[0])) <= (1e-12 if prec == N.F64 else 1e-4) * scale   # de-meaned DC bin
    assert np.all(X[:, :, K:] == 0)                                      # padded frames are zero
    t = np.arange(wl)
    assert np.max(np.abs(tw - np.exp(-2j * np.pi * t / wl))) < (1e-15 if prec == N.F64 else 1e-7)


@pytest.mark.parametrize("wl,overlap,window,detrend,scaling", [
    (64, 33, "blackman", True, "fftforward"),          # odd step (43): scalar load path
    (128, 75, "hamming", False, "fftorthogonal"),
    (512, 90, "flattop", True, "fftbackward"),
    (256, 0, "boxcar", False, "fftbackward"),
])
def test_rfft_parameter_variants(wl, overlap, window, detrend, scaling):
    rng = np.random.default_rng(1)
    x = rng.standard_normal((3001, 5)) + 0.3
    step = R.welch_step(wl, overlap)
    w = R.welch_window(window, wl)
    fft_scale = {"fftbackward": 1.0, "fftforward": 1.0 / wl, "fftorthogonal": wl ** -0.5}[scaling]
    ref = R.channel_spectra(x, wl, window, overlap, detrend, scaling)
    for transposed in (False, True):
        _, X, K, _ = run_k0_k1(x, wl, step, w, detrend, fft_scale, N.F64, transposed)
        got = np.swapaxes(X[:, :, :K], 1, 2)
        lo = 1 if detrend else 0
        assert np.max(np.abs(got[lo:] - ref[lo:])) <= 1e-12 * np.abs(ref).max()


def test_float32_input_and_strided_views():
    rng = np.random.default_rng(2)
    x32 = rng.standard_normal((2000, 4)).astype(np.float32)
    w = R.welch_window("hann", 128)
    _, X, K, _ = run_k0_k1(x32, 128, 64, w, True, 1.0, N.F64)
    ref = R.channel_spectra(x32.astype(np.float64), 128, "hann", 50, True, "fftbackward")
    assert np.max(np.abs(np.swapaxes(X[:, :, :K], 1, 2)[1:] - ref[1:])) <= 1e-12 * np.abs(ref).max()
