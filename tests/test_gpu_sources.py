"""Next row f2 parity: `MonopoleSource.get_signals_on_array` and `mix_sources_on_array` on the GPU (one
fractional-delay FIR kernel for all microphones and sources) against the golden vectors of the real reference and
against the oracle, and the reference's own test scenario (simulate -> CSM -> map peaks at the source)."""

import os
import warnings

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import dsptoolbox_b200 as dsp  # noqa: E402
from dsptoolbox_b200 import engine  # noqa: E402
from oracle import restatement as R  # noqa: E402

# the reference convolves with an overlap-add FFT, the kernel directly: both fp64, differences ~1e-16 of the peak
REL_TO_MAX = 1e-12


def _mics(xyz):
    return dsp.beamforming.MicArray(dict(x=xyz[:, 0], y=xyz[:, 1], z=xyz[:, 2]))


def test_golden_sources(golden_dir):
    g = np.load(os.path.join(golden_dir, "sources_on_array.npz"))
    fs, mics, c = int(g["fs"]), _mics(g["mic"]), float(g["c"])
    src = [dsp.beamforming.MonopoleSource(dsp.Signal(None, g[f"x_{i}"].copy(), fs), g["src"][i]) for i in range(int(g["n_src"]))]
    for i, s in enumerate(src):
        out = s.get_signals_on_array(mics, c)
        ref = g[f"on_array_{i}"]
        assert out.time_data.is_cuda and out.sampling_rate_hz == fs
        got = out.time_data.cpu().numpy()
        assert got.shape == ref.shape and got.dtype == ref.dtype
        assert np.max(np.abs(got - ref)) <= REL_TO_MAX * np.max(np.abs(ref)), i
    for tag in ("equal", "ragged", "ragged2"):
        lst = [src[i] for i in g[f"mix_{tag}_ids"]]
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            got = dsp.beamforming.mix_sources_on_array(lst, mics, c).time_data.cpu().numpy()
        ref = g[f"mix_{tag}"]
        assert (len(w) > 0) == (tag != "equal")                       # the reference warns when it trims
        assert got.shape == ref.shape and np.max(np.abs(got - ref)) <= REL_TO_MAX * np.max(np.abs(ref)), tag
        assert len(lst) == len(g[f"mix_{tag}_ids"])                   # the caller's list is left alone
    one = dsp.beamforming.mix_sources_on_array(src[0], mics, c).time_data.cpu().numpy()      # a bare source
    assert np.max(np.abs(one - g["on_array_0"])) <= REL_TO_MAX * np.max(np.abs(g["on_array_0"]))


@pytest.mark.parametrize("L,M,S", [(200000, 64, 1), (99991, 33, 3), (127, 5, 2), (128, 2, 1), (300, 256, 2)])
def test_against_oracle(L, M, S):
    """Ragged lengths against the 128-sample strips of the kernel, microphone counts that are not a multiple of the
    warp, several sources, fp32 and device-resident emitted signals -- against the oracle's direct convolution."""
    rng = np.random.default_rng(L + M)
    fs = 48000
    mic = rng.uniform(-0.5, 0.5, (M, 3)) * [1, 1, 0]
    src = rng.uniform(-1, 1, (S, 3)) + [0, 0, 2.0]
    if L < 1000:
        src[:, 2] = 0.3                              # the delay has to stay below the signal length
        mic *= 0.2
        src[:, :2] *= 0.1
    sig = [rng.standard_normal(L) for _ in range(S)]
    sources = [dsp.beamforming.MonopoleSource(dsp.Signal(None, sig[i].copy(), fs), src[i]) for i in range(S)]
    got = dsp.beamforming.mix_sources_on_array(sources, _mics(mic)).time_data.cpu().numpy()
    ref = R.mix_sources_on_array(sig, list(src), fs, mic, 343.0)
    assert got.shape == ref.shape == (L, M)
    assert np.max(np.abs(got - ref)) <= REL_TO_MAX * np.max(np.abs(ref))
    # device-resident fp32 emitted signal
    x32 = torch.from_numpy(sig[0].astype(np.float32)).cuda()
    n0 = engine.launches
    got = dsp.beamforming.MonopoleSource(dsp.Signal(None, x32, fs), src[0]).get_signals_on_array(_mics(mic)).time_data
    assert engine.launches == n0 + 1 and got.dtype == torch.float64
    ref = R.monopole_on_array(sig[0].astype(np.float32), fs, src[0], mic, 343.0)
    assert np.max(np.abs(got.cpu().numpy() - ref)) <= REL_TO_MAX * np.max(np.abs(ref))


def test_simulated_source_is_found_by_the_map():
    """The reference's own scenario (tests/test_beamforming.py:130-160): noise from a monopole source, simulated on
    a planar array, located by the delay-and-sum map -- everything from the emitted signal to the map on the GPU."""
    rng = np.random.default_rng(1)
    fs = 10000
    lin = np.arange(-0.3, 0.31, 0.1)
    xx, yy = np.meshgrid(lin, lin, indexing="ij")
    mics = dsp.beamforming.MicArray(dict(x=xx.ravel(), y=yy.ravel(), z=np.zeros(xx.size)))
    source = dsp.beamforming.MonopoleSource(dsp.Signal(None, rng.standard_normal(2 * fs), fs), [0.1, 0.3, 0.5])
    s = source.get_signals_on_array(mics)
    assert s.number_of_channels == 49 and s.time_data.is_cuda
    grid = dsp.beamforming.Regular2DGrid(np.arange(-0.4, 0.41, 0.1), np.arange(-0.5, 0.51, 0.1), ["x", "y"], value3=0.5)
    bf = dsp.beamforming.BeamformerDASFrequency(s, mics, grid, dsp.beamforming.SteeringVector())
    m = bf.get_beamformer_map(2000, 3)
    ix, iy = np.unravel_index(np.argmax(m), m.shape)
    assert np.isclose(grid.coordinates.reshape(9, 11, 3)[ix, iy, 0], 0.1) and np.isclose(grid.coordinates.reshape(9, 11, 3)[ix, iy, 1], 0.3)


def test_argument_checks():
    rng = np.random.default_rng(2)
    mic = rng.uniform(-0.5, 0.5, (4, 3))
    with pytest.raises(AssertionError):
        dsp.beamforming.MonopoleSource(dsp.Signal(None, rng.standard_normal((100, 2)), 8000), [0, 0, 1])
    with pytest.raises(AssertionError):
        dsp.beamforming.MonopoleSource(dsp.Signal(None, rng.standard_normal(100), 8000), [0, 0])
    far = dsp.beamforming.MonopoleSource(dsp.Signal(None, rng.standard_normal(100), 8000), [0, 0, 50.0])
    with pytest.raises(AssertionError):
        far.get_signals_on_array(_mics(mic))            # delay longer than the signal
    with pytest.raises(AssertionError):
        dsp.beamforming.mix_sources_on_array([], _mics(mic))
    with pytest.raises(AssertionError):
        dsp.beamforming.mix_sources_on_array([far, "source"], _mics(mic))
