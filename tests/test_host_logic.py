"""Host-side mirror of the reference interface: no GPU needed."""

import numpy as np
import pytest

import dsptoolbox_b200 as dsp
from dsptoolbox_b200.beamforming import simpson_weights
from dsptoolbox_b200.helpers.other import (_compute_number_frames, _euclidean_distance_matrix,
                                           _get_fractional_octave_bandwidth, _pad_trim,
                                           find_nearest_points_index_in_vector)
from oracle import restatement as R

x = np.arange(0, 1.1, 0.25)
xx, yy, zz = np.meshgrid(x, x, x, indexing="ij")
POINTS = dict(x=xx.flatten(), y=yy.flatten(), z=zz.flatten())


def test_grid_like_reference_test():
    # mirrors reference tests/test_beamforming.py:15-29
    g = dsp.beamforming.Grid(positions=POINTS)
    assert np.all([0, 1] == g.extent["x"])
    assert g.number_of_points == len(x) ** 3
    d = g.get_distances_to_point([0, 0, 0])
    assert d.shape == (125,) and d[0] == 0
    idx, coord = g.find_nearest_point([-0.2, 0.1, -1])
    assert idx == 0 and np.all(coord == 0)
    assert g.reconstruct_map_shape(d) is d


def test_regular_grids():
    # mirrors reference tests/test_beamforming.py:33-47 (+ layout facts of SURVEY.md 3.3)
    g = dsp.beamforming.Regular2DGrid(line1=x, line2=x[:3], dimensions=("x", "y"), value3=2)
    assert g.number_of_points == 15 and g.ndim == 2 and g.dim == ["x", "y"]
    c = g.coordinates
    assert np.all(c[:3, 0] == x[0]) and np.all(c[:3, 1] == x[:3]) and np.all(c[:, 2] == 2)   # line1 slowest
    assert g.reconstruct_map_shape(np.arange(15.0)).shape == (5, 3)
    g2 = dsp.beamforming.Regular2DGrid(x, x[:2], ("Z", "x"), 0.5)
    assert np.all(g2.coordinates[:, 1] == 0.5) and np.all(g2.coordinates[:2, 2] == x[0])
    g3 = dsp.beamforming.Regular3DGrid(x, x[:2], x[:3])
    assert g3.number_of_points == 30
    assert np.all(g3.coordinates[:3, 2] == x[:3])                                             # z fastest
    assert g3.reconstruct_map_shape(np.arange(30.0)).shape == (5, 2, 3)
    gl = dsp.beamforming.LineGrid(line=x, dimension="y", value2=0.3, value3=1)
    assert np.all(gl.coordinates[:, 1] == x) and np.all(gl.coordinates[:, 2] == 0.3) and np.all(gl.coordinates[:, 0] == 1)
    with pytest.raises(AssertionError):
        dsp.beamforming.Regular2DGrid(x, x, ("x", "x"), 1)
    with pytest.raises(AssertionError):
        g3.reconstruct_map_shape(np.zeros(7))


def test_mic_array():
    m = dsp.beamforming.MicArray(POINTS)
    ind, pc = R.array_center(m.coordinates)
    assert m.array_center_channel_number == ind and np.all(m.array_center_coordinates == pc)
    assert np.isclose(m.aperture, np.sqrt(3.0)) and np.isclose(m.min_distance, 0.25)
    lo, hi = m.get_maximum_frequency_range()
    assert np.isclose(lo, 4 * 343 / np.sqrt(3.0)) and np.isclose(hi, 343 / 0.25 / 2)
    c = m.coordinates
    c[0, 0] = 99.0
    assert m.coordinates[0, 0] == 0.0          # the getter hands out copies


def test_steering_vector_object():
    st = dsp.beamforming.SteeringVector()
    assert st.get_vector is dsp.beamforming.true_location_steering
    for t in dsp.beamforming.SteeringVectorType:
        dsp.beamforming.SteeringVector(formulation=t)
    with pytest.raises(ValueError):
        dsp.beamforming.SteeringVector(formulation="classic")


def test_helpers():
    f = np.fft.rfftfreq(1024, 1 / 48000)
    lo, hi = _get_fractional_octave_bandwidth(4000, 3)
    assert np.allclose([lo, hi], R.fractional_octave_band(4000, 3))
    assert np.all(_get_fractional_octave_bandwidth(500.0, 0) == 500.0)
    ids = find_nearest_points_index_in_vector([lo, hi], f)
    assert tuple(ids) == R.band_bins(f, 4000, 3) == (76, 96)
    assert find_nearest_points_index_in_vector(23.4375 * 1.5, f)[0] == 1          # first minimum on ties
    a = np.random.default_rng(0).standard_normal((5, 3))
    b = np.random.default_rng(1).standard_normal((4, 3))
    assert np.allclose(_euclidean_distance_matrix(a, b), R.distance_matrix(a, b), rtol=1e-15)
    for (wl, step, L) in [(1024, 512, 48000), (4096, 2048, 480000), (4096, 2048, 2880000), (128, 86, 43 * 86), (64, 43, 1000)]:
        assert _compute_number_frames(wl, step, L, True) == R.frame_count(wl, step, L)
    assert _compute_number_frames(1024, 512, 48000, True) == (94, 640)
    assert _compute_number_frames(4096, 2048, 2880000, True) == (1407, 3584)
    v = np.arange(6.0).reshape(3, 2)
    assert _pad_trim(v, 5).shape == (5, 2) and np.all(_pad_trim(v, 5)[3:] == 0)
    assert np.all(_pad_trim(v, 2) == v[:2]) and np.all(_pad_trim(np.arange(4.0), 6)[4:] == 0)


@pytest.mark.parametrize("n", [2, 3, 4, 5, 20, 21, 640, 2048])
def test_simpson_weights(n):
    from scipy.integrate import simpson

    assert np.allclose(simpson_weights(n, 23.4375), simpson(np.eye(n), dx=23.4375, axis=1), rtol=1e-14, atol=1e-12)


def test_scaling_enum_matches_oracle():
    w = R.welch_window("hann", 256)
    for s in dsp.SpectrumScaling:
        assert s.fft_norm() == R.fft_norm(s.name)
        assert s.is_amplitude_scaling() == (s.name.lower() in R.AMPLITUDE_SCALINGS)
        assert s.has_physical_units() == (s.name.lower() in R.PHYSICAL_SCALINGS)
        got = float(np.squeeze(s.get_scaling_factor(256, 8000, w)))
        assert np.isclose(got, R.scaling_factor(s.name, 256, 8000, w), rtol=1e-15)
    assert dsp.Window.Hann.to_scipy_format() == "hann"
    assert dsp.Window.KaiserBesselDerived.with_extra_parameter(3.0).to_scipy_format() == ("kaiser_bessel_derived", 3.0)
    assert dsp.Window.GeneralGaussian.with_extra_parameter((1.5, 7)).to_scipy_format() == ("general_gaussian", 1.5, 7)
    assert np.allclose(dsp.Window.Hann(64, False), R.welch_window("hann", 64))


def test_signal_container_semantics():
    rng = np.random.default_rng(3)
    data = rng.standard_normal((4, 1000))               # (channels, time): gets transposed
    s = dsp.Signal(None, data, 8000)
    assert s.time_data.shape == (1000, 4) and s.number_of_channels == 4 and len(s) == 1000
    assert np.shares_memory(s.time_data, data)           # a view of the caller's array
    assert s.length_samples == 1000 and s.length_seconds == 0.125
    with pytest.raises(AssertionError):
        dsp.Signal(None, data, 8000.0)                   # sampling rate must be int
    with pytest.raises(AssertionError):
        dsp.Signal(None, None, 8000)
    with pytest.raises(AssertionError):
        dsp.Signal(None, rng.standard_normal((2, 3, 4)), 8000)
    with pytest.warns(UserWarning):
        s2 = dsp.Signal(None, 3.0 * np.ones((10, 2)), 8000, constrain_amplitude=True)
    assert np.all(s2.time_data == 1.0) and np.isclose(s2.amplitude_scale_factor, 1 / 3)
    mono = dsp.Signal(None, rng.standard_normal(100), 8000)
    assert mono.time_data.shape == (100, 1)
    with pytest.raises(AssertionError):
        mono.get_csm()                                   # needs >= 2 channels (reference signal.py:969-972)
    assert s.set_spectrum_parameters(window_length_samples=256) is s
    assert s._csm_dirty
    with pytest.raises(AssertionError):
        s.spectrum_scaling = "psd"


def test_beamformer_constructor_checks():
    rng = np.random.default_rng(0)
    s = dsp.Signal(None, rng.standard_normal((500, 125)), 8000)
    ma = dsp.beamforming.MicArray(POINTS)
    g = dsp.beamforming.Grid(POINTS)
    st = dsp.beamforming.SteeringVector()
    dsp.beamforming.BeamformerDASFrequency(s, ma, g, st)
    with pytest.raises(AssertionError):
        dsp.beamforming.BeamformerDASFrequency(s, g, g, st)                   # mic_array must be MicArray exactly
    with pytest.raises(AssertionError):
        dsp.beamforming.BeamformerDASFrequency(s, ma, g, st, c=0)
    with pytest.raises(AssertionError):
        dsp.beamforming.BeamformerDASFrequency(s, ma, ma, st)                 # grid must be a Grid
    with pytest.raises(AssertionError):
        dsp.beamforming.BeamformerDASFrequency(dsp.Signal(None, rng.standard_normal((500, 3)), 8000), ma, g, st)


def test_spectrogram_parameters_and_stft_argument_checks():
    """reference signal.py:706-773 (names, defaults, state flag) and the asserts of _stft (:234-240); everything
    checked here fails before any device work, so it runs without a GPU."""
    from dsptoolbox_b200.standard._spectral_methods import _stft
    s = dsp.Signal(None, np.random.default_rng(0).standard_normal((2000, 2)), 8000)
    p = s._spectrogram_parameters
    assert p == dict(window_length_samples=1024, window_type=dsp.Window.Hann, overlap_percent=50.0,
                     fft_length_samples=None, detrend=False, padding=True, scaling=dsp.SpectrumScaling.FFTBackward)
    s._spectrogram_dirty = False
    assert s.set_spectrogram_parameters() is s and not s._spectrogram_dirty            # unchanged -> still valid
    s.set_spectrogram_parameters(window_length_samples=256, detrend=True)
    assert s._spectrogram_dirty and s._spectrogram_parameters["window_length_samples"] == 256
    x = s.time_data
    with pytest.raises(AssertionError):
        _stft(x, 8000, 100, dsp.Window.Hann, 50.0, None, False, True, dsp.SpectrumScaling.FFTBackward)
    with pytest.raises(AssertionError):
        _stft(x, 8000, 8, dsp.Window.Hann, 50.0, None, False, True, dsp.SpectrumScaling.FFTBackward)       # < 2**4
    with pytest.raises(AssertionError):
        _stft(x, 8000, 256, dsp.Window.Hann, 100.0, None, False, True, dsp.SpectrumScaling.FFTBackward)
    with pytest.raises(AssertionError):
        _stft(x, 8000, 1 << 17, dsp.Window.Hann, 50.0, None, False, True, dsp.SpectrumScaling.FFTBackward)  # > 2**16
    with pytest.raises(AssertionError):
        _stft(x, 8000, 256, dsp.Window.Hann, 50.0, 1, False, True, dsp.SpectrumScaling.FFTBackward)         # FFT length < 2
    from dsptoolbox_b200.standard._spectral_methods import spectrum_fft_device
    with pytest.raises(NotImplementedError):                  # smoothing is checked before any device work
        spectrum_fft_device(x, 8000, dsp.SpectrumScaling.FFTBackward, True, 3)
    # FFTBackward -> output scaling conversion of `_csm_fft` (reference enums.py:131-181): amplitude/amplitude is a
    # ratio of linear factors, amplitude -> power multiplies SQUARED data
    b = dsp.SpectrumScaling.FFTBackward
    assert np.allclose(b.conversion_factor(dsp.SpectrumScaling.FFTForward, 100, 8000, None), 1 / 100)
    assert np.allclose(b.conversion_factor(dsp.SpectrumScaling.PowerSpectrum, 100, 8000, None), 2 / 100 ** 2)
    assert np.allclose(b.conversion_factor(dsp.SpectrumScaling.AmplitudeSpectralDensity, 100, 8000, None), (2 / 100 / 8000) ** 0.5)


def test_spectrum_holder_and_transfer_function_argument_checks():
    """reference classes/spectrum.py:139-260, 871-885 (field validation) and the asserts of
    compute_transfer_function (transfer_functions.py:453-466), all raised before any device work."""
    from dsptoolbox_b200.transfer_functions import TransferFunctionType, compute_transfer_function
    f = np.linspace(0, 4000, 9)
    sp = dsp.Spectrum(f, np.ones((3, 9)) + 0j)                       # transposed input is turned
    assert sp.spectral_data.shape == (9, 3) and sp.number_of_channels == 3 and sp.number_frequency_bins == 9
    sp.set_coherence(np.zeros((9, 3)))
    with pytest.raises(AssertionError):
        sp.set_coherence(np.zeros((9, 2)))
    with pytest.raises(AssertionError):
        sp.set_coherence(np.zeros((9, 3)) + 0j)
    with pytest.raises(AssertionError):
        dsp.Spectrum(f[::-1], np.ones((9, 1)))                        # not ascending
    with pytest.raises(AssertionError):
        dsp.Spectrum(f, -np.ones((9, 1)))                             # negative magnitude
    with pytest.raises(AssertionError):
        dsp.Spectrum(f, np.ones((7, 2)))                              # length mismatch
    assert [m.name for m in TransferFunctionType] == ["H1", "H2", "H3"]
    x = np.random.default_rng(0).standard_normal((4000, 2))
    a, b = dsp.Signal(None, x, 8000), dsp.Signal(None, x[:3000], 8000)
    with pytest.raises(AssertionError, match="lengths"):
        compute_transfer_function(b, a, 256)
    with pytest.raises(AssertionError, match="Sampling rates"):
        compute_transfer_function(dsp.Signal(None, x, 16000), a, 256)
    with pytest.raises(AssertionError, match="Channel number"):
        compute_transfer_function(dsp.Signal(None, np.c_[x, x[:, :1]], 8000), a, 256)
    with pytest.raises(ValueError):
        compute_transfer_function(a, a, 256, "H1")


def test_fractional_delay_filters_and_source_checks():
    """Host side of the source simulator (next row f2): the vectorised filter design against the oracle's
    one-delay restatement of reference _standard_backend.py:430-493, and MonopoleSource's constructor asserts."""
    from dsptoolbox_b200.standard._standard_backend import _fractional_delay_filter, fractional_delay_filters
    rng = np.random.default_rng(0)
    d = np.concatenate([rng.uniform(0, 400, 500), [0.3, 0.5, 0.93, 15.0, 1e-4, 7.5, 2.999999]])
    shift, taps = fractional_delay_filters(d)
    assert taps.shape == (len(d), 31) and shift.dtype == np.int64
    for i, v in enumerate(d):
        s0, h0 = R.fractional_delay_filter(v)
        assert s0 == shift[i] and np.max(np.abs(h0 - taps[i])) < 4e-15
    for order in (7, 12, 31):                       # odd orders take the other centring rule
        s0, h0 = R.fractional_delay_filter(4.7, order, 40.0)
        s1, h1 = _fractional_delay_filter(4.7, order, 40.0)
        assert s0 == s1 and np.max(np.abs(h0 - h1)) < 4e-15
    sig = dsp.Signal(None, rng.standard_normal(100), 8000)
    assert dsp.beamforming.MonopoleSource(sig, [[0, 0, 1]]).coordinates.shape == (3,)
    with pytest.raises(AssertionError):
        dsp.beamforming.MonopoleSource(dsp.Signal(None, rng.standard_normal((100, 2)), 8000), [0, 0, 1])
    with pytest.raises(AssertionError):
        dsp.beamforming.MonopoleSource(sig, [0, 1])
    with pytest.raises(AssertionError):
        dsp.beamforming.mix_sources_on_array([], dsp.beamforming.MicArray(POINTS))
