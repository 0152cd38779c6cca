"""Grids, microphone array, steering vectors and the frequency-domain delay-and-sum beamformer.

Call-compatible with reference beamforming/beamforming.py for Grid (:35-75), Regular2DGrid
(:78-173), Regular3DGrid (:218-295), LineGrid (:366-422), MicArray (:425-602), SteeringVector
(:605-647), BaseBeamformer/BeamformerGridded (:650-796) and BeamformerDASFrequency (:799-880).
`get_beamformer_map` keeps the reference's side effects (stdout banner, `self.map`,
`self.f_range_hz`, returned copy) but runs on the GPU: the CSM stays on the device, steering
vectors are synthesised inside the map kernel, and the grid may be sharded over ranks (dist.py).
BeamformerMVDR (:1223-1314), BeamformerFunctional (:1127-1220) and BeamformerOrthogonal (:1010-1124)
and BeamformerCleanSC (:883-1007) reuse the same map kernel on per-bin transformed matrices
(SURVEY.md section 8f.1).  MonopoleSource (:1397-1458) and mix_sources_on_array (:1461-1511) -- the data path in
front of stage 1 -- run as one fractional-delay kernel (csrc/delay_fir.cu) and leave the
array recording on the device.  BeamformerDASTime is outside the scope.
"""

from __future__ import annotations

import numpy as np

from .. import engine
from .. import _native as N
from ..classes import Signal
from ..helpers.other import _get_fractional_octave_bandwidth, find_nearest_points_index_in_vector
from ._beamforming import BasePoints
from .enums import SteeringVectorType

nxs = np.newaxis


# =========================================================================== grids
class Grid(BasePoints):
    """Arbitrary set of grid points (right-handed cartesian coordinates in metres)."""

    def __init__(self, positions: dict):
        super().__init__(positions)

    def reconstruct_map_shape(self, map):
        """Identity for a plain grid; regular grids reshape the flat map to their lattice."""
        return map


class Regular2DGrid(Grid):
    """Rectangular lattice line1 x line2 in the plane spanned by `dimensions`, third coordinate fixed."""

    def __init__(self, line1, line2, dimensions, value3):
        assert len(dimensions) == 2, (
            "dimensions must contain exactly two strings specifying to which directions line1 and line2 correspond")
        assert len(np.unique(dimensions)) == len(dimensions), "There are repeated dimensions"
        dimensions = [d.lower() for d in dimensions]
        self.extent_dimensions = dimensions
        value3 = np.asarray(value3).squeeze()
        assert value3.ndim == 0, "value3 can only be a single value"
        line1, line2 = np.asarray(line1).squeeze(), np.asarray(line2).squeeze()
        self.original_lengths = (len(line1), len(line2))
        a, b = np.meshgrid(line1, line2, indexing="ij")       # line1 is the slow axis
        third = [ax for ax in ("x", "y", "z") if ax not in dimensions][0]
        super().__init__({dimensions[0]: a.ravel(), dimensions[1]: b.ravel(),
                          third: np.full(a.size, float(value3))})

    def reconstruct_map_shape(self, map_vector):
        assert map_vector.ndim == 1, "The passed map should be a vector (flattened)"
        assert len(map_vector) == self.number_of_points, "Length of passed vector does not match the number of points"
        return map_vector.reshape(self.original_lengths)


class Regular3DGrid(Grid):
    """Box lattice line_x x line_y x line_z (x slowest, z fastest)."""

    def __init__(self, line_x, line_y, line_z):
        self.lines = tuple(np.asarray(v).squeeze() for v in (line_x, line_y, line_z))
        assert all(v.ndim == 1 for v in self.lines), "Shape of lines is invalid"
        self.original_lengths = tuple(len(v) for v in self.lines)
        xx, yy, zz = np.meshgrid(*self.lines, indexing="ij")
        super().__init__(dict(x=xx.ravel(), y=yy.ravel(), z=zz.ravel()))

    def reconstruct_map_shape(self, map_vector):
        assert map_vector.ndim == 1, "The passed map should be a vector (flattened)"
        assert len(map_vector) == self.number_of_points, "Length of passed vector does not match the number of points"
        return map_vector.reshape(self.original_lengths)


class LineGrid(Grid):
    """Points along one coordinate axis; the next axes (cyclic x->y->z->x) are held at value2, value3."""

    def __init__(self, line, dimension: str, value2: float, value3: float):
        line = np.atleast_1d(np.squeeze(line))
        assert line.ndim == 1, "Line has an invalid shape"
        dimension = dimension.lower()
        order = ["x", "y", "z"]
        assert dimension in order, "Dimension should be x, y or z"
        i = order.index(dimension)
        second, third = order[(i + 1) % 3], order[(i + 2) % 3]
        self.extent_dimension = dimension
        super().__init__({dimension: line, second: np.full(len(line), float(value2)),
                          third: np.full(len(line), float(value3))})


# =========================================================================== microphones
class MicArray(BasePoints):
    """Microphone positions; channel order must match the multichannel signal."""

    def __init__(self, positions: dict):
        super().__init__(positions)
        self._center = None
        self._aperture_min = None

    def _pairwise(self):
        if self._aperture_min is None:
            d = np.array(self.get_distances_to_point(self._coordinates), dtype=float, ndmin=2)
            off = ~np.eye(d.shape[0], dtype=bool)
            self._aperture_min = (float(np.max(d[off])), float(np.min(d[off])))
        return self._aperture_min

    @property
    def aperture(self):
        return self._pairwise()[0]

    @property
    def min_distance(self):
        return self._pairwise()[1]

    def _locate_center(self):
        if self._center is None:
            centroid = np.mean(self._coordinates, axis=0)
            ind = int(np.argmin(self.get_distances_to_point(centroid)))
            self._center = (self._coordinates[ind, :].copy(), ind)
        return self._center

    @property
    def array_center_coordinates(self):
        """Coordinates of the microphone closest to the centroid (reference :517-541)."""
        return self._locate_center()[0]

    @property
    def array_center_channel_number(self):
        return self._locate_center()[1]

    def he_to_hz(self, he: float, c: float = 343) -> float:
        return he * c / self.aperture

    def hz_to_he(self, f_hz: float, c: float = 343) -> float:
        return f_hz * self.aperture / c

    def get_maximum_frequency_range(self, lowest_he: float = 4, c: float = 343) -> list:
        return [self.he_to_hz(lowest_he, c=c), c / self.min_distance / 2]


# =========================================================================== steering vectors
_STEER_CODE = {SteeringVectorType.Classic: N.STEER_CLASSIC, SteeringVectorType.Inverse: N.STEER_INVERSE,
               SteeringVectorType.TruePower: N.STEER_TRUE_POWER, SteeringVectorType.TrueLocation: N.STEER_TRUE_LOCATION}


def _dense_steering(code, wave_number, grid, mic):
    k = np.atleast_1d(wave_number)
    assert k.ndim == 1, "Wave number should be a 1D-array"
    h = engine.steering_vector(k, mic._coordinates, mic.array_center_coordinates, grid._coordinates, code)
    return h.cpu().numpy()


def classic_steering(wave_number, grid: Grid, mic: MicArray):
    """Formulation 1: exp(-jk(r_ti - r_t0)) / N -> complex128 (frequency, nmics, ngrid)."""
    return _dense_steering(N.STEER_CLASSIC, wave_number, grid, mic)


def inverse_steering(wave_number, grid: Grid, mic: MicArray):
    """Formulation 2: r_ti / (N r_t0) exp(-jk(r_ti - r_t0))."""
    return _dense_steering(N.STEER_INVERSE, wave_number, grid, mic)


def true_power_steering(wave_number, grid: Grid, mic: MicArray):
    """Formulation 3: exp(-jk(r_ti - r_t0)) / (r_t0 r_ti sum_j r_tj^-2)."""
    return _dense_steering(N.STEER_TRUE_POWER, wave_number, grid, mic)


def true_location_steering(wave_number, grid: Grid, mic: MicArray):
    """Formulation 4: exp(-jk(r_ti - r_t0)) / (r_ti sqrt(N sum_j r_tj^-2))."""
    return _dense_steering(N.STEER_TRUE_LOCATION, wave_number, grid, mic)


_STEER_FUNCS = {SteeringVectorType.Classic: classic_steering, SteeringVectorType.Inverse: inverse_steering,
                SteeringVectorType.TruePower: true_power_steering,
                SteeringVectorType.TrueLocation: true_location_steering}


class SteeringVector:
    """Holds the steering-vector formulation; `get_vector(k, grid=, mic=)` materialises it."""

    def __init__(self, formulation: SteeringVectorType = SteeringVectorType.TrueLocation):
        if formulation not in _STEER_FUNCS:
            raise ValueError("Incorrect formulation. Use either classic, inverse, true power or true location")
        self.formulation = formulation
        self.get_vector = _STEER_FUNCS[formulation]


def _steering_code(st_vec: SteeringVector) -> int:
    for kind, fn in _STEER_FUNCS.items():
        if st_vec.get_vector is fn:
            return _STEER_CODE[kind]
    raise ValueError("Incorrect formulation. Use either classic, inverse, true power or true location")


# =========================================================================== Simpson weights
def simpson_weights(n: int, dx: float):
    """w such that scipy.integrate.simpson(y, dx=dx) == y @ w for n equally spaced samples.

    Composite Simpson for an odd sample count; for an even count scipy (>= 1.11) integrates the
    first n-1 samples with Simpson and the last interval with the three-point correction
    (-1/12, 2/3, 5/12) -- the reference calls it at beamforming.py:876.
    """
    assert n >= 2
    w = np.zeros(n)
    if n == 2:
        w[:] = 0.5
        return w * dx
    m = n if n % 2 == 1 else n - 1
    w[:m] = 2.0 / 3.0
    w[1:m:2] = 4.0 / 3.0
    w[0] = w[m - 1] = 1.0 / 3.0
    if m != n:
        w[n - 3] += -1.0 / 12.0
        w[n - 2] += 2.0 / 3.0
        w[n - 1] += 5.0 / 12.0
    return w * dx


# =========================================================================== beamformers
class BaseBeamformer:
    def __init__(self, multi_channel_signal: Signal, mic_array: MicArray, c: float = 343):
        assert isinstance(multi_channel_signal, Signal), "multi_channel_signal should be of type Signal"
        assert type(mic_array) is MicArray, "mic_array should be of type MicArray"
        assert c > 0, "Speed of sound should be bigger than 0"
        assert multi_channel_signal.number_of_channels == mic_array.number_of_points, (
            "Number of channels in the signal and number of microphones do not match")
        self.signal = multi_channel_signal
        self.mics = mic_array
        self.c = c

    def get_frequency_range_from_he(self, range_he=[4, 10]) -> list:
        assert len(range_he) == 2, "Range in He should have length two"
        return [self.mics.he_to_hz(range_he[0], c=self.c), self.mics.he_to_hz(range_he[1], c=self.c)]

    def show_info(self):
        print(f"Beamformer: {getattr(self, 'beamformer_type', 'base')}; mics: {self.mics.number_of_points}; c: {self.c}")


class BeamformerGridded(BaseBeamformer):
    def __init__(self, multi_channel_signal: Signal, mic_array: MicArray, grid: Grid,
                 steering_vector: SteeringVector, c: float = 343):
        super().__init__(multi_channel_signal, mic_array, c)
        assert type(steering_vector) is SteeringVector, "steering_vector should be of type SteeringVector"
        assert issubclass(type(grid), Grid), "grid should be a Grid object"
        self.grid = grid
        self.st_vec = steering_vector


class BeamformerDASFrequency(BeamformerGridded):
    """Delay-and-sum beamforming in the frequency domain."""

    beamformer_type = "Delay-and-sum (Frequency)"

    def get_beamformer_map(self, center_frequency_hz: float, octave_fraction: int = 3,
                           remove_csm_diagonal: bool = True):
        """Map of integrated beamformer power in the (1/octave_fraction)-octave band around
        `center_frequency_hz`; float64 array shaped by `grid.reconstruct_map_shape`."""
        self.center_frequency_hz = center_frequency_hz
        self.octave_fraction = octave_fraction
        self.f_range_hz = _get_fractional_octave_bandwidth(center_frequency_hz, octave_fraction)
        txt = "Beamformer computation has started successfully:"
        print("\n" + txt)
        print("-" * len(txt))
        print("...csm...")
        f, csm_dev = self._csm_device()
        print("...Steering vector...")
        flat = self._map_from_device_csm(f, csm_dev, remove_csm_diagonal)
        self.map = self.grid.reconstruct_map_shape(flat)
        return self.map.copy()

    # -- pieces, overridable by the distributed front-end ------------------------------------
    def _csm_device(self):
        return self.signal.get_csm_device()

    def _select_bins(self, f):
        ids = find_nearest_points_index_in_vector(self.f_range_hz, f)
        id1, id2 = int(ids[0]), int(ids[1])
        if id1 == id2:               # a single bin
            id2 += 1
        return id1, id2              # half open: the upper band-edge bin is excluded

    def _grid_rows(self):
        """Rows of the grid this process evaluates (all of them on a single GPU)."""
        return 0, self.grid.number_of_points

    def _map_from_device_csm(self, f, csm_dev, remove_csm_diagonal, per_bin: bool = False):
        id1, id2 = self._select_bins(f)
        fsel = f[id1:id2]
        nbins = id2 - id1
        self.f_range_hz = np.array([fsel[0], fsel[-1]])
        weights = simpson_weights(nbins, fsel[1] - fsel[0]) if nbins > 1 else np.ones(1)
        mode = engine.resolve_map_mode()
        op = engine.das_prepare(csm_dev, id1, nbins, remove_csm_diagonal, mode)
        print("...Apply...")
        g0, g1 = self._grid_rows()
        out, pb = engine.das_map(op, self.mics.number_of_points, fsel, weights, self.c, self.mics._coordinates,
                                 self.mics.array_center_coordinates, self.grid._device_coordinates(csm_dev.device)[g0:g1],
                                 _steering_code(self.st_vec), remove_csm_diagonal, mode, per_bin=per_bin)
        self._per_bin_device = pb
        return self._gather_map(out)

    def _gather_map(self, out_dev):
        return out_dev.cpu().numpy()


# =========================================================================== next rows
class _QuadraticFormBeamformer(BeamformerGridded):
    """Shared front-end of the beamformers whose map is a function of Re(h^H A_f h) with a per-bin
    Hermitian matrix A_f derived from the CSM: the O(G F M^2) part runs in the same map kernel as
    delay-and-sum (per-bin output, no clamping); the O(F M^3) matrix function of each bin is a
    batched dense factorisation on the device (torch.linalg) and is not the hot loop."""

    def _band(self, center_frequency_hz, octave_fraction):
        self.center_frequency_hz = center_frequency_hz
        self.octave_fraction = octave_fraction
        self.f_range_hz = _get_fractional_octave_bandwidth(center_frequency_hz, octave_fraction)
        txt = "Beamformer computation has started successfully:"
        print("\n" + txt)
        print("-" * len(txt))
        print("...csm...")
        f, csm = self.signal.get_csm_device()
        print("...Steering vector...")
        ids = find_nearest_points_index_in_vector(self.f_range_hz, f)
        id1, id2 = int(ids[0]), int(ids[1])
        if id1 == id2:
            id2 += 1
        fsel = f[id1:id2]
        self.f_range_hz = np.array([fsel[0], fsel[-1]])
        return fsel, csm[id1:id2]

    def _per_bin_power(self, fsel, matrices, mode=None):
        """(G, F) float64 CUDA tensor of Re(h^H A_f h)."""
        mode = engine.resolve_map_mode(mode)
        # The map kernel evaluates Re(h^H A h) from the lower triangle of A, i.e. it assumes A Hermitian.  The
        # reference's full quadratic form equals h^H (A + A^H)/2 h for ANY A, so symmetrising first reproduces
        # it also when a matrix function (SVD of a rank-deficient CSM, an ill-conditioned inverse) comes back
        # visibly non-Hermitian.
        matrices = 0.5 * (matrices + matrices.mH)
        op = engine.das_prepare(matrices.contiguous(), 0, len(fsel), False, mode)
        _, pb = engine.das_map(op, self.mics.number_of_points, fsel, np.ones(len(fsel)), self.c,
                               self.mics._coordinates, self.mics.array_center_coordinates,
                               self.grid._device_coordinates(matrices.device), _steering_code(self.st_vec), False,
                               mode, per_bin=True)
        return pb

    def _finish(self, per_bin, fsel):
        import torch

        if len(fsel) > 1:
            w = torch.from_numpy(simpson_weights(len(fsel), fsel[1] - fsel[0])).to(per_bin.device)
            flat = per_bin @ w
        else:
            flat = per_bin[:, 0]
        self.map = self.grid.reconstruct_map_shape(flat.cpu().numpy())
        return self.map.copy()


class BeamformerMVDR(_QuadraticFormBeamformer):
    """Minimum-variance distortionless-response (Capon) beamformer: 1 / Re(h^H C^-1 h) per bin."""

    beamformer_type = "MVDR"

    def get_beamformer_map(self, center_frequency_hz: float, octave_fraction: int = 3, gamma: float = 10):
        import torch

        fsel, csm = self._band(center_frequency_hz, octave_fraction)
        inv = torch.linalg.inv(csm)
        print("...Apply...")
        return self._finish(1.0 / self._per_bin_power(fsel, inv), fsel)


class BeamformerFunctional(_QuadraticFormBeamformer):
    """Functional beamforming (Dougherty 2014): the CSM is replaced by u s^(1/gamma) v^H and the
    beamformer output is raised to the power gamma around the steering-vector norm."""

    beamformer_type = "Functional"

    def get_beamformer_map(self, center_frequency_hz: float, octave_fraction: int = 3, gamma: float = 10):
        import torch

        fsel, csm = self._band(center_frequency_hz, octave_fraction)
        # The reference takes u s^(1/gamma) v^H from an SVD (:1201).  The singular triplets of a Hermitian matrix
        # are (|lambda|, sign(lambda) v, v), so the same matrix comes out of the eigen-decomposition -- which stays
        # Hermitian also where the SVD of a rank-deficient CSM returns unrelated u and v for the null space.
        lam, v = engine.hermitian_eig(csm)
        powered = (v * (torch.sign(lam) * lam.abs() ** (1.0 / gamma)).to(v.dtype)[:, None, :]) @ v.mH
        print("...Apply...")
        power = self._per_bin_power(fsel, powered)
        # steering normalisation h^H h = sum_m |h_m|^2 does not depend on the bin: identity matrix, one bin
        M = self.mics.number_of_points
        eye = torch.eye(M, dtype=csm.dtype, device=csm.device)[None]
        norm = self._per_bin_power(fsel[:1], eye)
        return self._finish((power / norm) ** gamma * norm, fsel)


class BeamformerOrthogonal(_QuadraticFormBeamformer):
    """Orthogonal beamforming (Sarradj 2010): per bin, each of the largest eigenpairs of the CSM is
    mapped with |h^H v|^2 = h^H (v v^H) h and its eigenvalue is assigned to the grid point where
    that map peaks.  The O(G F E M^2) maps run on the tensor-core kernel; the peak location is an
    arg-max over the grid that a rounding difference of that kernel (~1e-6 of a column's maximum)
    could move, so every grid point within 1e-3 of its column's maximum is re-evaluated in fp64
    (a few device operations on the (point, bin) pairs) and the arg-max and the assigned power are
    taken from those values.  With a map mode forced by the caller the whole map runs on that
    kernel's fp64 CUDA-core sibling instead (`simt_f64`)."""

    beamformer_type = "Orthogonal"
    _recheck_margin = 1e-3

    def get_beamformer_map(self, center_frequency_hz: float, octave_fraction: int = 3,
                           number_eigenvalues: int | None = None):
        import torch

        if number_eigenvalues is None:
            number_eigenvalues = self.signal.number_of_channels // 2
        else:
            assert number_eigenvalues <= self.signal.number_of_channels, (
                "Number of eigenvalues cannot be more than number of microphones")
            assert number_eigenvalues > 0, "At least one eigenvalue of the CSM must be regarded"
        fsel, csm = self._band(center_frequency_hz, octave_fraction)
        F, G = len(fsel), self.grid.number_of_points
        dev = csm.device
        w, v = engine.hermitian_eig(csm)                       # ascending eigenvalues
        print("...Apply...")
        out = torch.zeros((G, F), dtype=torch.float64, device=dev)
        cols = torch.arange(F, device=dev)
        tensor_cores = engine.config.map_mode.lower() == "auto" and engine.tc_map_available()
        if tensor_cores:
            k_dev = torch.from_numpy(np.ascontiguousarray(fsel * np.pi * 2 / self.c, dtype=np.float64)).to(dev)
            mic_dev = torch.from_numpy(np.ascontiguousarray(self.mics._coordinates, dtype=np.float64)).to(dev)
            ctr_dev = torch.from_numpy(np.ascontiguousarray(self.mics.array_center_coordinates, dtype=np.float64)).to(dev)
            g_dev = self.grid._device_coordinates(dev)
            form = _steering_code(self.st_vec)
        # a few eigenpairs at a time keeps the batch of rank-1 matrices small: (chunk * F, M, M)
        chunk = max(1, min(number_eigenvalues, (1 << 28) // max(1, F * csm.shape[1] ** 2 * 16)))
        for e0 in range(0, number_eigenvalues, chunk):
            es = range(e0, min(number_eigenvalues, e0 + chunk))
            vecs = torch.stack([v[:, :, -e - 1] for e in es])                         # (E, F, M)
            rank1 = vecs[..., :, None] * vecs.conj()[..., None, :]                    # (E, F, M, M)
            rank1 = rank1.reshape(-1, *rank1.shape[2:])
            power = None
            if tensor_cores:
                power = self._per_bin_power(np.tile(fsel, len(es)), rank1)            # (G, E*F), FP16x3
                top = power.amax(dim=0)
                pts, cs = (power >= top - self._recheck_margin * top.abs()).nonzero(as_tuple=True)
                if pts.numel() <= 64 * power.shape[1] + 4096:
                    h = _steering_pairs(k_dev[cs % F], g_dev[pts], mic_dev, ctr_dev, form)      # (pairs, M)
                    exact = (h.conj() * vecs.reshape(-1, vecs.shape[-1])[cs]).sum(dim=1).abs() ** 2
                    power.fill_(-float("inf"))
                    power[pts, cs] = exact
                else:
                    power = None               # a flat column (no distinct peak): whole map in fp64
            if power is None:
                power = self._per_bin_power(np.tile(fsel, len(es)), rank1, mode="simt_f64")
            power = power.reshape(G, len(es), F)
            for i, e in enumerate(es):                                                 # eigenpair order matters
                peak = torch.argmax(power[:, i, :], dim=0)                             # (F,)
                out[peak, cols] = power[peak, i, cols] * w[:, -e - 1]
        return self._finish(out, fsel)


class BeamformerCleanSC(_QuadraticFormBeamformer):
    """CLEAN-SC deconvolution (Sijtsma 2007) of the delay-and-sum map.

    The reference deconvolves one bin after the other (`_clean_sc_deconvolve`,
    beamforming/_beamforming.py:194-297); here all bins of the band iterate in lock-step: every
    iteration is one arg-max per bin, ONE launch that forms the source component of every bin
    (`b2_cleansc_component`: stopping criterion, steering vector to the peak, the 20 fixed-point
    steps, the degraded CSM) and ONE launch that subtracts Re(h^H G_f h) of the rank-one G_f from
    every bin's dirty map in O(G F M) (`b2_das_rank1_update`).  The peak search makes the
    iteration sensitive to rounding, so everything stays in fp64 (the first, full-rank map runs on
    the fp64 CUDA-core kernel)."""

    beamformer_type = "CleanSC"

    def get_beamformer_map(self, center_frequency_hz: float, octave_fraction: int = 3,
                           maximum_iterations: int | None = None, safety_factor: float = 0.5,
                           remove_csm_diagonal: bool = False):
        import torch

        if maximum_iterations is None:
            maximum_iterations = self.signal.number_of_channels * 2
        else:
            assert maximum_iterations > 0, "Number of iterations must be positive"
        assert safety_factor > 0 and safety_factor <= 1, (
            f"{safety_factor} is not valid. The safety factor (loop gain) should be in ]0, 1]")
        fsel, csm = self._band(center_frequency_hz, octave_fraction)
        F, M = len(fsel), self.mics.number_of_points
        dev = csm.device
        csm = csm.clone()
        diag = torch.arange(M, device=dev)
        if remove_csm_diagonal:
            csm[:, diag, diag] = 0
        print("...Create and deconvolve map...")
        dirty = self._per_bin_power(fsel, csm, mode="simt_f64").clone().contiguous()          # (G, F)
        second = torch.zeros_like(dirty)
        D0, D1 = (2 * csm).contiguous(), csm.clone().contiguous()
        active = torch.ones(F, dtype=torch.bool, device=dev)
        cols = torch.arange(F, device=dev)
        zero = torch.zeros((), dtype=torch.float64, device=dev)
        k_dev = torch.from_numpy(np.ascontiguousarray(fsel * np.pi * 2 / self.c, dtype=np.float64)).to(dev)
        f_dev = torch.from_numpy(np.ascontiguousarray(fsel, dtype=np.float64)).to(dev)
        mic_dev = torch.from_numpy(np.ascontiguousarray(self.mics._coordinates, dtype=np.float64)).to(dev)
        ctr = np.ascontiguousarray(self.mics.array_center_coordinates, dtype=np.float64)
        g_dev = self.grid._device_coordinates(dev)
        form = _steering_code(self.st_vec)
        h_src = torch.empty((F, M), dtype=torch.complex128, device=dev)
        scale = torch.empty(F, dtype=torch.float64, device=dev)
        for it in range(maximum_iterations):
            peak = torch.argmax(dirty, dim=0)                                     # (F,)
            pmax = dirty[peak, cols]
            # the reference stores the peak BEFORE it looks at the stopping criterion (:250-260)
            second.index_put_((peak, cols), torch.where(active, pmax * safety_factor, zero), accumulate=True)
            # stopping criterion, steering vector to the peak, source component h, degraded CSM: one launch for all bins
            engine.cleansc_component(D0, D1, peak, pmax, active, k_dev, mic_dev, ctr, g_dev, form, safety_factor,
                                     remove_csm_diagonal, h_src, scale)
            # the only host round trip of the loop, every 8th iteration: finished bins are masked out (scale 0), so
            # running a few iterations past the last active bin changes nothing
            if it % 8 == 7 and not bool(active.any()):
                break
            # dirty -= Re(h_g^H G h_g) * safety_factor with the rank-one G = h h^H P_max of every bin
            engine.das_rank1_update(h_src, scale, f_dev, self.c, mic_dev, ctr, g_dev, form, remove_csm_diagonal,
                                    -safety_factor, dirty)
        return self._finish(second, fsel)


def _steering_pairs(k_dev, points, mic_dev, center_dev, steering: int):
    """Steering vectors h[f, m] of bin f towards ITS OWN point `points[f]` (reference beamforming.py:1515-1703,
    distances by the expansion formula of helpers/other.py:131-153) as a few fp64 device operations -- what CLEAN-SC
    needs per iteration (the dense (F, M, G) kernel would compute F times too much)."""
    import torch

    def dist(a, b):                    # (F, 3) x (M, 3) -> (F, M)
        return torch.sqrt((a * a).sum(1)[:, None] + (b * b).sum(1)[None, :] - 2.0 * a @ b.T)

    M = mic_dev.shape[0]
    rti = dist(points, mic_dev)
    rt0 = dist(points, center_dev[None, :])[:, 0]
    phase = torch.exp(-1j * k_dev[:, None] * (rti - rt0[:, None]))
    if steering == N.STEER_CLASSIC:
        return phase / M
    if steering == N.STEER_INVERSE:
        return rti / M / rt0[:, None] * phase
    rtj = (1.0 / rti ** 2).sum(dim=1)
    if steering == N.STEER_TRUE_POWER:
        return 1.0 / rt0[:, None] / rti / rtj[:, None] * phase
    return 1.0 / rti / torch.sqrt(M * rtj[:, None]) * phase


# =========================================================================== source simulation (next row f2)
def _source_filters(source, mics: "MicArray", c: float, length: int):
    """Taps (M, 31) with the 1 / (1 + r) gain folded in and integer shifts (M,) of one source -- the filter
    `fractional_delay(emitted, r / c, keep_length=True)` applies per microphone (reference :1443-1451)."""
    from ..standard._standard_backend import fractional_delay_filters

    distances = np.atleast_1d(mics.get_distances_to_point(source.coordinates)).astype(np.float64)
    delays = distances / c
    assert np.all(delays >= 0), "Delay must be positive"
    delay_samples = delays * source.emitted_signal.sampling_rate_hz
    assert np.all(delay_samples < length), "Delay too large for the given signal"
    shifts, taps = fractional_delay_filters(delay_samples, 30, 60.0)
    on_mic = delays == 0                                  # zero delay: the reference returns an unfiltered copy
    taps[on_mic] = 0.0
    taps[on_mic, 0] = 1.0
    shifts[on_mic] = 0
    return taps / (1.0 + distances)[:, None], shifts


class MonopoleSource:
    """Omnidirectional source: one emitted one-channel `Signal` at a position (reference :1397-1425)."""

    def __init__(self, signal: Signal, coordinates):
        assert signal.number_of_channels == 1, "Only signals with a single channel are supported"
        coordinates = np.squeeze(coordinates)
        assert len(coordinates) == 3 and coordinates.ndim == 1, "Coordinates should have exactly three values"
        self.emitted_signal = signal
        self.coordinates = coordinates

    def get_signals_on_array(self, mics: MicArray, c: float = 343) -> Signal:
        """The emitted signal as every microphone receives it: delayed by distance / c with a 31-tap windowed-sinc
        fractional-delay filter, divided by 1 + distance, same length as the emitted signal (reference :1427-1458).
        The returned `Signal` holds a CUDA tensor (time samples, microphones) -- ready for `get_csm`."""
        x = self.emitted_signal.time_data
        length = int(x.shape[0])
        taps, shifts = _source_filters(self, mics, c, length)
        out = engine.delay_mix([x[:, 0]], taps[None], shifts[None], length)
        return Signal(None, out, self.emitted_signal.sampling_rate_hz)


def mix_sources_on_array(sources, mics: MicArray, c: float = 343) -> Signal:
    """Sum of the sources' signals on the array (reference :1461-1511), one kernel launch for all sources.  When the
    emitted signals differ in length the reference trims the running sum and the next emitted signal to the
    shorter of the two, with a warning; the same lengths apply here (the emitted `Signal`s are not modified, and
    the caller's list is left as it is -- the reference pops its first element)."""
    from warnings import warn

    if type(sources) is MonopoleSource:
        sources = [sources]
    assert len(sources) > 0, "There must be at least one source to project on array"
    assert all(type(i) is MonopoleSource for i in sources), "All sources in list should be of type Source"
    fs = sources[0].emitted_signal.sampling_rate_hz
    total = int(sources[0].emitted_signal.time_data.shape[0])
    signals, taps, shifts = [], [], []
    for n, s in enumerate(sources):
        x = s.emitted_signal.time_data
        if total != int(x.shape[0]):
            warn("Emitted signals from sources differ in length. Trimming to shortest will be done")
            total = min(total, int(x.shape[0]))
        t, d = _source_filters(s, mics, c, total)          # this source is delayed at the running length
        signals.append(x[:total, 0])
        taps.append(t)
        shifts.append(d)
    out = engine.delay_mix(signals, np.stack(taps), np.stack(shifts), total)
    return Signal(None, out, fs)
