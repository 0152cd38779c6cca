"""Point containers behind grids and microphone arrays (reference beamforming/_beamforming.py:14-191).

Host-side data model only: coordinates live in a float (N, 3) array; the heavy geometry (all
grid-to-microphone distances) is recomputed inside the map kernel and never materialised.
"""

from __future__ import annotations

import numpy as np

from ..helpers.other import _euclidean_distance_matrix

_AXES = ("x", "y", "z")


class BasePoints:
    """N points given as a dict with 'x', 'y' and 'z' array-likes of equal shape."""

    def __init__(self, positions: dict):
        for axis in _AXES:
            assert axis in positions, f"{axis} values are missing"
        cols = [np.asarray(positions[a]).squeeze()[None, ...] for a in _AXES]
        assert cols[0].shape == cols[1].shape == cols[2].shape, "Shapes of x, y or z are not compatible"
        self.coordinates = np.concatenate(cols, axis=0).T

    @property
    def number_of_points(self) -> int:
        return self._coordinates.shape[0]

    @property
    def coordinates(self):
        return self._coordinates.copy()      # a copy on every access, like the reference (:59-61)

    @coordinates.setter
    def coordinates(self, new_r):
        assert type(new_r) is np.ndarray, "R vectors array should be of type numpy.ndarray"
        spans = [a for i, a in enumerate(_AXES[: new_r.shape[1]]) if len(np.unique(new_r[:, i])) > 1]
        self.dim = spans
        self.ndim = len(spans)
        self._coordinates = new_r
        self._coordinates_dev = None

    def _device_coordinates(self, device=None):
        """float64 (N, 3) CUDA copy of the coordinates, uploaded once and kept until they are reassigned
        (the map kernel reads the grid from device memory on every call; 24 MB at 10^6 points)."""
        import torch

        device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
        cached = getattr(self, "_coordinates_dev", None)
        if cached is None or cached.device != device:
            cached = torch.from_numpy(np.ascontiguousarray(self._coordinates, dtype=np.float64)).to(device)
            self._coordinates_dev = cached
        return cached

    @property
    def extent(self) -> dict:
        c = self._coordinates
        return {a: [np.min(c[:, i]), np.max(c[:, i])] for i, a in enumerate(_AXES)}

    def get_distances_to_point(self, point):
        """Euclidean distances (points, new_points), squeezed like the reference (:91-114)."""
        point = np.asarray(point)
        if point.ndim == 1:
            point = point[None, ...]
        assert point.shape[1] == self._coordinates.shape[1], (
            f"Invalid shapes: {point.shape}, {self._coordinates.shape}")
        return _euclidean_distance_matrix(self._coordinates, point).squeeze()

    def find_nearest_point(self, point):
        point = np.asarray(point).squeeze()
        assert point.ndim == 1, "Passed vector is not broadcastable to a 1D-array"
        assert len(point) == 3, "Point must have exactly 3 dimensions (x, y, z)"
        index = int(np.argmin(self.get_distances_to_point(point)))
        return index, self.coordinates[index, :]

    def plot_points(self, projection=None):
        raise NotImplementedError("plotting is outside the accelerated path")
