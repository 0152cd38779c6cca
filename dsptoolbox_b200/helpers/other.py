"""Small host-side index/geometry helpers with the reference's names (helpers/other.py).

These run on the host because they are O(1)..O(M^2) parameter logic, not the hot path:
  find_nearest_points_index_in_vector  reference :9-31
  _euclidean_distance_matrix           reference :131-153
  _get_fractional_octave_bandwidth     reference :156-178
  _compute_number_frames               reference :181-213
  _pad_trim                            reference :216-259
"""

from __future__ import annotations

import numpy as np


def find_nearest_points_index_in_vector(points, vector):
    """Index of the element of `vector` closest to each of `points` (first one on ties)."""
    pts = np.atleast_1d(np.asarray(points))
    vec = np.asarray(vector)
    return np.array([int(np.argmin(np.abs(p - vec))) for p in pts], dtype=np.int_)


def _euclidean_distance_matrix(x, y):
    """(points of x) x (points of y) distances via |x|^2 + |y|^2 - 2 x.y, like the reference."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    assert x.ndim == 2 and y.ndim == 2, "Inputs must have exactly two dimensions"
    assert x.shape[1] == y.shape[1], "Dimensions do not match"
    sq = (x * x).sum(axis=1)[:, None] + (y * y).sum(axis=1)[None, :] - 2.0 * (x @ y.T)
    return np.sqrt(sq)


def _get_fractional_octave_bandwidth(f_c: float, fraction=1):
    """Band edges f_c * 2^(-+1/(2*fraction)); fraction 0 collapses the band to f_c."""
    if fraction == 0:
        return np.array([f_c, f_c])
    half = 1 / fraction / 2
    return np.array([f_c * 2 ** (-half), f_c * 2 ** half])


def _compute_number_frames(window_length: int, step: int, signal_length: int, zero_padding: bool):
    """(frames, padding samples) of the framed representation."""
    if zero_padding:
        return int(np.ceil(signal_length / step)), window_length - int(signal_length % step)
    return int(np.ceil((signal_length - window_length) / step)), 0


def _pad_trim(vector, desired_length: int, axis: int = 0, in_the_end: bool = True):
    """Zero-pad or cut `vector` along `axis` to `desired_length` (1-D or 2-D arrays)."""
    v = np.asarray(vector)
    one_d = v.ndim < 2
    if one_d:
        assert axis == 0, "You can only pad along the 0 axis"
        v = v[:, None]
    elif v.ndim > 2:
        v = v.squeeze()
        if v.ndim > 2:
            raise ValueError("This function is only implemented for 1D and 2D arrays")
    v = np.moveaxis(v, axis, 0)
    n = v.shape[0]
    if desired_length > n:
        fill = np.zeros((desired_length - n,) + v.shape[1:], dtype=v.dtype)
        out = np.concatenate([v, fill] if in_the_end else [fill, v])
    elif in_the_end:
        out = v[:desired_length].copy()
    else:
        out = v[n - desired_length:].copy()
    out = np.moveaxis(out, 0, axis)
    return out[:, 0] if one_d else out
