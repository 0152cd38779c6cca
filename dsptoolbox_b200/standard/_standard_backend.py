"""Fractional-delay filter design for the source simulator.

`_fractional_delay_filter` follows reference standard/_standard_backend.py:430-493 (windowed sinc, Kaiser window of
:290-321 with the beta rule of :257-287).  `fractional_delay_filters` designs the filters of MANY delays at once
(one per microphone) as arrays; the convolution itself runs on the device (csrc/delay_fir.cu)."""

from __future__ import annotations

import numpy as np
from scipy.special import i0


def _kaiser_beta(side_lobe_suppression_db: float) -> float:
    a = abs(side_lobe_suppression_db)
    if a > 50:
        return 0.1102 * (a - 8.7)
    if a >= 21:
        return 0.5842 * (a - 21) ** 0.4 + 0.07886 * (a - 21)
    return 0.0


def fractional_delay_filters(delay_samples, filter_order: int = 30, side_lobe_suppression_db: float = 60.0):
    """Filters for an array of delays (in samples, >= 0) -> (integer_delay int64 [n], taps float64 [n, order + 1]).

    Row i delays by the fractional part of `delay_samples[i]` (plus the centring of the sinc); `integer_delay[i]`
    is the remaining whole-sample shift, which is negative for delays shorter than about half the order."""
    d = np.atleast_1d(np.asarray(delay_samples, dtype=np.float64))
    whole = np.trunc(d)
    frac = d - whole
    taps_idx = np.arange(filter_order + 1, dtype=np.float64)[None, :]
    if filter_order % 2:
        centre = np.trunc(frac) - (filter_order - 1) / 2
    else:
        centre = np.round(frac) - filter_order / 2
    sinc = np.sinc(taps_idx + centre[:, None] - frac[:, None])
    # Kaiser window shifted by the fractional delay; outside its support the argument of I0 turns imaginary
    # (I0(i y) = J0(y), real), which `i0` of a complex square root cannot give -- handle the sign explicitly
    beta = _kaiser_beta(side_lobe_suppression_db)
    alpha = filter_order / 2
    pos = taps_idx - frac[:, None]
    if filter_order % 2:
        pos = pos + 0.5
    else:
        pos = pos + (frac[:, None] > 0.5)
    arg = 1.0 - ((pos - alpha) / alpha) ** 2
    from scipy.special import j0

    inside = arg >= 0
    root = beta * np.sqrt(np.abs(arg))
    kaiser = np.where(inside, i0(root), j0(root)) / i0(beta)
    return (whole + centre).astype(np.int64), sinc * kaiser


def _fractional_delay_filter(delay_samples: float, filter_order: int, side_lobe_suppression_db: float):
    """Drop-in for reference _standard_backend.py:430-493: (integer_delay, taps) of one delay."""
    shift, taps = fractional_delay_filters([delay_samples], filter_order, side_lobe_suppression_db)
    return int(shift[0]), taps[0]
