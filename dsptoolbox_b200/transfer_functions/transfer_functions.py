"""`compute_transfer_function` of reference transfer_functions/transfer_functions.py:419-536.

The reference calls `_welch` four to five times per output channel (G_xx, G_yy, G_xy and, for
H2, G_yx).  All of those are entries of the Welch cross-spectral matrix of the stacked channels
[input | output] -- `_welch(a, b)` is element [b, a] and `_welch(a, None)` the (real) diagonal --
so ONE pass of the stage-1 kernels (framing + window + rFFT, Hermitian rank-K update, finalise)
yields every spectrum at once; only the few needed columns are copied to the host, where the
reference's element-wise formulas are applied in float64.
"""

from __future__ import annotations

import numpy as np

from ..classes.signal import Signal
from ..classes.spectrum import Spectrum
from ..standard import _spectral_methods as _sm
from .enums import TransferFunctionType


def compute_transfer_function(output: Signal, input: Signal, window_length_samples: int,
                              mode: TransferFunctionType = TransferFunctionType.H2) -> Spectrum:
    """Transfer function H1, H2 or H3 between `input` and every channel of `output`, with the coherence saved
    as `coherence` on the returned `Spectrum`.  A one-channel input is the input of all output channels.  The
    spectrum parameters of `input` are used (window, overlap, detrend, average, scaling), like the reference."""
    assert input.sampling_rate_hz == output.sampling_rate_hz, "Sampling rates do not match"
    assert input.time_data.shape[0] == output.time_data.shape[0], "Signal lengths do not match"
    if input.number_of_channels != 1:
        assert input.number_of_channels == output.number_of_channels, "Channel number does not match between signals"
        multichannel = False
    else:
        multichannel = True
    if not isinstance(mode, TransferFunctionType):
        raise ValueError("Unsupported transfer function type")
    p = input._spectrum_parameters
    n_in, n_out = input.number_of_channels, output.number_of_channels
    a, b = input.time_data, output.time_data
    if hasattr(a, "is_cuda") or hasattr(b, "is_cuda"):            # torch tensors (pinned host or device) stay torch
        import torch

        a, b = torch.as_tensor(a), torch.as_tensor(b)
        stacked = torch.cat([a, b.to(a.device)], dim=1)
    else:
        stacked = np.concatenate([np.asarray(a), np.asarray(b)], axis=1)
    f, csm = _sm.csm_welch_device(stacked, input.sampling_rate_hz, window_length_samples, p["window_type"],
                                  p["overlap_percent"], p["detrend"], p["average"], p["scaling"])
    import torch

    xi = torch.zeros(n_out, dtype=torch.long, device=csm.device) if multichannel else torch.arange(n_out, device=csm.device)
    yi = n_in + torch.arange(n_out, device=csm.device)
    g_xx = csm[:, xi, xi].real.cpu().numpy()          # _welch(x, None)
    g_yy = csm[:, yi, yi].real.cpu().numpy()          # _welch(y, None)
    g_xy = csm[:, yi, xi].cpu().numpy()               # _welch(x, y) = mean conj(X) Y
    # With detrend the Welch kernels write the DC bin of every spectrum as an exact 0 (the DFT of a de-meaned frame;
    # the reference carries rounding noise of ~1e-17 there and divides noise by noise).  0 / 0 is defined as 0 here:
    # transfer function and coherence of a bin without any energy.
    with np.errstate(divide="ignore", invalid="ignore"):
        if mode == TransferFunctionType.H1:
            tf = g_xy / g_xx
        elif mode == TransferFunctionType.H2:
            g_yx = csm[:, xi, yi].cpu().numpy()       # _welch(y, x)
            tf = g_yy / g_yx
        else:
            tf = g_xy / np.abs(g_xy) * (g_yy / g_xx) ** 0.5
        coherence = np.abs(g_xy) ** 2 / g_xx / g_yy
    dead = (g_xx == 0) | (g_yy == 0)
    tf = np.where(dead, 0.0, tf)
    coherence = np.where(dead, 0.0, coherence)
    spec = Spectrum(np.fft.rfftfreq(window_length_samples, 1 / input.sampling_rate_hz), tf.astype(np.complex128))
    spec.set_coherence(coherence)
    return spec
