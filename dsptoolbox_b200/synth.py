"""Seeded synthetic array recordings and geometries for tests and benchmarks.

Not part of the accelerated path: it only manufactures inputs of the shapes named in
BASELINE.json.  A monopole at `source_xyz` radiating white noise is
recorded by every microphone with the exact (circular) propagation delay d/c applied in the
frequency domain and the reference simulator's amplitude law 1/(1+d)
(reference beamforming/beamforming.py:1451), plus independent sensor noise.
"""

from __future__ import annotations

import numpy as np

DEFAULT_SOURCE = (0.1, -0.2, 1.0)


def disc_array(n_mics: int, radius: float, seed: int = 7) -> dict:
    """n_mics positions uniformly distributed on a disc of `radius` in the plane z = 0."""
    rng = np.random.default_rng(seed)
    r = radius * np.sqrt(rng.uniform(0.0, 1.0, n_mics))
    phi = rng.uniform(0.0, 2 * np.pi, n_mics)
    return dict(x=r * np.cos(phi), y=r * np.sin(phi), z=np.zeros(n_mics))


def positions_to_xyz(positions: dict) -> np.ndarray:
    return np.stack([np.asarray(positions[k], dtype=np.float64) for k in ("x", "y", "z")], axis=1)


def point_source_recording(
    mic_xyz: np.ndarray,
    length_samples: int,
    sampling_rate_hz: int = 48_000,
    source_xyz=DEFAULT_SOURCE,
    noise_sigma: float = 0.1,
    c: float = 343.0,
    seed: int = 1,
    dtype=np.float64,
    channel_block: int = 16,
) -> np.ndarray:
    """(length_samples, n_mics) recording of one white-noise monopole plus sensor noise."""
    mic_xyz = np.asarray(mic_xyz, dtype=np.float64)
    M = mic_xyz.shape[0]
    rng = np.random.default_rng(seed)
    s = rng.standard_normal(length_samples)
    S = np.fft.rfft(s)
    f = np.fft.rfftfreq(length_samples, 1.0 / sampling_rate_hz)
    d = np.sqrt(np.sum((mic_xyz - np.asarray(source_xyz, dtype=np.float64)[None, :]) ** 2, axis=1))
    out = np.empty((length_samples, M), dtype=dtype)
    for m0 in range(0, M, channel_block):
        dm = d[m0:m0 + channel_block]
        ph = np.exp(-2j * np.pi * f[:, None] * (dm[None, :] / c))
        sig = np.fft.irfft(S[:, None] * ph, n=length_samples, axis=0) / (1.0 + dm[None, :])
        sig += noise_sigma * rng.standard_normal(sig.shape)
        out[:, m0:m0 + channel_block] = sig
    return out


def point_source_recording_torch(
    mic_xyz: np.ndarray,
    length_samples: int,
    sampling_rate_hz: int = 48_000,
    source_xyz=DEFAULT_SOURCE,
    noise_sigma: float = 0.1,
    c: float = 343.0,
    seed: int = 1,
    device="cuda",
    dtype=None,
    channel_block: int = 8,
):
    """Same signal model generated on `device` with torch (for the multi-GB benchmark inputs,
    where numpy on the host would take minutes).  Returns a (length_samples, n_mics) tensor.
    The random streams differ from the numpy generator; the physics is identical."""
    import torch

    dtype = dtype or torch.float64
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    mic = torch.as_tensor(np.asarray(mic_xyz, dtype=np.float64), device=device)
    M = mic.shape[0]
    src = torch.as_tensor(np.asarray(source_xyz, dtype=np.float64), device=device)
    d = torch.sqrt(torch.sum((mic - src[None, :]) ** 2, dim=1))
    s = torch.randn(length_samples, generator=g, device=device, dtype=torch.float64)
    S = torch.fft.rfft(s)
    f = torch.fft.rfftfreq(length_samples, 1.0 / sampling_rate_hz, device=device, dtype=torch.float64)
    out = torch.empty((length_samples, M), device=device, dtype=dtype)
    for m0 in range(0, M, channel_block):
        dm = d[m0:m0 + channel_block]
        ph = torch.exp(-2j * np.pi * f[:, None] * (dm[None, :] / c))
        sig = torch.fft.irfft(S[:, None] * ph, n=length_samples, dim=0) / (1.0 + dm[None, :])
        sig += noise_sigma * torch.randn(sig.shape, generator=g, device=device, dtype=torch.float64)
        out[:, m0:m0 + channel_block] = sig.to(dtype)
        del ph, sig
    return out


# Named workloads of BASELINE.json `configs`.
def workload(name: str) -> dict:
    """Parameters of cfg1..cfg5 (geometry is built by the caller; cfg1 needs array.xml)."""
    table = {
        "cfg1": dict(mics=64, array="array.xml", seconds=1, wl=1024, grid=("2d", 50, 50), fc=4000.0, frac=3),
        "cfg2": dict(mics=64, array=("disc", 0.75), seconds=10, wl=4096, grid=None, fc=None, frac=None),
        "cfg3": dict(mics=64, array=("disc", 0.75), seconds=10, wl=4096, grid=("3d", 200, 200, 20), fc=2000.0, frac=0.25),
        "cfg4": dict(mics=256, array=("disc", 1.0), seconds=60, wl=4096, grid=None, fc=None, frac=None),
        "cfg5": dict(mics=256, array=("disc", 1.0), seconds=60, wl=4096, grid=("3d", 100, 100, 100), fc=219.089, frac=0.0738),
    }
    return dict(table[name], name=name, fs=48_000)


def workload_grid_lines(cfg: dict):
    """Lines of the regular grid of a workload (x, y in [-1, 1] or [-.5, .5]; z in [0.5, 1.5])."""
    g = cfg["grid"]
    if g is None:
        return None
    if g[0] == "2d":
        return (np.linspace(-0.5, 0.5, g[1]), np.linspace(-0.5, 0.5, g[2]))
    return (np.linspace(-1.0, 1.0, g[1]), np.linspace(-1.0, 1.0, g[2]), np.linspace(0.5, 1.5, g[3]))
