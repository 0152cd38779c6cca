"""One-process-per-GPU sharding of the two stages (torch.distributed; NCCL on the B200 box,
gloo in the CPU tests).

Stage 1 -- Welch frames are sharded: rank r owns the contiguous frame range
`frame_shard(K, r, R)`, reads only the samples those frames touch (a halo of wl - step samples
overlaps the neighbour), accumulates the UN-normalised sums S_f^(r) as a PACKED lower triangle, and an
all-reduce(sum) over NVLink -- issued per bin range so that it overlaps the rank-K update of the next range --
gives every rank the full sum.  The non-linear finalisation (mean, scaling,
element-wise sqrt, mirroring) runs after the reduction on every rank, so all ranks hold the
full CSM for stage 2.
Stage 2 -- grid points are sharded (`grid_shard`); the only communication is the final
all-gather of G floats.

The reference has no distributed path; results are identical to the
single-GPU path up to the summation order of the frame sums.
"""

from __future__ import annotations

import numpy as np

from .beamforming.beamforming import BeamformerDASFrequency
from .classes.signal import Signal
from .standard import _spectral_methods as _sm


# --------------------------------------------------------------------------- index math
def split_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [begin, end) share of range(n) for `rank` (first n % world ranks get one more)."""
    assert 0 <= rank < world
    base, extra = divmod(n, world)
    begin = rank * base + min(rank, extra)
    return begin, begin + base + (1 if rank < extra else 0)


def frame_shard(n_frames: int, rank: int, world: int) -> tuple[int, int]:
    return split_range(n_frames, rank, world)


def grid_shard(n_points: int, rank: int, world: int) -> tuple[int, int]:
    return split_range(n_points, rank, world)


def shard_sample_range(frames: tuple[int, int], step: int, wl: int, length: int) -> tuple[int, int]:
    """Samples [t0, t1) of the recording that the frames [k0, k1) read; the tail beyond
    `length` is the reference's zero padding (helpers/other.py:207-209) and is not stored."""
    k0, k1 = frames
    if k1 <= k0:
        return 0, 0
    return k0 * step, min(length, (k1 - 1) * step + wl)


# --------------------------------------------------------------------------- collectives
def _dist():
    import torch.distributed as dist

    return dist


def world() -> tuple[int, int]:
    dist = _dist()
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def allreduce_sum_(t):
    """In-place sum over ranks of a real or complex tensor (complex goes as interleaved reals)."""
    h = allreduce_sum_async_(t)
    if h is not None:
        h.wait()
    return t


def allreduce_sum_async_(t):
    """Start the in-place sum over ranks of `t` and return the work handle (None on a single rank).

    This is the `reduce_fn` of `engine.csm_partial_sums`: it is called once per bin range of the PACKED
    frame sums (the M(M+1)/2 valid entries per bin -- 1.08 GB instead of the 2.15 GB square at M = 256,
    nb = 2049), right after the rank-K update of that range was enqueued.  NCCL orders the collective after
    the work already on the current stream and runs it on its own stream, so the reduction of range i
    overlaps the update of range i+1; `handle.wait()` makes the current stream wait for the result."""
    import torch

    dist = _dist()
    if world()[1] == 1:
        return None
    return dist.all_reduce(torch.view_as_real(t) if t.is_complex() else t, op=dist.ReduceOp.SUM, async_op=True)


def allgather_rows(local, counts: list[int]):
    """Concatenate per-rank 1-D tensors of (possibly unequal) lengths `counts` on every rank."""
    import torch

    dist = _dist()
    if world()[1] == 1:
        return local
    nmax = max(counts)
    padded = torch.zeros(nmax, dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    parts = [torch.empty_like(padded) for _ in counts]
    dist.all_gather(parts, padded)
    return torch.cat([p[:c] for p, c in zip(parts, counts)])


# --------------------------------------------------------------------------- front-ends
class ShardedSignal(Signal):
    """A `Signal` of which this rank only holds the rows [row_offset, row_offset + len(rows)) -- the samples
    its Welch frames touch (`frame_shard` + `shard_sample_range`) -- of a recording of `total_length` samples.
    `rows` is a (samples, channels) numpy array (pinned or pageable) or torch tensor (host or device).
    `length_samples` reports the whole recording, so the frame arithmetic is that of the unsharded signal;
    `get_csm*` must be driven through `get_csm_device_sharded` / `DistributedBeamformerDASFrequency`."""

    def __init__(self, rows, row_offset: int, total_length: int, sampling_rate_hz: int, activate_cache: bool = False):
        assert len(rows.shape) == 2, "rows must have the shape (time samples, channels)"
        assert 0 <= row_offset and row_offset + rows.shape[0] <= total_length, "rows outside the recording"
        self.constrain_amplitude = False
        self.activate_cache = activate_cache
        self.sampling_rate_hz = sampling_rate_hz
        self._time_data = rows
        self.time_data_imaginary = None
        self._amplitude_scale_factor = 1.0
        self._row_offset, self._total_length = int(row_offset), int(total_length)
        self._invalidate()
        self.set_spectrum_parameters()
        self.set_spectrogram_parameters()

    @property
    def number_of_channels(self) -> int:
        return int(self._time_data.shape[1])

    @property
    def length_samples(self) -> int:
        return self._total_length

    def _compute_csm_device(self, frame_range=None, reduce_fn=None):
        p = self._spectrum_parameters
        rank, nranks = world()
        if frame_range is None:
            wl = p["window_length_samples"]
            step = wl - int(p["overlap_percent"] / 100 * wl)
            k_total = -(-self._total_length // step)
            frame_range, reduce_fn = frame_shard(k_total, rank, nranks), allreduce_sum_async_
        return _sm.csm_welch_device(self._time_data, self.sampling_rate_hz, p["window_length_samples"],
                                    p["window_type"], p["overlap_percent"], p["detrend"], p["average"], p["scaling"],
                                    frame_range=frame_range, reduce_fn=reduce_fn, total_length=self._total_length,
                                    row_offset=self._row_offset)


def get_csm_device_sharded(signal):
    """Frame-sharded `Signal.get_csm_device()`: every rank returns the full (f, csm)."""
    from . import engine

    rank, nranks = world()
    p = signal._spectrum_parameters
    wl = p["window_length_samples"]
    step = wl - int(p["overlap_percent"] / 100 * wl)
    k_total, _ = engine.welch_frames(signal.length_samples, wl, step)
    return signal._compute_csm_device(frame_range=frame_shard(k_total, rank, nranks), reduce_fn=allreduce_sum_async_)


class DistributedBeamformerDASFrequency(BeamformerDASFrequency):
    """`BeamformerDASFrequency` whose CSM is frame-sharded and whose grid is point-sharded over the
    ranks of the default process group.  Every rank returns the complete map."""

    def _csm_device(self):
        return get_csm_device_sharded(self.signal)

    def _grid_rows(self):
        rank, nranks = world()
        return grid_shard(self.grid.number_of_points, rank, nranks)

    def _gather_map(self, out_dev):
        _, nranks = world()
        G = self.grid.number_of_points
        counts = [b - a for a, b in (grid_shard(G, r, nranks) for r in range(nranks))]
        return allgather_rows(out_dev, counts).cpu().numpy()


def simulate_frame_sharding(n_frames: int, world_size: int) -> list[tuple[int, int]]:
    """All shards of a world (host-side helper for tests and for single-GPU emulation of R ranks)."""
    shards = [frame_shard(n_frames, r, world_size) for r in range(world_size)]
    assert shards[0][0] == 0 and shards[-1][1] == n_frames
    assert all(a[1] == b[0] for a, b in zip(shards[:-1], shards[1:]))
    return shards


__all__ = ["split_range", "frame_shard", "grid_shard", "shard_sample_range", "allreduce_sum_", "allreduce_sum_async_",
           "allgather_rows", "get_csm_device_sharded", "ShardedSignal", "DistributedBeamformerDASFrequency",
           "simulate_frame_sharding", "world"]
