"""Device-side orchestration of the two stages (host logic only; all arithmetic is in libb2dsp).

PyTorch is used for what it is good at here: owning device memory, streams and (in
dist.py) the NCCL process group.  Nothing in this file computes on the CPU.

Stage 1 (Welch CSM), replacing standard/_spectral_methods.py:285-371 of the reference:
    K0 b2_prepare_channels -> K1 b2_welch_rfft -> K2 b2_csm_herk -> [all-reduce] -> K2e b2_csm_finalize
Stage 2 (DAS map), replacing beamforming/beamforming.py:841-878:
    K3a b2_das_prepare -> K3 b2_das_map
"""

from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _native as N


# --------------------------------------------------------------------------- configuration
@dataclass
class Config:
    """Run-time switches of the accelerated path (the reference has none: SURVEY.md section 5)."""

    #: "fp64" = strict (fp64 FFT + fp64 tensor-core accumulation, meets rtol 1e-5 on any data),
    #: "fp32" = fast (fp32 FFT and spectra, fp64 accumulation).
    csm_precision: str = "fp64"
    #: "auto" (= "f16x3": two FP16 halves per operand, three tensor products), "mixed" (TF32 product + BF16
    #: corrections), "tf32x3", "tf32" (single pass, outside the parity contract), "simt_f32", "simt_f64"
    map_mode: str = "auto"
    #: device spectra buffer budget in bytes; longer recordings are processed in frame chunks
    spectra_budget_bytes: int = 24 << 30
    #: multi-GPU: the packed frame sums are all-reduced in this many bin ranges, each reduction
    #: overlapping the rank-K update of the next range (1 = one blocking all-reduce at the end)
    reduce_ranges: int = 8
    #: host recordings above this size are uploaded in frame chunks on a copy stream, so that the
    #: upload of chunk i+1 overlaps K0/K1/K2 of chunk i (0 disables the chunking)
    h2d_chunk_bytes: int = 768 << 20
    #: pageable host recordings above this size go through a ring of pinned staging buffers filled by
    #: `h2d_stage_threads` host threads (0 leaves the upload to the driver's own staging)
    h2d_stage_min_bytes: int = 64 << 20
    h2d_stage_threads: int = 6
    #: last-call statistics (kernel times are only filled when `profile` is on)
    profile: bool = False
    stats: dict = field(default_factory=dict)


config = Config()

#: number of libb2dsp kernel launches issued by this process (bench.py reports it as gpu_launches)
launches = 0


def _count(n: int = 1):
    global launches
    launches += n

_PREC = {"fp64": N.F64, "fp32": N.F32}
_MAP_MODES = {"simt_f64": N.MAP_SIMT_F64, "simt_f32": N.MAP_SIMT_F32, "tf32": N.MAP_TC_TF32,
              "tf32x3": N.MAP_TC_TF32X3, "mixed": N.MAP_TC_MIXED, "f16x3": N.MAP_TC_F16X3}


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise RuntimeError("dsptoolbox_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def tc_map_available() -> bool:
    """True when the tcgen05 map kernel is compiled into the library."""
    return N.lib().b2_das_operand_bytes(64, 1, N.MAP_TC_TF32X3) > 0


def resolve_map_mode(mode: str | None = None) -> int:
    mode = (mode or config.map_mode).lower()
    if mode == "auto":
        if not tc_map_available():
            # a library without the tcgen05 kernel is a broken build, not something to paper over with the
            # 50x slower CUDA-core kernel (ask for "simt_f32" / "simt_f64" explicitly if that is what you want)
            raise N.NativeError("libb2dsp.so was built without the tensor-core map kernel (das_map_tc.cu)")
        return N.MAP_TC_F16X3
    if mode not in _MAP_MODES:
        raise ValueError(f"unknown map mode {mode!r}")
    return _MAP_MODES[mode]


# --------------------------------------------------------------------------- helpers
def welch_frames(length: int, wl: int, step: int) -> tuple[int, int]:
    """(frames, zero padding) -- reference helpers/other.py:181-213 with zero_padding=True."""
    k, p = ctypes.c_int64(), ctypes.c_int64()
    N.check(N.lib().b2_welch_frames(length, wl, step, ctypes.byref(k), ctypes.byref(p)), "welch_frames")
    return k.value, p.value


_twiddle_cache: dict = {}

# Stage-1 temporaries (channel-major copy, spectra) and the map operand are multi-GB buffers that never leave the
# engine: they are kept per (role, device) and reused while the size fits, instead of going back to torch's caching
# allocator, whose block splitting turned the differently-sized requests of consecutive steps into cudaMalloc /
# cudaFree cycles (+40 ms per step at 256 mics x 60 s).  `release_workspaces()` drops them.
_workspaces: dict = {}


def _workspace(role: str, nbytes: int, device):
    torch = _torch()
    key = (role, str(device))
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes or buf.numel() > 2 * nbytes + (64 << 20):
        _workspaces.pop(key, None)
        buf = None
        buf = _workspaces[key] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return buf


def _workspace_tensor(role: str, shape, dtype, device):
    torch = _torch()
    n = int(np.prod(shape)) * torch.empty((), dtype=dtype).element_size()
    return _workspace(role, max(n, 16), device)[:n].view(dtype).reshape(shape)


def release_workspaces():
    """Free the engine's cached multi-GB temporaries (they are re-created on the next call)."""
    _workspaces.clear()


def _twiddles(wl: int, prec: int, device):
    torch = _torch()
    key = (wl, prec, str(device))
    tw = _twiddle_cache.get(key)
    if tw is None:
        tw = torch.empty(wl, dtype=torch.complex128 if prec == N.F64 else torch.complex64, device=device)
        N.check(N.lib().b2_fft_twiddles(wl, prec, N.ptr(tw), N.stream_handle()), "b2_fft_twiddles")
        _twiddle_cache[key] = tw
    return tw


class _Stager:
    """Pageable host memory -> device through pinned staging buffers.

    A `cudaMemcpy` from pageable memory is staged by the driver through one bounce buffer with one copying
    thread (measured: 13.7 GB/s for the 5.9 GB recording of the benchmark against 50+ GB/s from pinned memory).
    Here several host threads copy slices into a ring of pinned slots (numpy releases the GIL for the copy) and
    every filled slot goes to the device with an asynchronous copy on the current stream; a slot is reused when
    the event recorded behind its copy has passed."""

    SLOT_BYTES = 32 << 20
    SLOTS = 8

    def __init__(self):
        self.ring = None
        self.pool = None
        self.workers = 0

    def upload(self, src: np.ndarray, device, out=None):
        """`src`: C-contiguous numpy array; returns a CUDA tensor of the same shape and dtype (`out` if given)."""
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor

        torch = _torch()
        if self.ring is None:
            self.ring = torch.empty(self.SLOTS * self.SLOT_BYTES, dtype=torch.uint8).pin_memory()
            self.ring_np = self.ring.numpy()
        workers = max(1, int(config.h2d_stage_threads))
        if self.pool is None or self.workers != workers:
            if self.pool is not None:
                self.pool.shutdown(wait=True)
            self.pool, self.workers = ThreadPoolExecutor(max_workers=workers), workers
        if out is None:
            out = torch.empty(src.shape, dtype=torch.from_numpy(src[:0]).dtype, device=device)
        src_b = src.reshape(-1).view(np.uint8)
        dst_b = out.reshape(-1).view(torch.uint8)
        n, sb = src_b.nbytes, self.SLOT_BYTES
        pieces = [(o, min(sb, n - o)) for o in range(0, n, sb)]
        stream = torch.cuda.current_stream(device)
        events = [None] * self.SLOTS
        fill = lambda s, o, m: np.copyto(self.ring_np[s * sb:s * sb + m], src_b[o:o + m])   # noqa: E731
        pending, nxt = deque(), 0
        for i, (o, m) in enumerate(pieces):
            while nxt < len(pieces) and nxt < i + self.SLOTS:
                slot = nxt % self.SLOTS
                if events[slot] is not None:
                    events[slot].synchronize()               # the device copy out of this slot has finished
                pending.append(self.pool.submit(fill, slot, *pieces[nxt]))
                nxt += 1
            pending.popleft().result()
            slot = i % self.SLOTS
            dst_b[o:o + m].copy_(self.ring[slot * sb:slot * sb + m], non_blocking=True)
            events[slot] = torch.cuda.Event()
            events[slot].record(stream)
        return out


    def download(self, t) -> np.ndarray:
        """Contiguous CUDA tensor -> fresh (pageable) numpy array through the same ring, the other way round:
        asynchronous device copies into pinned slots, host threads empty them into the result."""
        from collections import deque
        from concurrent.futures import ThreadPoolExecutor

        torch = _torch()
        if self.ring is None:
            self.ring = torch.empty(self.SLOTS * self.SLOT_BYTES, dtype=torch.uint8).pin_memory()
            self.ring_np = self.ring.numpy()
        workers = max(1, int(config.h2d_stage_threads))
        if self.pool is None or self.workers != workers:
            if self.pool is not None:
                self.pool.shutdown(wait=True)
            self.pool, self.workers = ThreadPoolExecutor(max_workers=workers), workers
        t = t.contiguous()
        out = np.empty(tuple(t.shape), dtype=torch.empty((), dtype=t.dtype).numpy().dtype)
        src_b = t.reshape(-1).view(torch.uint8)
        dst_b = out.reshape(-1).view(np.uint8)
        n, sb = dst_b.nbytes, self.SLOT_BYTES
        pieces = [(o, min(sb, n - o)) for o in range(0, n, sb)]
        stream = torch.cuda.current_stream(t.device)
        drain = lambda s, o, m: np.copyto(dst_b[o:o + m], self.ring_np[s * sb:s * sb + m])   # noqa: E731
        copies, emptied, nxt = deque(), [None] * self.SLOTS, 0
        for i, (o, m) in enumerate(pieces):
            while nxt < len(pieces) and nxt < i + self.SLOTS:
                slot = nxt % self.SLOTS
                if emptied[slot] is not None:
                    emptied[slot].result()                   # the host copy out of this slot has finished
                po, pm = pieces[nxt]
                self.ring[slot * sb:slot * sb + pm].copy_(src_b[po:po + pm], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(stream)
                copies.append(ev)
                nxt += 1
            copies.popleft().synchronize()
            emptied[i % self.SLOTS] = self.pool.submit(drain, i % self.SLOTS, o, m)
        for f in emptied:
            if f is not None:
                f.result()
        return out


_stager = _Stager()


def to_host(t) -> np.ndarray:
    """CUDA tensor -> numpy array (what the reference API returns): large results go through the pinned staging
    ring (a plain `.cpu()` into pageable memory runs at ~10 GB/s), small ones directly."""
    if not (config.h2d_stage_threads > 0 and getattr(t, "is_cuda", False)
            and t.numel() * t.element_size() >= config.h2d_stage_min_bytes):
        return t.cpu().numpy()
    torch = _torch()
    if t.is_complex():
        real = _stager.download(torch.view_as_real(t.contiguous()))
        return real.view(np.complex128 if t.dtype == torch.complex128 else np.complex64).reshape(tuple(t.shape))
    return _stager.download(t)


def _host_to_device(arr: np.ndarray, device, slot: str | None = None):
    """C-contiguous host array -> CUDA tensor: asynchronous from pinned memory, staged by `_Stager` from large
    pageable arrays, the driver's own path otherwise.  With `slot` the destination is that workspace buffer
    (the caller orders its reuse) instead of a fresh allocation."""
    torch = _torch()
    t = torch.from_numpy(arr)
    out = None
    if slot is not None:
        buf = _workspaces.get((slot, str(device)))                 # never shrunk: the last chunk of a call is the small one
        if buf is None or buf.numel() < arr.nbytes:
            if buf is not None:
                torch.cuda.synchronize(device)                         # rare (a larger recording): nobody reads the old one
            _workspaces.pop((slot, str(device)), None)
            buf = _workspaces[(slot, str(device))] = torch.empty(max(arr.nbytes, 16), dtype=torch.uint8, device=device)
        out = buf[:arr.nbytes].view(t.dtype).reshape(arr.shape)
    if (config.h2d_stage_threads > 0 and arr.nbytes >= config.h2d_stage_min_bytes and not t.is_pinned()):
        return _stager.upload(arr, device, out)
    if out is None:
        return t.to(device, non_blocking=True)
    out.copy_(t, non_blocking=True)
    return out


def _as_device_block(x, t0: int, t1: int, device, slot: str | None = None):
    """Upload rows [t0, t1) of the (L, M) recording without changing its memory order.

    Returns (tensor, in_dtype code, stride_t, stride_m) for b2_prepare_channels.  `x` may be a
    numpy array (host; C- or F-ordered views are uploaded as they lie) or a CUDA tensor.
    """
    torch = _torch()
    if isinstance(x, torch.Tensor):
        if x.dtype not in (torch.float32, torch.float64):
            x = x.to(torch.float64)
        xs = x[t0:t1]
        if xs.device.type != "cuda":
            xs = xs.to(device, non_blocking=True)
        code = N.F64 if xs.dtype == torch.float64 else N.F32
        return xs, code, max(1, xs.stride(0)), max(1, xs.stride(1))
    x = np.asarray(x)
    if x.dtype not in (np.float32, np.float64):
        x = x.astype(np.float64)          # the reference upcasts when framing
    code = N.F64 if x.dtype == np.float64 else N.F32
    xs = x[t0:t1]
    if xs.flags.c_contiguous:
        return _host_to_device(xs, device, slot), code, xs.shape[1], 1
    if xs.T.flags.c_contiguous:          # (channels, time) storage seen through Signal's transpose
        return _host_to_device(xs.T, device, slot), code, 1, xs.shape[0]
    return _host_to_device(np.ascontiguousarray(xs), device, slot), code, xs.shape[1], 1


class _Timer:
    """CUDA-event stopwatch on the current stream, active only under config.profile."""

    def __init__(self):
        self.marks = []

    def mark(self, name):
        if config.profile:
            torch = _torch()
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.marks.append((name, e))

    def report(self) -> dict:
        if not config.profile or len(self.marks) < 2:
            return {}
        torch = _torch()
        torch.cuda.synchronize()
        out = {}
        for (_, a), (n, b) in zip(self.marks[:-1], self.marks[1:]):
            out[n] = out.get(n, 0.0) + a.elapsed_time(b)
        return out


# --------------------------------------------------------------------------- stage 1
def packed_size(M: int) -> int:
    """Entries of the packed lower triangle of an M x M Hermitian matrix: (j, i), j >= i, at j(j+1)/2 + i."""
    return M * (M + 1) // 2


def unpack_lower(S, M: int):
    """(nb, M, M) complex128 tensor whose lower triangle holds the packed sums `S` (nb, M(M+1)/2); the strict
    upper triangle is zero.  Debug / test helper -- the kernels consume the packed form directly."""
    torch = _torch()
    rows, cols = torch.tril_indices(M, M, device=S.device)          # row-major: index j(j+1)/2 + i
    out = torch.zeros((S.shape[0], M, M), dtype=S.dtype, device=S.device)
    out[:, rows, cols] = S
    return out


def _frame_chunks(k0: int, k1: int, tile: int, max_frames: int, host_bytes: int) -> list[tuple[int, int]]:
    """Frame ranges processed one after the other: bounded by the spectra budget and, for host recordings,
    small enough (config.h2d_chunk_bytes) that the upload of one range overlaps the kernels of the previous."""
    n = k1 - k0
    per = max_frames
    if host_bytes > 0 and config.h2d_chunk_bytes > 0 and host_bytes > config.h2d_chunk_bytes:
        parts = -(-host_bytes // config.h2d_chunk_bytes)
        frames_per_part = -(-n // parts)
        per = min(per, max(tile, -(-frames_per_part // tile) * tile))
    out, kc = [], k0
    while kc < k1:
        kn = min(k1, kc + per)
        out.append((kc, kn))
        kc = kn
    return out


def csm_partial_sums(x, length: int, wl: int, step: int, window: np.ndarray, detrend: bool,
                     fft_scale: float, frame_range: tuple[int, int] | None = None,
                     precision: str | None = None, device=None, row_offset: int = 0, diag_only: bool = False,
                     reduce_fn=None):
    """Un-normalised per-bin sums S_f = sum_k X_k X_k^H over the frames in `frame_range`.

    Returns (S, n_frames_total): S is a complex128 CUDA tensor (nb, M(M+1)/2) holding the PACKED lower
    triangle (row >= col, see `packed_size` / `unpack_lower`).  Partial sums of disjoint frame ranges add up
    to the full sum, which is what makes the stage frame-shardable.  `x` may hold only
    the rows [row_offset, row_offset + x.shape[0]) of the recording of `length` samples (a rank's shard).
    With `diag_only` only the auto-spectra are accumulated: S is then float64 (nb, M).

    `reduce_fn(t) -> handle | None` (dist.py) sums a slice of S over the ranks IN PLACE and may return an
    object with `.wait()`: the bins are processed in `config.reduce_ranges` ranges and the reduction of one
    range is started as soon as its rank-K update is enqueued, so it overlaps the update of the next range.
    """
    torch = _torch()
    lib = N.lib()
    prec = _PREC[(precision or config.csm_precision).lower()]
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    with torch.cuda.device(device):
        return _csm_partial_sums(torch, lib, prec, device, x, length, wl, step, window, detrend, fft_scale,
                                 frame_range, row_offset, diag_only, reduce_fn)


def _csm_partial_sums(torch, lib, prec, device, x, length, wl, step, window, detrend, fft_scale, frame_range,
                      row_offset, diag_only, reduce_fn):
    M = int(x.shape[1])
    K_total, _ = welch_frames(length, wl, step)
    k0, k1 = frame_range if frame_range is not None else (0, K_total)
    assert 0 <= k0 <= k1 <= K_total, "frame range outside the recording"
    nb = wl // 2 + 1
    rdt = torch.float64 if prec == N.F64 else torch.float32
    cdt = torch.complex128 if prec == N.F64 else torch.complex64
    cur = torch.cuda.current_stream()
    st = cur.cuda_stream
    timer = _Timer()

    tile = N.check(lib.b2_welch_frame_tile(wl, prec), "b2_welch_frame_tile")
    tile = max(16, tile)
    bytes_per_frame = (nb + 3) * M * (16 if prec == N.F64 else 8)      # bins + 3 unused slots (spectra layout)
    max_frames = max(tile, (config.spectra_budget_bytes // bytes_per_frame) // tile * tile)
    on_host = not (isinstance(x, torch.Tensor) and x.device.type == "cuda")
    itemsize = x.element_size() if isinstance(x, torch.Tensor) else int(getattr(getattr(x, "dtype", None), "itemsize", 8))
    host_bytes = (k1 - k0) * step * M * itemsize if on_host else 0
    chunks = _frame_chunks(k0, k1, tile, max_frames, host_bytes)

    win = _device_const(window, device).to(rdt)          # cached: a pageable upload would stall the host behind the stream
    tw = _twiddles(wl, prec, device)
    S = (torch.empty((nb, M), dtype=torch.float64, device=device) if diag_only
         else torch.empty((nb, packed_size(M)), dtype=torch.complex128, device=device))
    timer.mark("start")
    if k1 == k0:
        S.zero_()

    def rows_of(c):
        kc, kn = c
        t0 = kc * step
        t1 = min(length, (kn - 1) * step + wl)           # samples beyond `length` are the zero padding
        assert t0 >= row_offset and t1 - row_offset <= x.shape[0], "frame range reads samples outside this shard"
        return t0, t1

    # the upload of chunk i+1 is issued (on a copy stream) right after the kernels of chunk i were enqueued.
    # numpy recordings land in two alternating workspace buffers (fresh 768 MB tensors per chunk made the caching
    # allocator cudaMalloc 5 GB more on every call: profiles/r2_e2e_stage1_diag.txt); a buffer is written again
    # only after the K0 launch that read it has finished (`consumed`).
    copy_stream = torch.cuda.Stream(device=device) if (on_host and len(chunks) > 1) else None
    slots = ("h2d0", "h2d1") if isinstance(x, np.ndarray) else (None, None)
    consumed = [None, None]
    if copy_stream is not None:
        copy_stream.wait_stream(cur)                     # earlier calls may still be reading the buffers

    def upload(ci):
        t0, t1 = rows_of(chunks[ci])
        if copy_stream is None:
            return _as_device_block(x, t0 - row_offset, t1 - row_offset, device, slots[ci % 2]), None
        with torch.cuda.stream(copy_stream):
            if consumed[ci % 2] is not None:
                copy_stream.wait_event(consumed[ci % 2])
            blk = _as_device_block(x, t0 - row_offset, t1 - row_offset, device, slots[ci % 2])
            ev = torch.cuda.Event()
            ev.record(copy_stream)
        return blk, ev

    pending = []                                         # handles of the overlapped reductions
    nxt = upload(0) if chunks else None
    for ci, (kc, kn) in enumerate(chunks):
        (xs, code, s_t, s_m), ev = nxt
        if ev is not None:
            cur.wait_event(ev)
            if slots[0] is None:
                xs.record_stream(cur)
        nfr = kn - kc
        Kp = (nfr + tile - 1) // tile * tile
        t0, t1 = rows_of((kc, kn))
        Lp = (Kp - 1) * step + wl
        Lp += (-Lp) % 4
        timer.mark("h2d")
        xc = _workspace_tensor("xc", (M, Lp), rdt, device)
        N.check(lib.b2_prepare_channels(N.ptr(xs), code, s_t, s_m, 0, t1 - t0, M, N.ptr(xc), Lp, prec, st),
                "b2_prepare_channels")
        _count()
        if slots[0] is not None:
            consumed[ci % 2] = torch.cuda.Event()
            consumed[ci % 2].record(cur)
        timer.mark("k0_prepare")
        X = _workspace_tensor("spectra", (_spectra_elems(lib, wl, M, Kp),), cdt, device)
        N.check(lib.b2_welch_rfft(N.ptr(xc), Lp, M, wl, step, nfr, N.ptr(win), N.ptr(tw), int(bool(detrend)),
                                  float(fft_scale), N.ptr(X), Kp, prec, st), "b2_welch_rfft")
        _count()
        timer.mark("k1_rfft")
        acc = 0 if ci == 0 else 1
        last = ci == len(chunks) - 1
        if diag_only:
            N.check(lib.b2_welch_autopower(N.ptr(X), M, Kp, nb, N.ptr(S), acc, prec, st), "b2_welch_autopower")
            _count()
        else:
            nr = max(1, min(int(config.reduce_ranges), nb)) if (reduce_fn is not None and last) else 1
            for r in range(nr):
                b0, b1 = nb * r // nr, nb * (r + 1) // nr
                N.check(lib.b2_csm_herk(N.ptr(X), M, Kp, nb, b0, b1 - b0, N.ptr(S), acc, prec, st), "b2_csm_herk")
                _count()
                if reduce_fn is not None and last:
                    pending.append(reduce_fn(S[b0:b1]))
        timer.mark("k2_herk")
        if ci + 1 < len(chunks):
            nxt = upload(ci + 1)
        del xs, xc, X
    if reduce_fn is not None:
        if not chunks or diag_only:
            pending.append(reduce_fn(S))
        for h in pending:
            if h is not None:
                h.wait()
        timer.mark("allreduce_tail")
    config.stats["csm"] = dict(times_ms=timer.report(), frames=k1 - k0, frames_total=K_total, M=M, nb=nb,
                               precision="fp64" if prec == N.F64 else "fp32", frame_chunks=len(chunks))
    return S, K_total


def _spectra_elems(lib, wl: int, M: int, Kp: int) -> int:
    n = int(lib.b2_welch_spectra_elems(wl, M, Kp))
    if n < 0:
        raise N.NativeError(lib.b2_last_error().decode())
    return n


def frame_spectra(x, length: int, wl: int, step: int, window: np.ndarray, detrend: bool, fft_scale: float,
                  precision: str | None = None, device=None):
    """Spectra of ALL frames, complex128 CUDA tensor (bins, frames, channels): the framing / window /
    mean-removal / rFFT kernels of the Welch stage (K0, K1) without the averaging -- what the STFT
    (reference `_stft`, _spectral_methods.py:176-282) keeps.  Frames follow helpers/other.py:181-213
    (ceil(L/step) frames, zero-padded tail).  torch is only used to re-order the kernel's chunk-major
    spectra layout (include/b2dsp.h, b2_welch_rfft) into the reference's (bins, frames, channels).
    """
    torch = _torch()
    lib = N.lib()
    prec = _PREC[(precision or config.csm_precision).lower()]
    device = device or torch.device("cuda", torch.cuda.current_device())
    M = int(x.shape[1])
    K_total, _ = welch_frames(length, wl, step)
    nb = wl // 2 + 1
    rdt = torch.float64 if prec == N.F64 else torch.float32
    cdt = torch.complex128 if prec == N.F64 else torch.complex64
    st = N.stream_handle()
    tile = max(16, N.check(lib.b2_welch_frame_tile(wl, prec), "b2_welch_frame_tile"))
    bytes_per_frame = (nb + 3) * M * (16 if prec == N.F64 else 8)
    max_frames = max(tile, (config.spectra_budget_bytes // bytes_per_frame) // tile * tile)
    slots = _device_const(np.array([lib.b2_welch_bin_slot(wl, f) for f in range(nb)], dtype=np.float64), device).long()
    win = _device_const(window, device).to(rdt)          # cached: a pageable upload would stall the host behind the stream
    tw = _twiddles(wl, prec, device)
    out = torch.empty((nb, K_total, M), dtype=torch.complex128, device=device)
    kc = 0
    while kc < K_total:
        kn = min(K_total, kc + max_frames)
        nfr = kn - kc
        Kp = (nfr + tile - 1) // tile * tile
        t0 = kc * step
        t1 = min(length, (kn - 1) * step + wl)
        Lp = (Kp - 1) * step + wl
        Lp += (-Lp) % 4
        xs, code, s_t, s_m = _as_device_block(x, t0, t1, device)
        xc = torch.empty((M, Lp), dtype=rdt, device=device)
        N.check(lib.b2_prepare_channels(N.ptr(xs), code, s_t, s_m, 0, t1 - t0, M, N.ptr(xc), Lp, prec, st),
                "b2_prepare_channels")
        _count()
        X = torch.empty((Kp // 16, _spectra_elems(lib, wl, M, Kp) // (Kp // 16 * M * 64), M, 8, 4, 2), dtype=cdt, device=device)
        N.check(lib.b2_welch_rfft(N.ptr(xc), Lp, M, wl, step, nfr, N.ptr(win), N.ptr(tw), int(bool(detrend)),
                                  float(fft_scale), N.ptr(X), Kp, prec, st), "b2_welch_rfft")
        _count()
        # [chunk][group][channel][pair][slot][frame] -> (slots, frames, channels), then the bins' slots
        by_slot = X.permute(1, 4, 0, 3, 5, 2).reshape(-1, Kp, M)
        out[:, kc:kn, :] = by_slot.index_select(0, slots)[:, :nfr, :]
        kc = kn
        del xs, xc, X
    return out


# --------------------------------------------------------------------------- large / arbitrary-length FFT
_fft_plans: dict = {}


def _fft_plan(n: int, device):
    torch = _torch()
    lib = N.lib()
    key = (int(n), str(device))
    plan = _fft_plans.get(key)
    if plan is None:
        nbytes = N.check(lib.b2_fft_plan_bytes(n), "b2_fft_plan_bytes")
        plan = torch.empty(nbytes, dtype=torch.uint8, device=device)
        work = _workspace("fft_work", N.check(lib.b2_fft_work_bytes(n, 1), "b2_fft_work_bytes"), device)
        N.check(lib.b2_fft_plan(n, N.ptr(plan), N.ptr(work), N.stream_handle()), "b2_fft_plan")
        _count(3)
        if len(_fft_plans) >= 8:
            _fft_plans.pop(next(iter(_fft_plans)))
        _fft_plans[key] = plan
    return plan


def rfft_rows(x_dev, n: int, scale: float = 1.0, batch_bytes: int = 6 << 30):
    """np.fft.rfft(x, n=n, axis=0) * scale for a (L, M) float32/float64 CUDA tensor -> complex128 CUDA (n//2+1, M).
    Any length n (cropping or zero-padding the L samples like numpy); channels are processed in groups that keep the
    two work buffers under `batch_bytes`."""
    torch = _torch()
    lib = N.lib()
    assert x_dev.is_cuda and x_dev.ndim == 2 and x_dev.dtype in (torch.float32, torch.float64)
    device = x_dev.device
    L, M = int(x_dev.shape[0]), int(x_dev.shape[1])
    nb = n // 2 + 1
    out = torch.empty((nb, M), dtype=torch.complex128, device=device)
    with torch.cuda.device(device):
        plan = _fft_plan(n, device)
        per_row = N.check(lib.b2_fft_work_bytes(n, 1), "b2_fft_work_bytes")
        group = max(1, min(M, batch_bytes // per_row))
        code = N.F64 if x_dev.dtype == torch.float64 else N.F32
        st, sb = x_dev.stride(0), x_dev.stride(1)
        for m0 in range(0, M, group):
            b = min(group, M - m0)
            work = _workspace("fft_work", N.check(lib.b2_fft_work_bytes(n, b), "b2_fft_work_bytes"), device)
            xs = x_dev[:, m0:m0 + b]
            N.check(lib.b2_rfft_any(N.ptr(xs), code, st, sb, 0, b, L, n, b, None, 0, N.ptr(plan), N.ptr(work), float(scale),
                                    N.ptr(out[:, m0:m0 + b]), M, 1, 0, N.stream_handle()), "b2_rfft_any")
            _count()
    return out


def rfft_frames(xc, Lp: int, M: int, wl: int, step: int, n_frames: int, n: int, window, detrend: bool, scale: float,
                batch_bytes: int = 6 << 30):
    """Spectra of all frames for an FFT length n != wl (or beyond the shared-memory kernel): complex128 CUDA tensor
    (n//2+1, n_frames, M) = np.fft.rfft(windowed [, de-meaned] frames, n=n, axis=0) of reference `_stft`
    (_spectral_methods.py:257-263).  `xc` is the channel-major zero-padded copy (M, Lp) from b2_prepare_channels."""
    torch = _torch()
    lib = N.lib()
    device = xc.device
    nb = n // 2 + 1
    out = torch.empty((nb, n_frames, M), dtype=torch.complex128, device=device)
    win = _device_const(window, device)
    code = N.F64 if xc.dtype == torch.float64 else N.F32
    with torch.cuda.device(device):
        plan = _fft_plan(n, device)
        per_row = N.check(lib.b2_fft_work_bytes(n, 1), "b2_fft_work_bytes")
        ch_group = max(1, min(M, batch_bytes // (per_row * n_frames)))
        for m0 in range(0, M, ch_group):
            mc = min(ch_group, M - m0)
            b = mc * n_frames
            work = _workspace("fft_work", N.check(lib.b2_fft_work_bytes(n, b), "b2_fft_work_bytes"), device)
            # row (m, k): input at m * Lp + k * step, output out[f][k][m0 + m]
            N.check(lib.b2_rfft_any(N.ptr(xc[m0:m0 + mc]), code, 1, step, Lp, n_frames, wl, n, b, N.ptr(win), int(bool(detrend)),
                                    N.ptr(plan), N.ptr(work), float(scale), N.ptr(out[:, :, m0:m0 + mc]), n_frames * M, M, 1,
                                    N.stream_handle()), "b2_rfft_any")
            _count()
    return out


def frame_spectra_any(x, length: int, wl: int, step: int, window: np.ndarray, detrend: bool, n_fft: int, fft_scale: float,
                      device=None):
    """Like `frame_spectra` for an FFT length different from the window length (frames cropped or zero-padded to
    n_fft, reference _spectral_methods.py:263) or a window beyond the shared-memory kernel: K0 builds the
    channel-major zero-padded copy, the large / arbitrary-length FFT transforms every frame in float64.
    Returns complex128 (n_fft//2+1, frames, channels)."""
    torch = _torch()
    lib = N.lib()
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    M = int(x.shape[1])
    K_total, _ = welch_frames(length, wl, step)
    with torch.cuda.device(device):
        Lp = (K_total - 1) * step + wl
        Lp += (-Lp) % 4
        xs, code, s_t, s_m = _as_device_block(x, 0, length, device)
        xc = torch.empty((M, Lp), dtype=torch.float64, device=device)
        N.check(lib.b2_prepare_channels(N.ptr(xs), code, s_t, s_m, 0, length, M, N.ptr(xc), Lp, N.F64, N.stream_handle()),
                "b2_prepare_channels")
        _count()
        return rfft_frames(xc, Lp, M, wl, step, K_total, n_fft, window, detrend, fft_scale)


def delay_mix(signals, taps: np.ndarray, shifts: np.ndarray, out_length: int, device=None):
    """Mixed array signals of several sources: float64 CUDA tensor (out_length, M),
    out[n, m] = sum_s sum_t taps[s, m, t] x_s[n - shifts[s, m] - t] (x_s zero outside its samples).

    signals: one 1-D array / tensor per source (host or device, any float type); taps (S, M, T) float64 with the
    gains folded in; shifts (S, M) integers.  The per-microphone loop of `MonopoleSource.get_signals_on_array` and
    the per-source sum of `mix_sources_on_array` (reference beamforming/beamforming.py:1427-1511) as one kernel."""
    torch = _torch()
    device = torch.device(device) if device is not None else torch.device("cuda", torch.cuda.current_device())
    S, M, T = taps.shape
    assert len(signals) == S and shifts.shape == (S, M)
    with torch.cuda.device(device):
        parts, lens = [], []
        for x in signals:
            t = x if isinstance(x, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(x))
            t = t.reshape(-1).to(device=device, dtype=torch.float64, non_blocking=True)
            parts.append(t)
            lens.append(int(t.numel()))
        xcat = parts[0] if S == 1 else torch.cat(parts)
        offs = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.int64)
        meta = torch.from_numpy(np.concatenate([offs, np.asarray(lens, dtype=np.int64)])).to(device)
        taps_d = torch.from_numpy(np.ascontiguousarray(np.transpose(taps, (0, 2, 1)), dtype=np.float64)).to(device)   # [S][T][M]
        shifts_d = torch.from_numpy(np.ascontiguousarray(shifts, dtype=np.int32)).to(device)
        out = torch.empty((int(out_length), M), dtype=torch.float64, device=device)
        N.check(N.lib().b2_delay_mix(N.ptr(xcat), N.ptr(meta), N.ptr(meta[S:]), S, N.ptr(taps_d), N.ptr(shifts_d), T, M,
                                     int(out_length), N.ptr(out), N.stream_handle()), "b2_delay_mix")
    _count()
    return out


def cleansc_component(D0, D1, peak, pmax, active, k_dev, mic_dev, center_xyz, grid_dev, steering: int,
                      safety_factor: float, remove_diag: bool, h_out, scale_out):
    """One CLEAN-SC step of every bin on the degraded CSMs (in place): see b2_cleansc_component in include/b2dsp.h."""
    F, M = int(D1.shape[0]), int(D1.shape[1])
    assert D0.is_contiguous() and D1.is_contiguous() and h_out.is_contiguous() and active.dtype == _torch().bool
    with _torch().cuda.device(D1.device):
        N.check(N.lib().b2_cleansc_component(N.ptr(D0), N.ptr(D1), M, F, N.ptr(peak), N.ptr(pmax), N.ptr(active),
                                             N.ptr(k_dev), N.ptr(mic_dev), N.vec3(center_xyz), N.ptr(grid_dev),
                                             int(steering), float(safety_factor), int(bool(remove_diag)), N.ptr(h_out),
                                             N.ptr(scale_out), N.stream_handle()), "b2_cleansc_component")
    _count()


def das_rank1_update(vectors, scale, freqs_dev, c: float, mic_dev, center_xyz, grid_dev, steering: int,
                     remove_diag: bool, alpha: float, out):
    """out[g, f] += alpha * scale[f] * Re(h_g^H (v_f v_f^H [minus its diagonal]) h_g), float64, in O(G F M)."""
    F, M = int(vectors.shape[0]), int(vectors.shape[1])
    G = int(grid_dev.shape[0])
    assert out.is_contiguous() and tuple(out.shape) == (G, F) and vectors.is_contiguous()
    with _torch().cuda.device(out.device):
        N.check(N.lib().b2_das_rank1_update(N.ptr(vectors), N.ptr(scale), M, F, N.ptr(freqs_dev), float(c), N.ptr(mic_dev),
                                            N.vec3(center_xyz), N.ptr(grid_dev), G, int(steering), int(bool(remove_diag)),
                                            float(alpha), N.ptr(out), N.stream_handle()), "b2_das_rank1_update")
    _count()


def hermitian_eig(matrices):
    """(w, v) of a batch of Hermitian matrices, complex128 CUDA tensor (nb, M, M): eigenvalues in ASCENDING order
    (nb, M) float64 and the eigenvectors as columns (nb, M, M) -- the contract of np.linalg.eigh, which the reference
    calls per bin (beamforming/beamforming.py:1107).  Only the lower triangle is read.  Orders up to
    `b2_heig_max_order()` run in the shared-memory Jacobi kernel (all bins at once, one CTA each); larger matrices
    go to torch.linalg.eigh (cuSOLVER), one of the O(F M^3) per-bin factorisations outside the hot loop."""
    torch = _torch()
    lib = N.lib()
    nb, M = int(matrices.shape[0]), int(matrices.shape[1])
    if M > int(lib.b2_heig_max_order()):
        return torch.linalg.eigh(matrices)
    a = matrices.contiguous()
    w = torch.empty((nb, M), dtype=torch.float64, device=a.device)
    v = torch.empty((nb, M, M), dtype=torch.complex128, device=a.device)
    with torch.cuda.device(a.device):
        N.check(lib.b2_heig_jacobi(N.ptr(a), M, nb, N.ptr(w), N.ptr(v), None, N.stream_handle()), "b2_heig_jacobi")
    _count()
    w, order = torch.sort(w, dim=1)
    return w, torch.gather(v, 2, order[:, None, :].expand(nb, M, M))


def csm_rank1(spectrum_dev, factor: float, halve_edges: bool, amplitude_sqrt: bool):
    """(nb, M, M) complex128 CSM of one spectrum per channel -- `_csm_fft` (_spectral_methods.py:374-443)."""
    torch = _torch()
    nb, M = spectrum_dev.shape
    out = torch.empty((nb, M, M), dtype=torch.complex128, device=spectrum_dev.device)
    with torch.cuda.device(spectrum_dev.device):
        N.check(N.lib().b2_csm_rank1(N.ptr(spectrum_dev.contiguous()), M, nb, float(factor), int(bool(halve_edges)),
                                     int(bool(amplitude_sqrt)), N.ptr(out), N.stream_handle()), "b2_csm_rank1")
    _count()
    return out


def csm_finalize(S, M: int, n_frames: int, factor: float, halve_edges: bool, amplitude_sqrt: bool):
    """Average / scale / sqrt / mirror the packed summed spectra (nb, M(M+1)/2) -> complex128 (nb, M, M)
    CUDA tensor."""
    torch = _torch()
    nb = S.shape[0]
    assert S.shape[1] == packed_size(M), "packed sums do not match the channel count"
    out = torch.empty((nb, M, M), dtype=torch.complex128, device=S.device)
    N.check(N.lib().b2_csm_finalize(N.ptr(S), M, nb, 1.0 / float(n_frames), float(factor), int(bool(halve_edges)),
                                    int(bool(amplitude_sqrt)), N.ptr(out), N.stream_handle()), "b2_csm_finalize")
    _count()
    return out


def auto_finalize(P, n_frames: int, factor: float, halve_edges: bool, amplitude_sqrt: bool):
    """Average / scale / sqrt the summed auto-spectra -> float64 (nb, M) CUDA tensor."""
    torch = _torch()
    nb, M = P.shape
    out = torch.empty_like(P)
    N.check(N.lib().b2_auto_finalize(N.ptr(P), M, nb, 1.0 / float(n_frames), float(factor), int(bool(halve_edges)),
                                     int(bool(amplitude_sqrt)), N.ptr(out), N.stream_handle()), "b2_auto_finalize")
    _count()
    return out


# --------------------------------------------------------------------------- stage 2
def das_prepare(csm_dev, bin0: int, nbins: int, remove_diag: bool, mode: int):
    """Build the map kernel's operand for bins [bin0, bin0+nbins) of a device CSM."""
    torch = _torch()
    lib = N.lib()
    nb, M, _ = csm_dev.shape
    nbytes = N.check(lib.b2_das_operand_bytes(M, nbins, mode), "b2_das_operand_bytes")
    # one operand buffer per device: a map call consumes it before the next one is prepared (same stream)
    op = _workspace("map_operand", nbytes, csm_dev.device)[:nbytes]
    N.check(lib.b2_das_prepare(N.ptr(csm_dev), M, nb, bin0, nbins, int(bool(remove_diag)), N.ptr(op), mode,
                               N.stream_handle()), "b2_das_prepare")
    _count()
    return op


_const_cache: dict = {}


def _device_const(arr, device):
    """float64 device copy of a SMALL host array (window, bin frequencies, Simpson weights, microphone positions),
    cached by content so that repeated calls do not re-upload them from pageable memory: such a copy makes the host
    wait until everything queued on the stream has finished, and the kernels enqueued after it then start late
    (measured: +55 ms per step of stage 1 when it follows a 1.6 s map kernel, profiles/r2_stage1_host_stall.txt)."""
    torch = _torch()
    a = np.ascontiguousarray(arr, dtype=np.float64)
    if a.nbytes > (1 << 20):
        return torch.from_numpy(a).to(device)
    key = (a.shape, a.tobytes(), str(device))
    t = _const_cache.get(key)
    if t is None:
        if len(_const_cache) >= 64:
            _const_cache.pop(next(iter(_const_cache)))
        t = _const_cache[key] = torch.from_numpy(a).to(device)
    return t


def das_map(op, M: int, freqs: np.ndarray, weights: np.ndarray, c: float, mic_xyz: np.ndarray,
            center_xyz: np.ndarray, grid_xyz, steering: int, clamp_negative: bool, mode: int,
            per_bin: bool = False):
    """Run K3 over all rows of `grid_xyz` ((G,3) numpy or CUDA float64 tensor).

    Returns (map [G] float64 CUDA tensor, per-bin [G, F] tensor or None).  `op` (from `das_prepare`) also
    holds the kernel's per-CTA scratch tables: it must not be shared by two launches in flight at once.
    """
    torch = _torch()
    lib = N.lib()
    device = op.device
    nbins = len(freqs)
    f_dev = _device_const(freqs, device)
    w_dev = _device_const(weights, device)
    mic_dev = _device_const(mic_xyz, device)
    if isinstance(grid_xyz, torch.Tensor):
        g_dev = grid_xyz.to(device=device, dtype=torch.float64).contiguous()
    else:
        g_dev = torch.from_numpy(np.ascontiguousarray(grid_xyz, dtype=np.float64)).to(device)
    G = g_dev.shape[0]
    out = torch.empty(G, dtype=torch.float64, device=device)
    pb = torch.empty((G, nbins), dtype=torch.float64, device=device) if per_bin else None
    if G == 0:                        # an empty grid shard (more ranks than grid points) contributes nothing
        return out, pb
    with torch.cuda.device(device):
        N.check(lib.b2_das_map(N.ptr(op), M, nbins, N.ptr(f_dev), N.ptr(w_dev), float(c), N.ptr(mic_dev),
                               N.vec3(center_xyz), N.ptr(g_dev), G, steering, int(bool(clamp_negative)),
                               N.ptr(out), N.ptr(pb), mode, N.stream_handle()), "b2_das_map")
    _count(2 if mode in (N.MAP_TC_TF32, N.MAP_TC_TF32X3, N.MAP_TC_MIXED, N.MAP_TC_F16X3) else 1)     # bin-frequency table + map kernel
    return out, pb


def steering_vector(wave_numbers: np.ndarray, mic_xyz: np.ndarray, center_xyz: np.ndarray,
                    grid_xyz: np.ndarray, steering: int):
    """Dense complex128 (F, M, G) steering array on the device (small problems only)."""
    torch = _torch()
    device = torch.device("cuda", torch.cuda.current_device())
    k_dev = torch.from_numpy(np.ascontiguousarray(wave_numbers, dtype=np.float64)).to(device)
    mic_dev = torch.from_numpy(np.ascontiguousarray(mic_xyz, dtype=np.float64)).to(device)
    g_dev = torch.from_numpy(np.ascontiguousarray(grid_xyz, dtype=np.float64)).to(device)
    F, M, G = len(wave_numbers), mic_xyz.shape[0], grid_xyz.shape[0]
    h = torch.empty((F, M, G), dtype=torch.complex128, device=device)
    N.check(N.lib().b2_steering_vector(N.ptr(k_dev), F, N.ptr(mic_dev), M, N.vec3(center_xyz), N.ptr(g_dev), G,
                                       steering, N.ptr(h), N.stream_handle()), "b2_steering_vector")
    return h
