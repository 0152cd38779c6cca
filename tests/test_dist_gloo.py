"""N > 1 host logic on CPU: world_size-2 gloo process group.  Checks the frame/grid sharding index
math (halo, zero-padded tail) and the collectives of dsptoolbox_b200/dist.py; the per-rank compute
is stood in for by the oracle (tests may do that), so the test proves that
shard -> partial sums -> all-reduce -> (non-linear) finalise equals the unsharded result."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

from dsptoolbox_b200 import dist as bdist
from oracle import restatement as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_split_range_properties():
    for n in (0, 1, 7, 94, 235, 1407, 1_000_000):
        for world in (1, 2, 3, 4, 8):
            parts = [bdist.split_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts[:-1], parts[1:]))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
    assert bdist.simulate_frame_sharding(1407, 8)[3] == bdist.frame_shard(1407, 3, 8)


def test_shard_sample_range_halo_and_tail():
    L, wl, step = 48000, 1024, 512
    K, pad = R.frame_count(wl, step, L)
    assert (K, pad) == (94, 640)
    s0 = bdist.shard_sample_range((0, 47), step, wl, L)
    s1 = bdist.shard_sample_range((47, 94), step, wl, L)
    assert s0 == (0, 46 * 512 + 1024)
    assert s1 == (47 * 512, L)                    # last frames run into the zero padding, not stored
    assert s0[1] - s1[0] == wl - step             # halo
    assert bdist.shard_sample_range((5, 5), step, wl, L) == (0, 0)


def _worker(rank, world, port, x, wl, overlap, scaling, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    tdist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        L, M = x.shape
        step = R.welch_step(wl, overlap)
        K, _ = R.frame_count(wl, step, L)
        k0, k1 = bdist.frame_shard(K, rank, world)
        t0, t1 = bdist.shard_sample_range((k0, k1), step, wl, L)
        # this rank only looks at its own samples (+ zero padding for frames past the end)
        need = (k1 - k0 - 1) * step + wl
        xs = np.zeros((need, M))
        xs[: t1 - t0] = x[t0:t1]
        w = R.welch_window("hann", wl)
        fr = xs[np.arange(wl)[:, None] + step * np.arange(k1 - k0)[None, :], :] * w[:, None, None]
        fr -= fr.mean(axis=0)
        X = np.fft.rfft(fr, axis=0, norm=R.fft_norm(scaling))               # (nb, k, M)
        S = torch.from_numpy(np.matmul(np.swapaxes(X, 1, 2), X.conj()))      # un-normalised partial sums
        # the product reduces the PACKED lower triangle bin range by bin range with asynchronous handles
        # (engine.csm_partial_sums + dist.allreduce_sum_async_): same protocol here on the oracle's sums
        rows, cols = torch.tril_indices(M, M)
        Sp = S[:, rows, cols].contiguous()
        nb = Sp.shape[0]
        handles = [bdist.allreduce_sum_async_(Sp[nb * r // 3: nb * (r + 1) // 3]) for r in range(3)]
        for h in handles:
            h.wait()
        S = bdist.allreduce_sum_(S)
        assert torch.equal(Sp, S[:, rows, cols])
        csm = R._finish_welch(S.numpy() / K, scaling, wl, 8000, w)           # non-linear part AFTER the reduce
        # grid sharding + all-gather of unequal pieces
        G = 37
        g0, g1 = bdist.grid_shard(G, rank, world)
        local = torch.arange(g0, g1, dtype=torch.float64) * 1.5
        counts = [b - a for a, b in (bdist.grid_shard(G, r, world) for r in range(world))]
        gathered = bdist.allgather_rows(local, counts)
        if rank == 0:
            out["csm"] = csm
            out["gathered"] = gathered.numpy()
            out["world"] = bdist.world()
    finally:
        tdist.destroy_process_group()


@pytest.mark.parametrize("scaling", ["fftbackward", "powerspectraldensity"])
def test_two_rank_frame_sharding_matches_unsharded(scaling):
    rng = np.random.default_rng(0)
    x = rng.standard_normal((5000, 4)) + 0.7 * rng.standard_normal((5000, 1))
    wl, overlap = 256, 50
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), x, wl, overlap, scaling, out), nprocs=2, join=True)
    _, full = R.csm_welch(x, 8000, wl, "hann", overlap, True, scaling)
    low = np.tril(np.ones((4, 4), bool))
    got = out["csm"]
    assert np.max(np.abs(got[1:][:, low] - full[1:][:, low]) / np.abs(full[1:][:, low])) < 1e-10
    assert np.array_equal(out["gathered"], np.arange(37) * 1.5)
    assert out["world"] == (0, 2)
