#!/usr/bin/env python
"""Correctness of the N-GPU path on hardware (the `-m gpu` suite is single-GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29533 \
        scripts/check_multigpu.py [--mics 64] [--seconds 4] [--points 60000]

Every rank builds the same seeded recording, keeps only the rows its Welch frames touch
(`dist.ShardedSignal`), and runs `dist.DistributedBeamformerDASFrequency.get_beamformer_map`: frame-sharded
CSM with the bin-range-pipelined NCCL all-reduce of the packed sums, grid-sharded map, all-gather.  Rank 0 then
(a) recomputes CSM and map on its own GPU alone and compares them with the sharded results (the sums are
re-associated: fp64 rounding only), (b) checks seeded grid points and channel pairs against the oracle with
the north-star tolerances.  Prints one JSON line; exit code 0 = all checks passed, 3 = mismatch.
"""

import argparse
import contextlib
import io
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mics", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--wl", type=int, default=2048)
    ap.add_argument("--points", type=int, default=60_000)
    a = ap.parse_args()

    import torch
    import torch.distributed as dist

    import dsptoolbox_b200 as dsp
    from dsptoolbox_b200 import dist as bdist, engine, synth
    from oracle import restatement as R

    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("WORLD_SIZE", "1"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    fs, wl, M = 48_000, a.wl, a.mics
    L = int(a.seconds * fs)
    pos = synth.disc_array(M, 0.75, seed=7)
    mic = synth.positions_to_xyz(pos)
    x = synth.point_source_recording_torch(mic, L, fs, noise_sigma=0.1, seed=1, device=dev)
    step = wl // 2
    K, _ = engine.welch_frames(L, wl, step)
    k0, k1 = bdist.frame_shard(K, rank, world)
    t0, t1 = bdist.shard_sample_range((k0, k1), step, wl, L)
    rng = np.random.default_rng(4)
    g = np.stack([rng.uniform(-1, 1, a.points), rng.uniform(-1, 1, a.points), rng.uniform(0.5, 1.5, a.points)], axis=1)
    grid = dsp.beamforming.Grid(dict(x=g[:, 0], y=g[:, 1], z=g[:, 2]))
    fc, frac = 3000.0, 1.0

    sig = bdist.ShardedSignal(x[t0:t1], t0, L, fs, activate_cache=True).set_spectrum_parameters(window_length_samples=wl)
    bf = bdist.DistributedBeamformerDASFrequency(sig, dsp.beamforming.MicArray(pos), grid, dsp.beamforming.SteeringVector())
    with contextlib.redirect_stdout(io.StringIO()):
        m_sharded = bf.get_beamformer_map(fc, frac, True)
    f, csm_sharded = bdist.get_csm_device_sharded(sig)
    # host-memory rows through the same path (chunked upload on a copy stream + overlapped reduction)
    sig_h = bdist.ShardedSignal(x[t0:t1].cpu().numpy(), t0, L, fs).set_spectrum_parameters(window_length_samples=wl)
    old = engine.config.h2d_chunk_bytes
    engine.config.h2d_chunk_bytes = max(1, (t1 - t0) * M * 8 // 3)
    _, csm_host = bdist.get_csm_device_sharded(sig_h)
    engine.config.h2d_chunk_bytes = old
    torch.cuda.synchronize()
    dist.barrier()
    res = {"world": world}
    ok = True
    if rank == 0:
        full = dsp.Signal(None, x, fs, activate_cache=True).set_spectrum_parameters(window_length_samples=wl)
        bf1 = dsp.beamforming.BeamformerDASFrequency(full, dsp.beamforming.MicArray(pos), grid, dsp.beamforming.SteeringVector())
        with contextlib.redirect_stdout(io.StringIO()):
            m_single = bf1.get_beamformer_map(fc, frac, True)
        _, csm_single = full.get_csm_device()
        res["csm_sharded_vs_single"] = float(((csm_sharded[1:] - csm_single[1:]).abs() / csm_single[1:].abs()).max())
        res["csm_hostrows_vs_single"] = float(((csm_host[1:] - csm_single[1:]).abs() / csm_single[1:].abs()).max())
        res["map_sharded_vs_single"] = float(np.max(np.abs(m_sharded - m_single)) / m_single.max())
        ok &= res["csm_sharded_vs_single"] <= 1e-10 and res["csm_hostrows_vs_single"] <= 1e-10
        ok &= res["map_sharded_vs_single"] <= 1e-6
        # oracle: channel pairs from the recording, grid points from the sharded CSM
        chans = sorted(int(v) for v in rng.choice(M, size=4, replace=False))
        xs = x[:, chans].cpu().numpy()
        c_np = csm_sharded.cpu().numpy()
        worst = 0.0
        for i in range(4):
            for j in range(i, 4):
                want = R.welch_pair(xs[:, i], None if i == j else xs[:, j], fs, "hann", wl, 50, True, "fftbackward")
                got = c_np[:, chans[j], chans[i]]
                worst = max(worst, float(np.max(np.abs(got[1:] - want[1:]) / np.abs(want[1:]))))
        res["csm_vs_oracle_max_rel"] = worst
        pick = np.unique(np.concatenate([rng.choice(a.points, size=40, replace=False), [int(np.argmax(m_sharded))]]))
        ref = R.das_map(f, c_np, mic, g[pick], fc, frac, True, 343.0, "truelocation", chunk_points=16)
        mx = m_sharded.max()
        res["map_vs_oracle_rel_to_max"] = float(np.max(np.abs(m_sharded[pick] - ref)) / mx)
        sel = ref >= mx * 1e-2
        res["map_vs_oracle_db_top20"] = float(np.max(np.abs(10 * np.log10(m_sharded[pick][sel] / ref[sel])))) if sel.any() else 0.0
        ok &= worst <= 1e-5 and res["map_vs_oracle_rel_to_max"] <= 1e-4 and res["map_vs_oracle_db_top20"] <= 0.01
        res["pass"] = bool(ok)
        print(json.dumps(res), flush=True)
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 3)


if __name__ == "__main__":
    main()
