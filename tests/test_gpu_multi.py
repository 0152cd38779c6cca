"""Multi-device behaviour (needs >= 2 GPUs; skipped on a single-GPU box -- run scripts/check_multigpu.py under
torchrun there is no substitute for): (a) the C ABI driven on two devices from ONE process (per-device kernel
attributes / SM counts, include/b2dsp.h "re-entrant across devices"), (b) the torchrun-launched 2-rank check."""

import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
need2 = pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs")


@need2
def test_two_devices_from_one_process():
    import dsptoolbox_b200 as dsp
    from dsptoolbox_b200 import engine, synth
    from oracle import restatement as R

    mic = synth.positions_to_xyz(synth.disc_array(40, 0.5, seed=3))
    x = synth.point_source_recording(mic, 30_000, 16_000, noise_sigma=0.1, seed=5)
    _, want = R.csm_welch(x, 16_000, 1024, "hann", 50, True, "fftbackward")
    w = R.welch_window("hann", 1024)
    g = np.random.default_rng(0).uniform(-0.5, 0.5, (20_000, 3)) + np.array([0, 0, 1.0])
    maps = []
    for d in (1, 0, 1):                      # device 1 FIRST: it must not inherit device 0's setup
        with torch.cuda.device(d):
            S, K = engine.csm_partial_sums(x, 30_000, 1024, 512, w, True, 1.0)
            c = engine.csm_finalize(S, 40, K, 1.0, False, True)
            assert c.device.index == d
            got = c.cpu().numpy()
            assert np.max(np.abs(got[1:] - want[1:]) / np.abs(want[1:])) <= 1e-9
            mode = engine.resolve_map_mode("f16x3")
            op = engine.das_prepare(c, 60, 40, True, mode)
            f = np.fft.rfftfreq(1024, 1 / 16_000)[60:100]
            out, _ = engine.das_map(op, 40, f, np.ones(40), 343.0, mic, mic[0], g, 3, True, mode)
            assert out.device.index == d
            maps.append(out.cpu().numpy())
            lam, vec = engine.hermitian_eig(c[60:64])                      # shared-memory opt-in is per device, too
            assert lam.device.index == d
            assert np.max(np.abs(lam.cpu().numpy() - np.linalg.eigvalsh(got[60:64]))) <= 1e-12 * np.abs(got[60:64]).max() * 40
    assert np.max(np.abs(maps[0] - maps[1])) <= 1e-12 * maps[0].max()
    assert np.max(np.abs(maps[0] - maps[2])) <= 1e-12 * maps[0].max()


@need2
def test_two_rank_torchrun_check():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "scripts", "check_multigpu.py"), "--points", "40000"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"pass": true' in r.stdout
