#!/usr/bin/env python
"""Benchmark of the CSM + DAS-map hot path (BASELINE.json metric) -- prints ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg45|cfg3|cfg1]

A "step" is one pass of the hot path over one synthetic recording: stage 1 = Welch CSM
(`Signal.get_csm`), stage 2 = delay-and-sum map from that CSM
(`BeamformerDASFrequency.get_beamformer_map`).  Default workload "cfg45" = BASELINE.json
configs[3]+[4]: 256 mics, 60 s @ 48 kHz, nfft 4096 / hann / 50 % (CSM), then 1 000 000 grid points
x 2048 bins (map).  It fits one B200; with N > 1 the SAME total problem is sharded (frames for
stage 1 with a bin-range-pipelined NCCL all-reduce of the packed sums, grid points for stage 2 with
one all-gather): "scaling": "strong".

  value     grid-point*bins/s of stage 2, device-resident inputs, CUDA-event timed, max over ranks
  stages    per-stage times; stages.csm.value = channel-samples/s of stage 1 (the second BASELINE metric)
  e2e       the same step through the public host API: pinned host recording in, host map out
            (H2D of the samples and D2H of the result inside the timed region); e2e_pageable = the
            same with an ordinary (pageable) numpy recording in the stock `dsp.Signal`
  roofline  dominant kernel (the map kernel): algorithmic flops 4*M^2*G*F / its CUDA-event time against
            MEASURED_PEAKS.json's sustained bf16 GEMM rate / tensor products per algorithmic product
            (the same rule at every N); the GEMM / DGEMM rates seen in this run are recorded beside it
  parity    the timed configuration checked after the timed region: seeded grid points of the gathered map
            and channel pairs of the CSM recomputed by the oracle (exit code 3 on failure)
  cpu_baseline  the reference itself (baseline/_ref, unmodified) timed here on a bounded sample,
            plus the vectorised float64 restatement (T1) as the strong CPU baseline
"""

from __future__ import annotations

import os
import sys

if "reference" in sys.argv:
    # `--impl reference` runs on the host cores: torchrun exports OMP_NUM_THREADS=1, which would pin the
    # BLAS behind numpy to one thread at N > 1 and make the CPU arm incomparable between N = 1 and N > 1
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import argparse
import json
import subprocess
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48_000
C_SOUND = 343.0
WORKLOADS = {
    "cfg45": dict(mics=256, radius=1.0, seconds=60, wl=4096, grid=(100, 100, 100), fc=219.089, frac=0.0738,
                  label="BASELINE configs[3]+[4]: 256 mics x 60 s @ 48 kHz, nfft 4096 hann 50% -> CSM; 1e6 grid points x 2048 bins -> DAS map"),
    "cfg3": dict(mics=64, radius=0.75, seconds=10, wl=4096, grid=(200, 200, 20), fc=2000.0, frac=0.25,
                 label="BASELINE configs[1]+[2]: 64 mics x 10 s @ 48 kHz, nfft 4096 -> CSM; 200x200x20 grid x 640 bins (500 Hz-8 kHz) -> DAS map"),
    "cfg1": dict(mics=64, radius=0.75, seconds=1, wl=1024, grid=(50, 50, 1), fc=4000.0, frac=3,
                 label="BASELINE configs[0] shape: 64 mics x 1 s, nfft 1024, 50x50 grid x 20 bins"),
}
MODE_NAMES = {0: "simt_f64", 1: "simt_f32", 2: "tf32", 3: "tf32x3", 4: "mixed", 5: "f16x3"}
# tensor products per algorithmic product (the divisor of the 16-bit GEMM peak): three FP16 products in
# f16x3; split-TF32 = 3 TF32 products = 6 bf16-rate equivalents; mixed = TF32 + double-depth BF16 = 3
PASSES_BF16 = {"f16x3": 3, "tf32x3": 6, "mixed": 3, "tf32": 2}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg45", choices=list(WORKLOADS))
    ap.add_argument("--csm-precision", default=None, choices=["fp64", "fp32"])
    ap.add_argument("--map-mode", default=None)
    ap.add_argument("--map-points", type=int, default=0, help="debug: evaluate only this many grid points")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    return ap.parse_args()


# --------------------------------------------------------------------------- geometry / problem
def problem(args):
    from dsptoolbox_b200 import synth

    w = WORKLOADS[args.workload]
    pos = synth.disc_array(w["mics"], w["radius"], seed=7)
    L = w["seconds"] * FS
    gx, gy, gz = w["grid"]
    lines = (np.linspace(-1.0, 1.0, gx), np.linspace(-1.0, 1.0, gy),
             np.linspace(0.5, 1.5, gz) if gz > 1 else np.array([1.0]))
    return w, pos, L, lines


def band(f, fc, frac):
    lo, hi = (fc, fc) if frac == 0 else (fc * 2 ** (-1 / frac / 2), fc * 2 ** (1 / frac / 2))
    i1, i2 = int(np.argmin(np.abs(lo - f))), int(np.argmin(np.abs(hi - f)))
    return i1, (i2 + 1 if i2 == i1 else i2)


def config_block(args, w, L, mode_name, csm_precision, world):
    """The `config` object: identical in both arms (the driver compares them)."""
    M, wl = w["mics"], w["wl"]
    K = -(-L // (wl // 2))
    nb = wl // 2 + 1
    f = np.fft.rfftfreq(wl, 1 / FS)
    i1, i2 = band(f, w["fc"], w["frac"])
    gx, gy, gz = w["grid"]
    G_full = gx * gy * gz
    G = args.map_points if args.map_points else G_full
    return {"workload": w["label"], "M": M, "L": L, "nfft": wl, "frames": K, "bins": nb, "G": G, "F": i2 - i1,
            "G_full": G_full, "map_mode": mode_name, "csm_precision": csm_precision,
            "parallelism": (f"frames/{world} + pipelined all-reduce of the packed sums, grid/{world} + all-gather"
                            if world > 1 else "single GPU"),
            "l2": "inputs larger than L2 (recording %.2f GB, spectra %.2f GB, CSM %.2f GB)"
                  % (M * L * 8 / 1e9, (16 if csm_precision == "fp64" else 8) * M * K * nb / 1e9, nb * M * M * 16 / 1e9)}


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        # "under load" = samples in the upper half of the observed clock range
        load = [c for c in sm if c >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------- the CPU arms
class CpuArms:
    """The reference's own CPU implementation of the path on a bounded sample of the workload.

    kind "reference": the UNMODIFIED dsptoolbox v0.8 from baseline/_ref (oracle/install_reference.py;
    I/O and plotting imports stubbed by oracle/ref_loader.py) -- `_welch` per channel pair
    (standard/_spectral_methods.py:10-173, the body of `_csm_welch`'s double loop :351-369) and
    `BeamformerDASFrequency.get_beamformer_map` (beamforming/beamforming.py:805-880) on a plain `Grid` of
    sampled grid points with the CSM served from the Signal's cache.  kind "port": the oracle's
    restatement with the same loop structure, only where the reference package is not importable.
    The reference's cost is exactly linear in the sampled dimension (one `_welch` per channel pair, one
    multi_dot per grid point and bin), so rates measured on the sample are the rates of the full problem.
    """

    def __init__(self, w, pos, L, lines, channels=8):
        from dsptoolbox_b200 import synth
        from oracle import ref_loader

        self.w, self.L, self.lines = w, L, lines
        self.mic = synth.positions_to_xyz(pos)
        self.pos = pos
        self.M = self.mic.shape[0]
        rng = np.random.default_rng(0)
        self.sel = sorted(int(v) for v in rng.choice(self.M, size=min(self.M, channels), replace=False))
        self.x = synth.point_source_recording(self.mic[self.sel], L, FS, noise_sigma=0.1, seed=1)
        self.f = np.fft.rfftfreq(w["wl"], 1 / FS)
        i1, i2 = band(self.f, w["fc"], w["frac"])
        self.F = i2 - i1
        self.G = len(lines[0]) * len(lines[1]) * len(lines[2])
        # any Hermitian matrix of the right shape costs the same in the map loop
        a = rng.standard_normal((len(self.f), self.M, 4)) + 1j * rng.standard_normal((len(self.f), self.M, 4))
        self.csm = a @ np.conj(np.swapaxes(a, 1, 2))
        n = len(self.sel)
        self.pairs = [(i, j) for i in range(n) for j in range(i + 1, n)] + [(i, i) for i in range(n)]
        self.pair_cursor = 0
        self.ref = None
        if ref_loader.available():
            try:
                self.ref = ref_loader.load()
            except Exception as e:                                   # noqa: BLE001
                print(f"[bench] reference import failed ({e}); using the oracle port", file=sys.stderr)
        self.kind = "reference" if self.ref is not None else "port"
        if self.ref is not None:
            ref = self.ref
            sig = ref.Signal(None, np.zeros((self.w["wl"], self.M)), FS, activate_cache=True)
            sig.set_spectrum_parameters(window_length_samples=w["wl"])
            # serve the CSM from the Signal's cache (signal.py:973-977, 1003-1007): the map timing below is
            # the reference's own get_beamformer_map from its `self.signal.get_csm()` call onwards
            sig.csm = [self.f.copy(), self.csm]
            sig._Signal__csm_state_update = False
            self.ref_sig = sig

    def _points(self, n, seed):
        rng = np.random.default_rng(1000 + seed)
        gx, gy, gz = self.lines
        gi = rng.choice(self.G, size=min(n, self.G), replace=False)
        ix, iy, iz = np.unravel_index(gi, (len(gx), len(gy), len(gz)))
        return np.stack([gx[ix], gy[iy], gz[iz]], axis=1)

    def csm_sample(self, n_pairs):
        """Time `n_pairs` `_welch` calls on the full-length recording -> (seconds, cross-equivalents)."""
        w = self.w
        todo = [self.pairs[(self.pair_cursor + k) % len(self.pairs)] for k in range(n_pairs)]
        self.pair_cursor += n_pairs
        t0 = time.perf_counter()
        if self.ref is not None:
            from dsptoolbox.standard._spectral_methods import _welch

            for i, j in todo:
                _welch(self.x[:, i], None if i == j else self.x[:, j], FS, self.ref.Window.Hann, w["wl"], 50, True, "mean",
                       self.ref.SpectrumScaling.FFTBackward)
        else:
            from oracle import restatement as R

            R.csm_welch_pairwise(self.x, FS, w["wl"], "hann", 50, True, "fftbackward", pairs=todo)
        dt = time.perf_counter() - t0
        # an auto-spectrum call transforms one channel instead of two: count it as half a cross call
        return dt, sum(0.5 if i == j else 1.0 for i, j in todo)

    def map_sample(self, n_points, seed):
        g = self._points(n_points, seed)
        t0 = time.perf_counter()
        if self.ref is not None:
            import contextlib
            import io

            bfm = self.ref.beamforming
            bf = bfm.BeamformerDASFrequency(self.ref_sig, bfm.MicArray(self.pos),
                                            bfm.Grid(dict(x=g[:, 0], y=g[:, 1], z=g[:, 2])), bfm.SteeringVector())
            with contextlib.redirect_stdout(io.StringIO()):
                bf.get_beamformer_map(self.w["fc"], self.w["frac"], True)
        else:
            from oracle import restatement as R

            R.das_map_loop(self.f, self.csm, self.mic, g, self.w["fc"], self.w["frac"], True)
        return time.perf_counter() - t0, len(g)

    def map_marginal(self, n_small, n_large, seed):
        """Two map calls on n_small and n_large sampled points: every `get_beamformer_map` call pays a fixed
        cost (copy of the cached CSM, diagonal removal over all bins, ...: ~1 s at M = 256) that the full-size
        job pays once per 10^6 points, so the rate of the full problem is the MARGINAL one,
        (n_large - n_small) * F / (t_large - t_small).  Returns (seconds of both calls, marginal seconds,
        marginal points, fixed seconds)."""
        t1, n1 = self.map_sample(n_small, 2 * seed)
        t2, n2 = self.map_sample(n_large, 2 * seed + 1)
        per_point = max(t2 - t1, 1e-9) / max(1, n2 - n1)
        return t1 + t2, t2 - t1, n2 - n1, max(0.0, t1 - n1 * per_point)

    def step(self, pairs, points, seed=0):
        t_c, eq = self.csm_sample(pairs)
        t_both, t_m, npts, t_fixed = self.map_marginal(max(2, points // 8), points + max(2, points // 8), seed)
        return dict(t_csm=t_c, cross_equiv=eq, t_map=t_m, points=npts, t_map_calls=t_both, t_map_fixed=t_fixed)

    def vectorised(self, points=256, bins=64, channels=16):
        """T1: the oracle's vectorised float64 restatement on all host cores (BLAS), sampled along linear
        dimensions: map = `points` grid points x all bins; CSM = spectra of `channels` channels (linear in M)
        + one zgemm per bin for `bins` bins at the full M and K (linear in bins)."""
        from oracle import restatement as R

        w, M, L = self.w, self.M, self.L
        g = self._points(points, 77)
        t0 = time.perf_counter()
        R.das_map(self.f, self.csm, self.mic, g, w["fc"], w["frac"], True, chunk_points=64)
        t_map = time.perf_counter() - t0
        nch = min(channels, self.x.shape[1])
        t0 = time.perf_counter()
        X = R.channel_spectra(self.x[:, :nch], w["wl"], "hann", 50, True, "fftbackward")       # (nb, K, nch)
        t_fft = (time.perf_counter() - t0) * M / nch
        K = X.shape[1]
        rng = np.random.default_rng(5)
        Xb = rng.standard_normal((bins, K, M)) + 1j * rng.standard_normal((bins, K, M))
        t0 = time.perf_counter()
        np.matmul(np.swapaxes(Xb, 1, 2), Xb.conj())
        t_gemm = (time.perf_counter() - t0) * len(self.f) / bins
        return {"map_value": len(g) * self.F / t_map, "map_unit": "grid-pt*bins/s",
                "csm_value": M * L / (t_fft + t_gemm), "csm_unit": "ch-samples/s",
                "sample": f"map: {len(g)} grid points x {self.F} bins; CSM: spectra of {nch} channels scaled to {M} "
                          f"+ zgemm of {bins} of {len(self.f)} bins at full M={M}, K={K}"}


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation on the host cores, on this arm's config,
    each step a bounded sample of the workload (rank 0 only; other ranks exit 0 without work)."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w, pos, L, lines = problem(args)
    cores = os.cpu_count()
    arm = CpuArms(w, pos, L, lines, channels=12)
    big = w["mics"] >= 128
    pairs, points = (2, 32) if big else (8, 32)
    per = []
    t_steps = []
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = arm.step(pairs=pairs, points=points, seed=it)
        if it >= args.warmup:
            per.append(r); t_steps.append(time.perf_counter() - t0)
    M = w["mics"]
    G, F = arm.G, arm.F
    # rates over the union of the timed steps' samples
    cross_full = M * (M - 1) / 2 + 0.5 * M
    t_c, eq = sum(r["t_csm"] for r in per), sum(r["cross_equiv"] for r in per)
    t_m, npts = sum(r["t_map"] for r in per), sum(r["points"] for r in per)
    csm_rate = M * L / (t_c / eq * cross_full)
    map_rate = npts * F / t_m
    full_step_s = G * F / map_rate + M * L / csm_rate
    mode_name = args.map_mode or "f16x3"
    sample = (f"per step: {pairs} `_welch` channel pairs of the full {L}-sample recording + two get_beamformer_map calls on "
              f"{max(2, points // 8)} and {points + max(2, points // 8)} of {G} grid points x all {F} bins; {args.steps} timed steps = "
              f"{len(per) * pairs} pairs and {npts} marginal grid points in total; rates = sampled units / their time, the map "
              f"rate being the marginal one (t_large - t_small: the per-call fixed cost of "
              f"{float(np.mean([r['t_map_fixed'] for r in per])):.2f} s is paid once per 10^6 points at full size); "
              f"the full-size step would take {full_step_s:.0f} s")
    line = {
        "impl": "reference", "metric": "das_map_grid_point_bins_per_s", "value": map_rate, "unit": "grid-pt*bins/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(t_steps)) * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args, w, L, mode_name, args.csm_precision or "fp64", max(1, args.gpus)),
        "stages": {"csm": {"value": csm_rate, "unit": "ch-samples/s"}, "das_map": {"value": map_rate, "unit": "grid-pt*bins/s"}},
        "cpu_baseline": {"value": map_rate, "unit": "grid-pt*bins/s", "cores": cores, "kind": arm.kind, "sample": sample,
                         "csm_value": csm_rate, "csm_unit": "ch-samples/s", "full_size_step_s": full_step_s,
                         "threads": threads_info()},
        # whole step (CSM + map) of the full problem at the sampled rates, in the map metric: the same
        # quantity as the GPU arm's e2e (no device, hence no copies)
        "e2e": {"value": G * F / full_step_s, "unit": "grid-pt*bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def threads_info():
    try:
        from threadpoolctl import threadpool_info

        return [{"api": i.get("internal_api"), "threads": i.get("num_threads")} for i in threadpool_info()]
    except Exception:
        return None


# --------------------------------------------------------------------------- B200 arm
def gemm_rates(torch, sustained_seconds=1.5):
    """cuBLAS rates of this GPU in this run, recorded beside the roofline (the denominators themselves come
    from MEASURED_PEAKS.json): FP16 8192^3 burst (best of 5) + sustained loop, TF32 8192^3 burst, DGEMM 4096^3."""
    out = {}

    def timed(fn, n=1):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n

    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        a = torch.randn(8192, 8192, device="cuda").half(); b = torch.randn(8192, 8192, device="cuda").half()
        timed(lambda: a @ b, 2)
        best = min(timed(lambda: a @ b) for _ in range(5))
        out["fp16_burst_tflops"] = 2 * 8192 ** 3 / best / 1e9
        n = max(8, int(sustained_seconds * 1e3 / best))
        out["fp16_sustained_tflops"] = 2 * 8192 ** 3 / timed(lambda: a @ b, n) / 1e9
        torch.backends.cuda.matmul.allow_tf32 = True
        a, b = a.float(), b.float()
        timed(lambda: a @ b, 2)
        out["tf32_burst_tflops"] = 2 * 8192 ** 3 / min(timed(lambda: a @ b) for _ in range(3)) / 1e9
        a = torch.randn(4096, 4096, device="cuda", dtype=torch.float64); b = torch.randn(4096, 4096, device="cuda", dtype=torch.float64)
        timed(lambda: a @ b, 2)
        out["fp64_dgemm_tflops"] = 2 * 4096 ** 3 / min(timed(lambda: a @ b) for _ in range(3)) / 1e9
        del a, b
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    return out


def parity_check(torch, R, w, mic, x_dev, f, csm_dev, map_gathered, grid_xyz, n_points=40, n_channels=5):
    """The configuration that was timed, checked against the oracle (north-star tolerances, tests/parity.py):
    `n_points` seeded grid points (+ the map's arg-max) of the full-size gathered map recomputed from the
    device CSM, and the n(n+1)/2 entries of `n_channels` seeded channels recomputed from the recording."""
    M = mic.shape[0]
    rng = np.random.default_rng(12345)
    m = map_gathered.detach().cpu().numpy()
    G = m.shape[0]
    idx = np.unique(np.concatenate([rng.choice(G, size=min(n_points, G), replace=False), [int(np.argmax(m))]]))
    csm = csm_dev.cpu().numpy()
    ref = R.das_map(f, csm, mic, grid_xyz[idx], w["fc"], w["frac"], True, C_SOUND, "truelocation", chunk_points=8)
    mx = float(m.max())
    map_rel = float(np.max(np.abs(m[idx] - ref)) / mx)
    top = ref >= mx * 1e-2
    map_db = float(np.max(np.abs(10 * np.log10(m[idx][top] / ref[top])))) if top.any() else 0.0
    chans = sorted(int(v) for v in rng.choice(M, size=min(n_channels, M), replace=False))
    xs = x_dev[:, chans].cpu().numpy()
    worst = 0.0
    pairs = 0
    for a in range(len(chans)):
        for b in range(a, len(chans)):
            want = R.welch_pair(xs[:, a], None if a == b else xs[:, b], FS, "hann", w["wl"], 50, True, "fftbackward")
            got = csm[:, chans[b], chans[a]]
            worst = max(worst, float(np.max(np.abs(got[1:] - want[1:]) / np.abs(want[1:]))))
            pairs += 1
    ok = bool(map_rel <= 1e-4 and map_db <= 0.01 and worst <= 1e-5 and np.all(np.isfinite(m)) and np.all(m >= 0))
    return {"pass": ok, "csm_max_rel": worst, "csm_pairs": pairs, "csm_tol": 1e-5, "map_rel_to_max": map_rel,
            "map_db_top20": map_db, "map_tol": [1e-4, 0.01], "points": int(len(idx)), "points_in_top20db": int(top.sum()),
            "checker": "oracle/restatement.py das_map (from the device CSM) and welch_pair (from the recording)"}


def run_b200(args):
    import torch
    import torch.distributed as dist

    import dsptoolbox_b200 as dsp
    from dsptoolbox_b200 import dist as bdist, engine, synth, _native as N

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    assert world == args.gpus or world == 1, "launch with torchrun --nproc-per-node N for --gpus N"
    if args.csm_precision:
        engine.config.csm_precision = args.csm_precision
    if args.map_mode:
        engine.config.map_mode = args.map_mode
    mode = engine.resolve_map_mode()
    mode_name = MODE_NAMES[mode]

    w, pos, L, lines = problem(args)
    M, wl = w["mics"], w["wl"]
    mic = synth.positions_to_xyz(pos)
    step_samples = wl // 2
    K, _ = engine.welch_frames(L, wl, step_samples)
    nb = wl // 2 + 1
    f = np.fft.rfftfreq(wl, 1 / FS)
    i1, i2 = band(f, w["fc"], w["frac"])
    F = i2 - i1
    grid = dsp.beamforming.Regular3DGrid(*lines)
    G_full = grid.number_of_points
    if args.map_points:
        grid = dsp.beamforming.Grid({k: grid._coordinates[: args.map_points, i] for i, k in enumerate("xyz")})
    G = grid.number_of_points

    # ---- synthetic recording, generated on the device; this rank's shard also in pinned host memory
    x_dev = synth.point_source_recording_torch(mic, L, FS, noise_sigma=0.1, seed=1, device=dev)
    k0, k1 = bdist.frame_shard(K, rank, world)
    t0s, t1s = bdist.shard_sample_range((k0, k1), step_samples, wl, L)
    x_host = torch.empty((t1s - t0s, M), dtype=torch.float64, pin_memory=True)
    x_host.copy_(x_dev[t0s:t1s])
    torch.cuda.synchronize()
    x_np = x_host.numpy()

    def make_sig(rows):
        """The public Signal: stock `dsp.Signal` on one GPU, `dist.ShardedSignal` (this rank's rows) on N."""
        if world == 1:
            s = dsp.Signal(None, rows, FS)
        else:
            s = bdist.ShardedSignal(rows, t0s, L, FS)
        return s.set_spectrum_parameters(window_length_samples=wl)

    BFClass = bdist.DistributedBeamformerDASFrequency if world > 1 else dsp.beamforming.BeamformerDASFrequency

    def make_bf(sig):
        return BFClass(sig, dsp.beamforming.MicArray(pos), grid, dsp.beamforming.SteeringVector())

    sig_res = make_sig(x_dev[t0s:t1s])        # device-resident inputs
    sig_e2e = make_sig(x_np)                  # pinned host inputs
    bf_res, bf_e2e = make_bf(sig_res), make_bf(sig_e2e)
    devnull = open(os.devnull, "w")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def csm_resident():
        return bf_res._csm_device()

    def map_resident(fq, csm):
        """Stage 2 from a device CSM; returns the gathered flat map as a CUDA tensor."""
        bf_res.f_range_hz = (np.array([w["fc"] * 2 ** (-1 / w["frac"] / 2), w["fc"] * 2 ** (1 / w["frac"] / 2)])
                             if w["frac"] else np.array([w["fc"]] * 2))
        g0, g1 = bf_res._grid_rows()
        id1, id2 = bf_res._select_bins(fq)
        weights = dsp.beamforming.simpson_weights(id2 - id1, fq[1] - fq[0]) if id2 - id1 > 1 else np.ones(1)
        op = engine.das_prepare(csm, id1, id2 - id1, True, mode)
        out, _ = engine.das_map(op, M, fq[id1:id2], weights, C_SOUND, mic, bf_res.mics.array_center_coordinates,
                                grid._device_coordinates(dev)[g0:g1], N.STEER_TRUE_LOCATION, True, mode)
        if world > 1:
            counts = [b - a for a, b in (bdist.grid_shard(G, r, world) for r in range(world))]
            out = bdist.allgather_rows(out, counts)
        return out

    def step_resident():
        """Stage 1 then stage 2, device-resident input; returns CUDA events around each stage."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        fq, csm = csm_resident()
        ev[1].record()
        out = map_resident(fq, csm)
        ev[2].record()
        return ev, out, fq, csm

    def step_e2e(bf=None):
        so = sys.stdout
        sys.stdout = devnull
        try:
            return (bf or bf_e2e).get_beamformer_map(w["fc"], w["frac"], True)
        finally:
            sys.stdout = so

    # ---- warm-up (>= 3), then K timed steps bracketed by barrier + synchronize
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    if os.environ.get("B2_BENCH_TRACE") == "1":
        engine.config.profile = True
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = engine.launches
    t_csm, t_map = [], []
    wall0 = time.perf_counter()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    evs = []
    prof = None
    if os.environ.get("B2_BENCH_TRACE") == "2":  # debug: where the HOST spends its time inside the timed loop
        import cProfile
        prof = cProfile.Profile()
        prof.enable()
    for _ in range(args.steps):
        ev, out, fq_last, csm_last = step_resident()
        evs.append(ev)
        if os.environ.get("B2_BENCH_TRACE") == "1":     # debug: per-kernel times of stage 1 inside the timed loop (adds syncs)
            print("trace", engine.config.stats["csm"].get("times_ms"), file=sys.stderr)
    e_stop.record()
    if prof is not None:
        import pstats
        prof.disable()
        pstats.Stats(prof, stream=sys.stderr).sort_stats("tottime").print_stats(18)
    barrier()
    wall = time.perf_counter() - wall0
    launches = engine.launches - launches0
    total_ms = e_start.elapsed_time(e_stop)
    for ev in evs:
        t_csm.append(ev[0].elapsed_time(ev[1])); t_map.append(ev[1].elapsed_time(ev[2]))
    clocks = sampler.stop()

    def maxr(v):
        if world == 1:
            return float(v)
        t = torch.tensor([v], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_step = maxr(total_ms / args.steps)
    ms_csm = maxr(float(np.mean(t_csm)))
    ms_map = maxr(float(np.mean(t_map)))

    # ---- parity of the configuration that was just timed (rank 0 holds the gathered map and the full CSM)
    parity = None
    if not args.no_parity and rank == 0:
        from oracle import restatement as R

        parity = parity_check(torch, R, w, mic, x_dev, fq_last, csm_last, out, grid._coordinates)
    del csm_last

    # ---- kernel-level timing of the dominant kernel (map) and of the stage-1 kernels, one more step
    engine.config.profile = True
    csm_resident()
    k_times = dict(engine.config.stats["csm"]["times_ms"])
    engine.config.profile = False
    fq, csm = csm_resident()
    id1, id2 = band(fq, w["fc"], w["frac"])
    weights = dsp.beamforming.simpson_weights(id2 - id1, fq[1] - fq[0]) if id2 - id1 > 1 else np.ones(1)
    g0, g1 = bf_res._grid_rows()
    op = engine.das_prepare(csm, id1, id2 - id1, True, mode)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    engine.das_map(op, M, fq[id1:id2], weights, C_SOUND, mic, bf_res.mics.array_center_coordinates,
                   grid._device_coordinates(dev)[g0:g1], N.STEER_TRUE_LOCATION, True, mode)
    b.record(); torch.cuda.synchronize()
    ms_map_kernel = maxr(a.elapsed_time(b))
    del op, csm

    # ---- the all-reduce of the packed sums alone (NCCL bus bandwidth at this message size)
    nccl = None
    if world > 1:
        S = torch.zeros((nb, engine.packed_size(M)), dtype=torch.complex128, device=dev)
        for _ in range(2):
            bdist.allreduce_sum_(S)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            bdist.allreduce_sum_(S)
        b.record(); torch.cuda.synchronize()
        ms_ar = maxr(a.elapsed_time(b) / 3)
        nbytes = S.numel() * 16
        nccl = {"allreduce_alone_ms": ms_ar, "bytes": nbytes,
                "busbw_gbs": nbytes * 2 * (world - 1) / world / (ms_ar * 1e-3) / 1e9}
        del S

    # ---- end-to-end through the public host API
    e2e = None
    e2e_pageable = None
    if not args.no_e2e:
        m_e2e = step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            m_e2e = step_e2e()
        barrier()
        t_e2e = maxr((time.perf_counter() - t0) / n_e2e)
        h2d = x_np.nbytes
        if world > 1:
            tb = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
            dist.all_reduce(tb)
            h2d = int(tb.item())
        same = float(np.max(np.abs(m_e2e.ravel() - out.cpu().numpy())) / out.max().item())
        e2e = {"value": G * F / t_e2e, "unit": "grid-pt*bins/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(G * 8),
               "api": ("BeamformerDASFrequency.get_beamformer_map(fc, frac, True) on dsp.Signal(None, ndarray, fs)"
                       if world == 1 else
                       "dist.DistributedBeamformerDASFrequency.get_beamformer_map on dist.ShardedSignal (this rank's rows)")
                      + ", float64 samples in pinned host memory; grid / microphone uploads are cached on their objects",
               "csm_value": M * L / t_e2e, "csm_unit": "ch-samples/s (whole step)",
               "max_diff_to_resident_map_rel_to_max": same}
        if world == 1:
            # the same call with an ordinary (pageable) numpy array: the upload is staged by the driver
            x_page = np.array(x_np)
            bf_p = make_bf(make_sig(x_page))
            step_e2e(bf_p)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step_e2e(bf_p)
            torch.cuda.synchronize()
            t_p = time.perf_counter() - t0
            e2e_pageable = {"value": G * F / t_p, "unit": "grid-pt*bins/s", "ms_per_step": t_p * 1e3,
                            "h2d_bytes_per_step": int(x_page.nbytes)}
            del x_page, bf_p

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    src = "measured" if peaks else "fallback"
    hbm = peaks.get("hbm_gbs", 6650.0)
    # the step runs for seconds under the 1 kW cap: sustained figure at every N (fallback: 1400 sustained)
    bf16_sustained = peaks.get("bf16_tflops_sustained", 1400.0)
    in_run = gemm_rates(torch)
    passes = PASSES_BF16.get(mode_name, 1)
    Gr = g1 - g0
    alg_flops = 4.0 * M * M * Gr * F
    achieved = alg_flops / (ms_map_kernel * 1e-3) / 1e12
    tensor_peak = bf16_sustained / passes
    # DRAM traffic of one launch of this kernel from committed ncu captures (same workload, mode, points per GPU)
    traffic = None
    try:
        cap = json.load(open(os.path.join(ROOT, "profiles", "map_kernel_traffic.json")))
        key = f"{args.workload}:{mode_name}:G{Gr}"
        if key in cap:
            traffic = cap[key]["dram_bytes_read"] + cap[key]["dram_bytes_write"]
    except Exception:
        pass
    roofline = {"kernel": f"das_map[{mode_name}]", "bound": "tensor", "achieved": achieved, "peak": tensor_peak,
                "unit": "TFLOP/s", "frac": achieved / tensor_peak, "traffic": traffic,
                "algorithmic_flops_per_launch": alg_flops, "kernel_ms": ms_map_kernel,
                "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained {bf16_sustained} (of {src}) / {passes} tensor "
                               f"products per algorithmic product ({mode_name}); same rule at every N",
                "in_run_cublas": in_run}
    s_in, s_x = (8, 16) if engine.config.csm_precision == "fp64" else (4, 8)
    Kloc = k1 - k0
    fft_bytes = s_in * M * (t1s - t0s) + s_x * M * Kloc * nb
    herk_flops = 4.0 * M * M * nb * Kloc
    stage1 = {}
    if "k1_rfft" in k_times:
        stage1["welch_rfft"] = {"bound": "hbm", "ms": k_times["k1_rfft"], "achieved": fft_bytes / k_times["k1_rfft"] / 1e6,
                                "peak": hbm, "unit": "GB/s", "frac": fft_bytes / k_times["k1_rfft"] / 1e6 / hbm,
                                "peak_source": f"MEASURED_PEAKS.json hbm_gbs (of {src})"}
    if "k2_herk" in k_times:
        ach = herk_flops / k_times["k2_herk"] / 1e9
        stage1["csm_herk"] = {"bound": "tensor(fp64 DMMA)", "ms": k_times["k2_herk"], "achieved": ach,
                              "peak": in_run.get("fp64_dgemm_tflops"), "unit": "TFLOP/s",
                              "frac": ach / in_run["fp64_dgemm_tflops"] if in_run.get("fp64_dgemm_tflops") else None,
                              "algorithmic_flops": herk_flops,
                              "peak_source": "cuBLAS DGEMM 4096^3 in this run (MEASURED_PEAKS.json has no fp64 figure)"}
    if "k0_prepare" in k_times:
        b0 = (8 + s_in) * M * (t1s - t0s)
        stage1["prepare_channels"] = {"bound": "hbm", "ms": k_times["k0_prepare"], "achieved": b0 / k_times["k0_prepare"] / 1e6,
                                      "peak": hbm, "unit": "GB/s", "frac": b0 / k_times["k0_prepare"] / 1e6 / hbm}

    cpu = None
    if not args.no_cpu_baseline and world == 1:       # reported at N = 1 only
        arm = CpuArms(w, pos, L, lines, channels=12)
        big = M >= 128
        npairs = 16 if big else 48
        t_c, eq = arm.csm_sample(npairs)
        arm.map_sample(2, 99)                                        # first call touches the cached CSM's pages
        t_both, t_m, npts, t_fixed = arm.map_marginal(8, 104 if big else 264, 0)
        cross_full = M * (M - 1) / 2 + 0.5 * M
        cpu = {"value": npts * arm.F / t_m, "unit": "grid-pt*bins/s", "cores": os.cpu_count(), "kind": arm.kind,
               "sample": f"stage 1: {npairs} `_welch` calls on channel pairs of the full {L}-sample recording scaled to "
                         f"M(M+1)/2={M*(M+1)//2} pairs; stage 2: marginal rate between two get_beamformer_map calls on 8 and "
                         f"{npts + 8} of {arm.G} grid points x all {arm.F} bins (fixed cost per call {t_fixed:.2f} s excluded: "
                         f"it is paid once per full-size map)",
               "csm_value": M * L / (t_c / eq * cross_full), "csm_unit": "ch-samples/s", "sample_seconds": t_c + t_both,
               "t1_vectorised": arm.vectorised(), "threads": threads_info()}

    line = {
        "metric": "das_map_grid_point_bins_per_s", "value": G * F / (ms_map * 1e-3), "unit": "grid-pt*bins/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {3: "tf32x3 (fp32 accumulate) map; f64 CSM", 4: "tf32+bf16x2 (fp32 accumulate) map; f64 CSM",
                  5: "f16x3 (two FP16 halves per operand, 3 products, fp32 accumulate) map; f64 CSM"}.get(
            mode, f"{mode_name} map; {engine.config.csm_precision} CSM"),
        "data": "synthetic",
        "config": config_block(args, w, L, mode_name, engine.config.csm_precision, world),
        "stages": {"csm": {"value": M * L / (ms_csm * 1e-3), "unit": "ch-samples/s", "ms": ms_csm, "kernels_ms": k_times,
                           "rooflines": stage1, "nccl": nccl},
                   "das_map": {"value": G * F / (ms_map * 1e-3), "unit": "grid-pt*bins/s", "ms": ms_map}},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_pageable": e2e_pageable, "parity": parity,
        "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()
    if parity is not None and not parity["pass"]:
        print(f"[bench] PARITY FAILED: {parity}", file=sys.stderr)
        sys.exit(3)


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the process's original stdout; everything else any library
    prints on fd 1 (e.g. NCCL's version banner) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, data)


def main():
    global _REAL_STDOUT
    args = parse()
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
