#!/usr/bin/env python
"""Benchmark of the CSM + DAS-map hot path (BASELINE.json metric) -- prints ONE JSON line.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg45|cfg3|cfg1]

A "step" is one pass of the hot path over one synthetic recording: stage 1 = Welch CSM
(`Signal.get_csm`), stage 2 = delay-and-sum map from that CSM
(`BeamformerDASFrequency.get_beamformer_map`).  Default workload "cfg45" = BASELINE.json
configs[3]+[4]: 256 mics, 60 s @ 48 kHz, nfft 4096 / hann / 50 % (CSM), then 1 000 000 grid points
x 2048 bins (map).  It fits one B200; with N > 1 the SAME total problem is sharded (frames for
stage 1 with one NCCL all-reduce, grid points for stage 2 with one all-gather): "scaling": "strong".

  value     grid-point*bins/s of stage 2, device-resident inputs, CUDA-event timed, max over ranks
  stages    per-stage times; stages.csm.value = channel-samples/s of stage 1 (the second BASELINE metric)
  e2e       the same step through the public host API: pinned host recording in, host map out
            (H2D of the samples and D2H of the result inside the timed region)
  roofline  dominant kernel (the map kernel): algorithmic flops 4*M^2*G*F / its CUDA-event time,
            against the cuBLAS GEMM peak measured in this run in the mode's arithmetic
            (FP16 / 3 for the default FP16x3 mode, TF32 / 3 split-TF32, TF32 / 2 mixed)
  cpu_baseline  the oracle port with the reference's loop structure, timed here on a bounded sample
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FS = 48_000
WORKLOADS = {
    # name: (cfg for stage 1, cfg for stage 2)
    "cfg45": dict(mics=256, radius=1.0, seconds=60, wl=4096, grid=(100, 100, 100), fc=219.089, frac=0.0738,
                  label="BASELINE configs[3]+[4]: 256 mics x 60 s @ 48 kHz, nfft 4096 hann 50% -> CSM; 1e6 grid points x 2048 bins -> DAS map"),
    "cfg3": dict(mics=64, radius=0.75, seconds=10, wl=4096, grid=(200, 200, 20), fc=2000.0, frac=0.25,
                 label="BASELINE configs[1]+[2]: 64 mics x 10 s @ 48 kHz, nfft 4096 -> CSM; 200x200x20 grid x 640 bins (500 Hz-8 kHz) -> DAS map"),
    "cfg1": dict(mics=64, radius=0.75, seconds=1, wl=1024, grid=(50, 50, 1), fc=4000.0, frac=3,
                 label="BASELINE configs[0] shape: 64 mics x 1 s, nfft 1024, 50x50 grid x 20 bins"),
}


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


# --------------------------------------------------------------------------- reference arm (CPU)
class CpuReference:
    """The oracle port that keeps the reference's loop structure, timed on a bounded sample and
    scaled linearly (the reference's cost is exactly linear in the sampled dimension: one `_welch`
    per channel pair, one multi_dot per grid point and bin)."""

    def __init__(self, w, pos, L, lines, channels=6):
        from dsptoolbox_b200 import synth

        self.w, self.L, self.lines = w, L, lines
        self.mic = synth.positions_to_xyz(pos)
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

    def step(self, pairs, points, seed=0):
        from oracle import restatement as R

        w, M, L = self.w, self.M, self.L
        n = len(self.sel)
        plist = [(i, j) for i in range(n) for j in range(i, n)][:pairs]
        t0 = time.perf_counter()
        R.csm_welch_pairwise(self.x, FS, w["wl"], "hann", 50, True, "fftbackward", pairs=plist)
        t_pairs = time.perf_counter() - t0
        n_auto = sum(1 for i, j in plist if i == j)
        # an auto-spectrum call transforms one channel instead of two: count it as half a cross call
        per_cross = t_pairs / (len(plist) - 0.5 * n_auto)
        csm_rate = M * L / (per_cross * (M * (M - 1) / 2 + 0.5 * M))
        rng = np.random.default_rng(seed)
        gx, gy, gz = self.lines
        gi = rng.choice(self.G, size=min(points, self.G), replace=False)
        ix, iy, iz = np.unravel_index(gi, (len(gx), len(gy), len(gz)))
        gsel = np.stack([gx[ix], gy[iy], gz[iz]], axis=1)
        t0 = time.perf_counter()
        R.das_map_loop(self.f, self.csm, self.mic, gsel, w["fc"], w["frac"], True)
        t_map = time.perf_counter() - t0
        return dict(csm_rate=csm_rate, map_rate=len(gi) * self.F / t_map, t_sample_s=t_pairs + t_map, G=self.G, F=self.F,
                    sample=f"stage 1: {len(plist)} channel pairs of the full {L}-sample recording scaled to "
                           f"M(M+1)/2={M*(M+1)//2} pairs; stage 2: {len(gi)} of {self.G} grid points x all {self.F} bins "
                           f"scaled to the full grid")


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU algorithm (oracle port, reference loop
    structure; the reference itself is pure Python and cannot travel to the GPU box)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w, pos, L, lines = problem(args)
    cores = os.cpu_count()
    rates, times = [], []
    ref = CpuReference(w, pos, L, lines, channels=3)
    for it in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        r = ref.step(pairs=4, points=12, seed=it)
        if it >= args.warmup:
            rates.append(r); times.append(time.perf_counter() - t0)
    map_rate = float(np.mean([r["map_rate"] for r in rates]))
    csm_rate = float(np.mean([r["csm_rate"] for r in rates]))
    G, F = rates[0]["G"], rates[0]["F"]
    M = w["mics"]
    full_step_s = G * F / map_rate + M * L / csm_rate
    line = {
        "impl": "reference", "metric": "das_map_grid_point_bins_per_s", "value": map_rate, "unit": "grid-pt*bins/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": full_step_s * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": w["label"], "M": M, "L": L, "G": G, "F": F},
        "stages": {"csm": {"value": csm_rate, "unit": "ch-samples/s"}, "das_map": {"value": map_rate, "unit": "grid-pt*bins/s"}},
        "cpu_baseline": {"value": map_rate, "unit": "grid-pt*bins/s", "cores": cores, "kind": "port",
                         "sample": rates[0]["sample"] + "; ms_per_step is the extrapolated full-size step",
                         "csm_value": csm_rate, "csm_unit": "ch-samples/s", "sample_seconds_per_step": float(np.mean(times))},
        "e2e": {"value": map_rate, "unit": "grid-pt*bins/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# --------------------------------------------------------------------------- B200 arm
def gemm_peak(torch, dtype="tf32", sustained_seconds=2.5):
    """cuBLAS GEMM throughput (8192^3) in TF32 or FP16: best of 5 (burst) and a back-to-back loop of
    `sustained_seconds` (sustained, under the power cap) -- the tensor roofline denominators measured
    on the same GPU in the same run (MEASURED_PEAKS.json carries bf16 only).
    Returns (burst, sustained) in TFLOP/s."""
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    tdt = torch.float16 if dtype == "f16" else torch.float32
    a = torch.randn(8192, 8192, device="cuda").to(tdt); b = torch.randn(8192, 8192, device="cuda").to(tdt)
    for _ in range(2):
        a @ b
    best = 1e9
    for _ in range(5):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a @ b; e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    n = max(8, int(sustained_seconds * 1e3 / best))
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n):
        a @ b
    e.record(); torch.cuda.synchronize()
    sustained = 2 * 8192 ** 3 * n / s.elapsed_time(e) / 1e9
    torch.backends.cuda.matmul.allow_tf32 = prev
    del a, b
    return 2 * 8192 ** 3 / best / 1e9, sustained


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
    mode_name = {0: "simt_f64", 1: "simt_f32", 2: "tf32", 3: "tf32x3", 4: "mixed", 5: "f16x3"}[mode]

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

    from dsptoolbox_b200.standard._spectral_methods import csm_welch_device

    class _Sig(dsp.Signal):
        """Signal whose samples are this rank's rows of the recording (device tensor or pinned host view)."""

        def __init__(self, data, row_offset):
            self._rows, self._row_offset = data, row_offset
            super().__init__(None, np.zeros((2, M)), FS)
            self.set_spectrum_parameters(window_length_samples=wl)

        @property
        def number_of_channels(self):
            return M

        @property
        def length_samples(self):
            return L

        def _compute_csm_device(self, frame_range=None, reduce_fn=None):
            q = self._spectrum_parameters
            return csm_welch_device(self._rows, FS, q["window_length_samples"], q["window_type"], q["overlap_percent"],
                                    q["detrend"], q["average"], q["scaling"], frame_range=(k0, k1),
                                    reduce_fn=bdist.allreduce_sum_ if world > 1 else None, total_length=L,
                                    row_offset=self._row_offset)

    BFClass = bdist.DistributedBeamformerDASFrequency if world > 1 else dsp.beamforming.BeamformerDASFrequency

    class _BF(BFClass):
        def _csm_device(self):
            return self.signal.get_csm_device()

    def make_bf(sig):
        return _BF(sig, dsp.beamforming.MicArray(pos), grid, dsp.beamforming.SteeringVector())

    sig_res = _Sig(x_dev[t0s:t1s], t0s)       # device-resident inputs
    sig_e2e = _Sig(x_np, t0s)                 # pinned host inputs
    bf_res, bf_e2e = make_bf(sig_res), make_bf(sig_e2e)
    devnull = open(os.devnull, "w")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        """Stage 1 then stage 2, device-resident input; returns CUDA events around each stage."""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        fq, csm = sig_res.get_csm_device()
        ev[1].record()
        bf_res.f_range_hz = np.array([w["fc"] * 2 ** (-1 / w["frac"] / 2), w["fc"] * 2 ** (1 / w["frac"] / 2)]) if w["frac"] else np.array([w["fc"]] * 2)
        g0, g1 = bf_res._grid_rows()
        id1, id2 = bf_res._select_bins(fq)
        weights = dsp.beamforming.simpson_weights(id2 - id1, fq[1] - fq[0]) if id2 - id1 > 1 else np.ones(1)
        op = engine.das_prepare(csm, id1, id2 - id1, True, mode)
        out, _ = engine.das_map(op, M, fq[id1:id2], weights, 343.0, mic, bf_res.mics.array_center_coordinates,
                                grid._coordinates[g0:g1], N.STEER_TRUE_LOCATION, True, mode)
        if world > 1:
            counts = [b - a for a, b in (bdist.grid_shard(G, r, world) for r in range(world))]
            out = bdist.allgather_rows(out, counts)
        ev[2].record()
        return ev, out

    def step_e2e():
        so = sys.stdout
        sys.stdout = devnull
        try:
            return bf_e2e.get_beamformer_map(w["fc"], w["frac"], True)
        finally:
            sys.stdout = so

    # ---- warm-up (>= 3), then K timed steps bracketed by barrier + synchronize
    for _ in range(max(3, args.warmup)):
        step_resident()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = engine.launches
    t_csm, t_map = [], []
    wall0 = time.perf_counter()
    e_start, e_stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e_start.record()
    evs = []
    for _ in range(args.steps):
        ev, out = step_resident()
        evs.append(ev)
    e_stop.record()
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

    # ---- kernel-level timing of the dominant kernel (map) and of the stage-1 kernels, one more step
    engine.config.profile = True
    sig_res.get_csm_device()
    k_times = dict(engine.config.stats["csm"]["times_ms"])
    engine.config.profile = False
    fq, csm = sig_res.get_csm_device()
    id1, id2 = band(fq, w["fc"], w["frac"])
    weights = dsp.beamforming.simpson_weights(id2 - id1, fq[1] - fq[0]) if id2 - id1 > 1 else np.ones(1)
    g0, g1 = bf_res._grid_rows()
    op = engine.das_prepare(csm, id1, id2 - id1, True, mode)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    engine.das_map(op, M, fq[id1:id2], weights, 343.0, mic, bf_res.mics.array_center_coordinates,
                   grid._coordinates[g0:g1], N.STEER_TRUE_LOCATION, True, mode)
    b.record(); torch.cuda.synchronize()
    ms_map_kernel = maxr(a.elapsed_time(b))
    del op, csm

    # ---- end-to-end through the public host API
    e2e = None
    if not args.no_e2e:
        for _ in range(1):
            step_e2e()
        barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(args.steps, 3))
        for _ in range(n_e2e):
            m = step_e2e()
        barrier()
        t_e2e = maxr((time.perf_counter() - t0) / n_e2e)
        h2d = x_np.nbytes
        if world > 1:
            tb = torch.tensor([float(h2d)], device=dev, dtype=torch.float64)
            dist.all_reduce(tb)
            h2d = int(tb.item())
        e2e = {"value": G * F / t_e2e, "unit": "grid-pt*bins/s", "ms_per_step": t_e2e * 1e3,
               "h2d_bytes_per_step": int(h2d + G * 24 + M * 24), "d2h_bytes_per_step": int(G * 8),
               "api": "BeamformerDASFrequency.get_beamformer_map(fc, frac, True) on a Signal holding pinned host float64 samples",
               "csm_value": M * L / t_e2e, "csm_unit": "ch-samples/s (whole step)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # FP16x3 mode runs three kind::f16 products per algorithmic product -> FP16 GEMM peak / 3; the TF32-based
    # modes are counted in TF32-equivalents against the TF32 GEMM peak
    peak_kind = "f16" if mode == N.MAP_TC_F16X3 else "tf32"
    peak_tf32_burst, peak_tf32_sustained = gemm_peak(torch, peak_kind)
    # the map kernel runs for seconds under the 1 kW cap: the sustained figure is its denominator;
    # a kernel shorter than ~0.5 s is compared with the burst figure
    peak_tf32 = peak_tf32_sustained if ms_map_kernel > 500.0 else peak_tf32_burst
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm = peaks.get("hbm_gbs", 6650.0)
    # TF32-equivalent tensor products per algorithmic product: 3 (split TF32), 2 (TF32 + one BF16 product of
    # double depth at twice the TF32 rate), 1 (single pass)
    passes = {N.MAP_TC_TF32X3: 3, N.MAP_TC_MIXED: 2, N.MAP_TC_F16X3: 3}.get(mode, 1)
    Gr = g1 - g0
    alg_flops = 4.0 * M * M * Gr * F
    achieved = alg_flops / (ms_map_kernel * 1e-3) / 1e12
    tensor_peak = peak_tf32 / passes
    # DRAM traffic of one launch of this kernel from the committed `ncu --set full` capture (same
    # workload, same mode, one GPU); null when no capture matches
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
                "traffic_note": "dram__bytes_read+write of one launch (ncu, profiles/map_kernel_traffic.json); the operand images "
                                "are re-streamed from HBM per wave of tiles -> ~1 % of HBM bandwidth, compute-bound",
                "algorithmic_flops_per_launch": alg_flops, "kernel_ms": ms_map_kernel,
                "peak_source": f"cuBLAS {peak_kind.upper()} 8192^3 measured in this run: burst {peak_tf32_burst:.1f}, sustained {peak_tf32_sustained:.1f} TFLOP/s; used {peak_tf32:.1f}"
                               + (" / 3 (FP16x3: three FP16 tensor products per algorithmic product)" if mode == N.MAP_TC_F16X3 else "")
                               + (" / 3 (split-TF32: three tensor products per algorithmic product)" if mode == N.MAP_TC_TF32X3 else "")
                               + (" / 2 (mixed: one TF32 product + one double-depth BF16 product = 2 TF32-equivalents)" if passes == 2 else "")
                               + f"; MEASURED_PEAKS bf16 {peaks.get('bf16_tflops')} (of {'measured' if peaks else 'fallback'})"}
    s_in, s_x = (8, 16) if engine.config.csm_precision == "fp64" else (4, 8)
    Kloc = k1 - k0
    fft_bytes = s_in * M * (t1s - t0s) + s_x * M * Kloc * nb
    herk_flops = 4.0 * M * M * nb * Kloc
    stage1 = {}
    if "k1_rfft" in k_times:
        stage1["welch_rfft"] = {"bound": "hbm", "ms": k_times["k1_rfft"], "achieved": fft_bytes / k_times["k1_rfft"] / 1e6,
                                "peak": hbm, "unit": "GB/s", "frac": fft_bytes / k_times["k1_rfft"] / 1e6 / hbm}
    if "k2_herk" in k_times:
        stage1["csm_herk"] = {"bound": "tensor(fp64 DMMA)", "ms": k_times["k2_herk"], "achieved": herk_flops / k_times["k2_herk"] / 1e9,
                              "unit": "TFLOP/s", "algorithmic_flops": herk_flops}
    if "k0_prepare" in k_times:
        b0 = (8 + s_in) * M * (t1s - t0s)
        stage1["prepare_channels"] = {"bound": "hbm", "ms": k_times["k0_prepare"], "achieved": b0 / k_times["k0_prepare"] / 1e6,
                                      "peak": hbm, "unit": "GB/s", "frac": b0 / k_times["k0_prepare"] / 1e6 / hbm}

    cpu = None
    if not args.no_cpu_baseline and world == 1:       # reported at N = 1 only
        r = CpuReference(w, pos, L, lines, channels=6).step(pairs=21, points=96)
        cpu = {"value": r["map_rate"], "unit": "grid-pt*bins/s", "cores": os.cpu_count(), "kind": "port",
               "sample": r["sample"], "csm_value": r["csm_rate"], "csm_unit": "ch-samples/s", "sample_seconds": r["t_sample_s"]}

    line = {
        "metric": "das_map_grid_point_bins_per_s", "value": G * F / (ms_map * 1e-3), "unit": "grid-pt*bins/s",
        "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": {3: "tf32x3 (fp32 accumulate) map; f64 CSM", 4: "tf32+bf16x2 (fp32 accumulate) map; f64 CSM",
                  5: "f16x3 (two FP16 halves per operand, 3 products, fp32 accumulate) map; f64 CSM"}.get(
            mode, f"{mode_name} map; {engine.config.csm_precision} CSM"),
        "data": "synthetic",
        "config": {"workload": w["label"], "M": M, "L": L, "nfft": wl, "frames": K, "bins": nb, "G": G, "F": F,
                   "G_full": G_full, "map_mode": mode_name, "csm_precision": engine.config.csm_precision,
                   "parallelism": f"frames/{world} + all-reduce, grid/{world} + all-gather" if world > 1 else "single GPU",
                   "l2": "inputs larger than L2 (recording %.2f GB, spectra %.2f GB, CSM %.2f GB)" % (M * L * 8 / 1e9, s_x * M * K * nb / 1e9, nb * M * M * 16 / 1e9)},
        "stages": {"csm": {"value": M * L / (ms_csm * 1e-3), "unit": "ch-samples/s", "ms": ms_csm, "kernels_ms": k_times, "rooflines": stage1},
                   "das_map": {"value": G * F / (ms_map * 1e-3), "unit": "grid-pt*bins/s", "ms": ms_map}},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
        "wall_s_timed_region": wall,
    }
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
