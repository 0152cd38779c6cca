"""Timing of the other gridded beamformers (MVDR, Functional, Orthogonal, CleanSC) through the public
API, on a cfg2/cfg3-shaped problem: 64 microphones, 10 s at 48 kHz, nfft 4096, a 100 x 100 grid, the third-octave band
around 2 kHz (40 bins).  One JSON line: per beamformer the time of `get_beamformer_map` with the CSM already cached on
the device (CUDA events around the call, median of the timed calls), grid-pt*bins/s, and the share of that time spent
in this library's map kernels (engine.config.stats).  Usage: python scripts/bench_f1.py [steps] [warmup]"""
import contextlib, io, json, sys
import numpy as np, torch

sys.path.insert(0, ".")
import dsptoolbox_b200 as dsp
from dsptoolbox_b200 import beamforming as BF, engine, synth

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
warmup = int(sys.argv[2]) if len(sys.argv) > 2 else 1
FS, M, SECONDS, WL = 48_000, 64, 10, 4096
pos = synth.disc_array(M, 0.75, seed=7)
x = synth.point_source_recording_torch(synth.positions_to_xyz(pos), SECONDS * FS, FS, source_xyz=(0.3, -0.2, 1.0), seed=3)
sig = dsp.Signal(None, x, FS, activate_cache=True)
sig.set_spectrum_parameters(window_length_samples=WL, overlap_percent=50, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
grid = BF.Regular2DGrid(np.linspace(-1, 1, 100), np.linspace(-1, 1, 100), ("x", "y"), 1.0)
ma, st = BF.MicArray(pos), BF.SteeringVector(BF.SteeringVectorType.TrueLocation)
G = grid.number_of_points


def timed(make, *args, **kw):
    times, last = [], None
    for i in range(warmup + steps):
        bf = make()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        with contextlib.redirect_stdout(io.StringIO()):
            last = bf.get_beamformer_map(*args, **kw)
        e1.record()
        torch.cuda.synchronize()
        if i >= warmup:
            times.append(e0.elapsed_time(e1))
    F = 1 + int(round((bf.f_range_hz[1] - bf.f_range_hz[0]) / (FS / WL)))
    ms = float(np.median(times))
    peak = np.unravel_index(int(np.argmax(last)), last.shape)
    return {"ms": ms, "bins": F, "grid_pt_bins_per_s": G * F / (ms * 1e-3), "peak_at": [int(p) for p in peak]}


with contextlib.redirect_stdout(io.StringIO()):
    sig.get_csm_device()      # cached on the device: the CSM stage is not what is timed here
res = {
    "das": timed(lambda: BF.BeamformerDASFrequency(sig, ma, grid, st), 2000.0, 3, True),
    "mvdr": timed(lambda: BF.BeamformerMVDR(sig, ma, grid, st), 2000.0, 3),
    "functional": timed(lambda: BF.BeamformerFunctional(sig, ma, grid, st), 2000.0, 3, 10),
    "orthogonal": timed(lambda: BF.BeamformerOrthogonal(sig, ma, grid, st), 2000.0, 3, 8),
    "cleansc": timed(lambda: BF.BeamformerCleanSC(sig, ma, grid, st), 2000.0, 3, 16),
    "cleansc_128_iterations": timed(lambda: BF.BeamformerCleanSC(sig, ma, grid, st), 2000.0, 3),
}
print(json.dumps({"metric": "f1 beamformer maps (public API, CSM cached on the device)", "unit": "ms per get_beamformer_map",
                  "config": {"M": M, "seconds": SECONDS, "nfft": WL, "G": G, "fc": 2000.0, "octave_fraction": 3,
                             "map_mode": engine.config.map_mode, "steps": steps, "warmup": warmup},
                  "results": res}))
