"""Where does BeamformerCleanSC.get_beamformer_map spend its time?  torch.profiler kernel table (bench_f1 problem)."""
import contextlib, io, sys
import numpy as np, torch
sys.path.insert(0, ".")
import dsptoolbox_b200 as dsp
from dsptoolbox_b200 import beamforming as BF, synth
from torch.profiler import profile, ProfilerActivity

FS, M, WL = 48_000, 64, 4096
pos = synth.disc_array(M, 0.75, seed=7)
x = synth.point_source_recording_torch(synth.positions_to_xyz(pos), 10 * FS, FS, source_xyz=(0.3, -0.2, 1.0), seed=3)
sig = dsp.Signal(None, x, FS, activate_cache=True)
sig.set_spectrum_parameters(window_length_samples=WL, overlap_percent=50, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
grid = BF.Regular2DGrid(np.linspace(-1, 1, 100), np.linspace(-1, 1, 100), ("x", "y"), 1.0)
ma, st = BF.MicArray(pos), BF.SteeringVector(BF.SteeringVectorType.TrueLocation)
with contextlib.redirect_stdout(io.StringIO()):
    BF.BeamformerCleanSC(sig, ma, grid, st).get_beamformer_map(2000.0, 3, 16)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        BF.BeamformerCleanSC(sig, ma, grid, st).get_beamformer_map(2000.0, 3, 16)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=60))
