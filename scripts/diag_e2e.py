"""Stage 1 from a pinned / pageable HOST recording (the e2e path): wall time to the finished device CSM, the engine's
per-kernel event times, and allocator counters.  usage: python scripts/diag_e2e.py"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import dsptoolbox_b200 as dsp
from dsptoolbox_b200 import engine

M, L, wl = 256, 2_880_000, 4096
pin = torch.empty((L, M), dtype=torch.float64).pin_memory()
pin.normal_()
x_pin = pin.numpy()


def stats():
    m = torch.cuda.memory_stats()
    return m.get("num_device_alloc", 0), m.get("num_device_free", 0), torch.cuda.memory_reserved() >> 20


for name, arr in (("pinned", x_pin), ("pageable", np.array(x_pin)), ("resident", pin.cuda())):
    s = dsp.Signal(None, arr, 48000)
    s.set_spectrum_parameters(window_length_samples=wl, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
    for prof in (False, False, False, True):
        engine.config.profile = prof
        torch.cuda.synchronize(); a0 = stats(); t0 = time.perf_counter()
        f, csm = s.get_csm_device()
        t1 = time.perf_counter()
        torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"{name}: host {1e3 * (t1 - t0):.1f} ms, to-sync {1e3 * (t2 - t0):.1f} ms, alloc/free/reservedMiB {a0} -> {stats()}"
              + (f"  {engine.config.stats['csm']}" if prof else ""), flush=True)
        del csm
    engine.config.profile = False
