"""Where does the time of the stage-1 step go?  Runs Signal.get_csm_device on a resident recording a few times
and prints, per call: host wall time of the call (no sync), GPU event time, host time per sub-phase with syncs,
and the caching allocator's cudaMalloc / cudaFree counters.  usage: python scripts/diag_csm.py [M L wl]"""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import dsptoolbox_b200 as dsp
from dsptoolbox_b200 import engine
from dsptoolbox_b200.standard import _spectral_methods as sm

M, L, wl = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (256, 2880000, 4096)
x = torch.randn((L, M), device="cuda", dtype=torch.float64)
s = dsp.Signal(None, x, 48000)
s.set_spectrum_parameters(window_length_samples=wl, scaling=dsp.SpectrumScaling.PowerSpectralDensity)
big = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")


def stats():
    m = torch.cuda.memory_stats()
    return m.get("num_device_alloc", 0), m.get("num_device_free", 0), m.get("num_alloc_retries", 0), torch.cuda.memory_reserved() >> 20


for it in range(6):
    torch.cuda.synchronize()
    a0 = stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    big.zero_()                                  # some GPU work in front, like the previous step's map kernel
    e0.record()
    t0 = time.perf_counter()
    f, csm = s.get_csm_device()
    t1 = time.perf_counter()
    e1.record()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"call {it}: host {1e3 * (t1 - t0):.2f} ms, gpu {e0.elapsed_time(e1):.2f} ms, to-sync {1e3 * (t2 - t0):.2f} ms, "
          f"alloc/free/retries/reservedMiB {a0} -> {stats()}  kernels {engine.config.stats['csm']['times_ms'] if engine.config.profile else ''}")
    del csm

# sub-phases with syncs
p = s._spectrum_parameters
window, step = sm._window_and_step(p["window_type"], wl, p["overlap_percent"])
for it in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    S, K = engine.csm_partial_sums(x, L, wl, step, window, True, 1.0)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    csm = engine.csm_finalize(S, M, K, 1.0, True, False)
    torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"phases {it}: partial_sums {1e3 * (t1 - t0):.2f} ms, finalize {1e3 * (t2 - t1):.2f} ms")
    del S, csm
engine.config.profile = True
S, K = engine.csm_partial_sums(x, L, wl, step, window, True, 1.0)
torch.cuda.synchronize()
print(engine.config.stats["csm"])
