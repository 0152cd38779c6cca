"""Signal.get_csm() to HOST arrays at cfg4 size (2.15 GB complex128 result): plain .cpu() vs the pinned staging ring."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import dsptoolbox_b200 as dsp
from dsptoolbox_b200 import engine

x = torch.randn((2_880_000, 256), device="cuda", dtype=torch.float64)
s = dsp.Signal(None, x, 48000)
s.set_spectrum_parameters(window_length_samples=4096)
for threads in (6, 6, 0, 0, 6):
    engine.config.h2d_stage_threads = threads
    torch.cuda.synchronize(); t0 = time.perf_counter()
    f, c = s.get_csm()
    dt = time.perf_counter() - t0
    print(f"threads {threads}: get_csm() -> numpy {c.shape} {c.dtype} in {dt*1e3:.1f} ms", flush=True)
ref = s.get_csm_device()[1].cpu().numpy()
engine.config.h2d_stage_threads = 6
print("identical:", np.array_equal(s.get_csm()[1], ref))
