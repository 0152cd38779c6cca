"""Upload paths for the 5.9 GB benchmark recording: direct cudaMemcpyAsync from one large pinned allocation vs. the
pinned staging ring of engine._Stager (fed from pinned and from pageable memory), in 768 MB chunks like the engine."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine

L, M = 2_880_000, 256
dev = torch.device("cuda", 0)
pin = torch.empty((L, M), dtype=torch.float64).pin_memory()
pin.normal_()
x_pin = pin.numpy()
x_page = np.array(x_pin)
rows = (768 << 20) // (M * 8)
torch.cuda.synchronize()

def run(name, fn, reps=3):
    for r in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        keep = [fn(x0) for x0 in range(0, L, rows)]
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(f"{name}: {dt*1e3:.1f} ms  {L*M*8/dt/1e9:.1f} GB/s", flush=True)
        del keep

run("pinned, direct .to()", lambda r0: torch.from_numpy(x_pin[r0:r0 + rows]).to(dev, non_blocking=True))
run("pinned, staging ring", lambda r0: engine._stager.upload(x_pin[r0:r0 + rows], dev))
run("pageable, staging ring", lambda r0: engine._stager.upload(x_page[r0:r0 + rows], dev))
run("pageable, driver .to()", lambda r0: torch.from_numpy(x_page[r0:r0 + rows]).to(dev, non_blocking=True), reps=2)
for thr in (2, 4, 8, 12):
    engine.config.h2d_stage_threads = thr
    run(f"pageable, staging ring, {thr} threads", lambda r0: engine._stager.upload(x_page[r0:r0 + rows], dev), reps=2)
