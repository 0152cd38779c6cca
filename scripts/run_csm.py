"""Run the stage-1 pipeline a few times on synthetic noise (for ncu / timing): M L wl precision reps."""
import sys, numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine
from oracle import restatement  # noqa  (not used: keeps import cost comparable)

M, L, wl = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
prec = sys.argv[4] if len(sys.argv) > 4 else "fp64"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
engine.config.csm_precision = prec
engine.config.profile = True
x = torch.randn((L, M), device="cuda", dtype=torch.float64)
from scipy.signal.windows import get_window
w = get_window("hann", wl, fftbins=True)
for _ in range(reps):
    S, K = engine.csm_partial_sums(x, L, wl, wl // 2, w, True, 1.0)
    csm = engine.csm_finalize(S, M, K, 1.0, False, True)
    torch.cuda.synchronize()
    print(engine.config.stats["csm"]["times_ms"])
