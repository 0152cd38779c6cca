"""Debug: per-role blocked cycles of the map kernel (needs the -DB2_MAP_PROFILE build via B2DSP_LIB)."""
import ctypes, os, subprocess, sys
import numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N
lib = N.lib()
M, G, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = {"mixed": N.MAP_TC_MIXED, "f16x3": N.MAP_TC_F16X3, "tf32x3": N.MAP_TC_TF32X3, "tf32": N.MAP_TC_TF32}[sys.argv[4]]
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randn((F, M, 4), device="cuda", dtype=torch.complex128, generator=g)
csm = a @ a.conj().transpose(1, 2)
rng = np.random.default_rng(0)
mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), np.zeros(M)], 1)
grid = torch.from_numpy(np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1)).cuda()
op = engine.das_prepare(csm, 0, F, True, mode)
buf = (ctypes.c_ulonglong * 8)()
for rep in range(2):
    lib.b2_debug_map_profile(buf)
    engine.das_map(op, M, 11.71875 * np.arange(F), np.ones(F), 343.0, mic, mic[0], grid, 3, True, mode)
    torch.cuda.synchronize()
    lib.b2_debug_map_profile(buf)
v = list(buf)
names = ["synth wait empty", "kernel total", "mma wait full", "mma wait tmem-empty", "mma total", "epi wait tmem-full", "epi total", "mma wait B full"]
tot = v[1]
for n, x in zip(names, v):
    print(f"{n:22s} {x:14d} cycles  {100.0 * x / tot:5.1f}% of kernel")
