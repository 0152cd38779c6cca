"""Debug: event timeline of CTA 0 of the map kernel (needs the -DB2_MAP_TRACE build via B2DSP_LIB)."""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N
lib = N.lib()
M, G, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = engine._MAP_MODES[sys.argv[4]]
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randn((F, M, 4), device="cuda", dtype=torch.complex128, generator=g)
csm = a @ a.conj().transpose(1, 2)
rng = np.random.default_rng(0)
mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), np.zeros(M)], 1)
grid = torch.from_numpy(np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1)).cuda()
op = engine.das_prepare(csm, 0, F, True, mode)
buf = (ctypes.c_longlong * (4 * 2 * 1024))()
for rep in range(2):
    lib.b2_debug_trace(buf, 4096)
    engine.das_map(op, M, 11.71875 * np.arange(F), np.ones(F), 343.0, mic, mic[0], grid, 3, True, mode)
    torch.cuda.synchronize()
lib.b2_debug_trace(buf, 4096)
raw = np.array(buf[:], dtype=np.int64).reshape(4, 1024, 2)
names = {0: "synth", 1: "epi", 2: "issue0", 3: "issue1"}
ev = [(int(raw[r, k, 1]), r, int(raw[r, k, 0]) >> 16, int(raw[r, k, 0]) & 0xffff) for r in range(4) for k in range(1024) if raw[r, k, 1] != 0]
ev.sort()
t0 = ev[0][0]
for t, r, e, i in ev:
    print(f"{t - t0:8d} {names[r]:7s} ev{e} idx{i}")
