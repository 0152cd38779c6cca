"""Debug: steady-state event timeline of CTA 0 of the map kernel (needs a -DB2_MAP_PROFILE -DB2_MAP_TRACE build via
B2DSP_LIB): the probe points of the role profile with their clock64 stamps, for the two issue warps and one warp of
each synthesis group.  usage: M G F mode [rows]"""
import ctypes, sys
import numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N
lib = N.lib()
M, G, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = engine._MAP_MODES[sys.argv[4]]
rows = int(sys.argv[5]) if len(sys.argv) > 5 else 1500
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randn((F, M, 4), device="cuda", dtype=torch.complex128, generator=g)
csm = a @ a.conj().transpose(1, 2)
rng = np.random.default_rng(0)
mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), np.zeros(M)], 1)
grid = torch.from_numpy(np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1)).cuda()
op = engine.das_prepare(csm, 0, F, True, mode)
engine.das_map(op, M, 11.71875 * np.arange(F), np.ones(F), 343.0, mic, mic[0], grid, 3, True, mode)
torch.cuda.synchronize()
LEN = 2048
buf = (ctypes.c_longlong * (5 * LEN * 2))()
lib.b2_debug_map_trace(buf)
raw = np.array(buf[:], dtype=np.int64).reshape(5, LEN, 2)
names = {0: "W0", 1: "W1", 2: "G0", 3: "G1", 4: "G2"}
what = {1: "tempty.", 2: "fullB.", 3: "fullA.", 4: "issued", 5: "released",
        25: "tdone.", 26: "reduced", 27: "emptyA.", 28: "synth", 29: "arrived", 31: "table"}
ev = []
for r in range(5):
    k = 0
    for i in range(LEN):
        if raw[r, i, 1] == 0:
            continue
        slot = int(raw[r, i, 0])
        ev.append((int(raw[r, i, 1]), r, what.get(slot if slot >= 24 else slot % 8, str(slot))))
ev.sort()
tmin = min(t for t, _, _ in ev)
for rr in (0, 1):                       # the issue warps log 5 events per step: their windows end before the groups' begin
    print(f"---- {names[rr]} alone")
    last = None
    for t, r, w in [e for e in ev if e[1] == rr][:240]:
        print(f"{t - tmin:9d} {names[r]} {w}" + (f"  (+{t - last})" if last is not None else ""))
        last = t
print("---- all roles")
t0 = max(min(t for t, r, _ in ev if r == rr) for rr in range(2, 5))     # all groups are being logged from here on
n = 0
for t, r, w in ev:
    if t < t0:
        continue
    print(f"{t - t0:9d} " + "            " * r + f"{names[r]} {w}")
    n += 1
    if n >= rows:
        break
