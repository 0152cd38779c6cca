"""Role-level cycle profile of the map kernel (needs a library built with -DB2_MAP_PROFILE, passed as B2DSP_LIB):
where CTA 0's issue warps and its first synthesis warp spend their cycles.  usage: M G F mode"""
import ctypes, sys, numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N

M, G, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = engine._MAP_MODES[sys.argv[4] if len(sys.argv) > 4 else "f16x3"]
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randn((F, M, 4), device="cuda", dtype=torch.complex128, generator=g)
csm = a @ a.conj().transpose(1, 2)
rng = np.random.default_rng(0)
mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), np.zeros(M)], 1)
grid = torch.from_numpy(np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1)).cuda()
freqs = 11.71875 * np.arange(F)
op = engine.das_prepare(csm, 0, F, True, mode)
lib = N.lib()
buf = (ctypes.c_ulonglong * 32)()
for rep in range(2):
    engine.das_map(op, M, freqs, np.ones(F), 343.0, mic, mic[0], grid, 3, True, mode)
    torch.cuda.synchronize()
    lib.b2_debug_map_profile(buf)
v = list(buf)
for w in (0, 1, 2):
    b = v[8 * w: 8 * w + 8]
    steps = max(1, b[6])
    print(f"issue warp {w}: total {b[0]} cyc, steps {b[6]}; per step: tempty {b[1]/steps:.0f} fullB {b[2]/steps:.0f} fullA {b[3]/steps:.0f} "
          f"issue+commit {b[4]/steps:.0f}; releases of A tiles it does not multiply {b[5]} cyc total ({b[5]/max(1,b[0])*100:.1f} %)")
    if b[0] == 0:
        continue
    print(f"   shares: tempty {b[1]/b[0]*100:.1f}% fullB {b[2]/b[0]*100:.1f}% fullA {b[3]/b[0]*100:.1f}% issue {b[4]/b[0]*100:.1f}% release {b[5]/b[0]*100:.1f}%")
b = v[24:32]
ch = max(1, b[6])
print(f"synthesis warp 0 (group 0): total {b[0]} cyc, own chunks {b[6]}; per own chunk: wait tdone {b[1]/ch:.0f} ld+fma {b[2]/ch:.0f} wait emptyA {b[3]/ch:.0f} "
      f"synth+stores {b[4]/ch:.0f} fence+arrive {b[5]/ch:.0f}; table {b[7]} cyc total")
print(f"   shares: tdone {b[1]/b[0]*100:.1f}% ld+fma {b[2]/b[0]*100:.1f}% emptyA {b[3]/b[0]*100:.1f}% synth {b[4]/b[0]*100:.1f}% fence {b[5]/b[0]*100:.1f}% table {b[7]/b[0]*100:.1f}%")
