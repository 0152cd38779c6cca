"""Run the stage-2 kernels on a synthetic Hermitian CSM (for ncu / timing): M G F mode reps."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N

M, G, F = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
mode = {"mixed": N.MAP_TC_MIXED, "f16x3": N.MAP_TC_F16X3, "tf32x3": N.MAP_TC_TF32X3, "tf32": N.MAP_TC_TF32, "simt_f32": N.MAP_SIMT_F32, "simt_f64": N.MAP_SIMT_F64}[sys.argv[4]]
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
g = torch.Generator(device="cuda"); g.manual_seed(0)
a = torch.randn((F, M, 4), device="cuda", dtype=torch.complex128, generator=g)
csm = a @ a.conj().transpose(1, 2)
rng = np.random.default_rng(0)
mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), np.zeros(M)], 1)
grid = np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1)
freqs = 11.71875 * np.arange(F)
w = np.ones(F)
op = engine.das_prepare(csm, 0, F, True, mode)
gdev = torch.from_numpy(grid).cuda()
for _ in range(reps):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out, _ = engine.das_map(op, M, freqs, w, 343.0, mic, mic[0], gdev, 3, True, mode)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    passes = 3 if mode == N.MAP_TC_TF32X3 else 1
    print(f"{sys.argv[4]} M={M} G={G} F={F}: {dt*1e3:.2f} ms  {G*F/dt/1e6:.1f} M gpb/s  alg {4*M*M*G*F/dt/1e12:.1f} TF/s", flush=True)
