"""Quick numeric check of the tcgen05 map kernel against the fp64 CUDA-core kernel."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
from dsptoolbox_b200 import engine, _native as N

def run(M, G, F, mode, seed=0, steering=3, remove_diag=True, csm_scale=1.0, geo=1.0, bin_ramp=False):
    rng = np.random.default_rng(seed)
    nb = F + 3
    a = rng.standard_normal((nb, M, 6)) + 1j * rng.standard_normal((nb, M, 6))
    csm = a @ np.conj(np.swapaxes(a, 1, 2)) * csm_scale
    if bin_ramp:                                   # very different levels per bin
        csm = csm * (10.0 ** np.linspace(-8, 8, nb))[:, None, None]
    csm_dev = torch.from_numpy(csm).cuda()
    mic = np.stack([rng.uniform(-0.7, 0.7, M), rng.uniform(-0.7, 0.7, M), rng.uniform(-0.05, 0.05, M)], 1)
    grid = np.stack([rng.uniform(-1, 1, G), rng.uniform(-1, 1, G), rng.uniform(0.5, 1.5, G)], 1) * geo
    freqs = 500.0 + 23.4375 * np.arange(F)
    w = rng.uniform(0.5, 1.5, F)
    op = engine.das_prepare(csm_dev, 2, F, remove_diag, mode)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out, pb = engine.das_map(op, M, freqs, w, 343.0, mic, mic[0], grid, steering, remove_diag, mode, per_bin=True)
    torch.cuda.synchronize()
    return out.cpu().numpy(), pb.cpu().numpy(), time.perf_counter() - t0

MODES = [(N.MAP_TC_F16X3, "f16x3"), (N.MAP_TC_TF32X3, "tf32x3"), (N.MAP_TC_MIXED, "mixed"), (N.MAP_TC_TF32, "tf32")]
for (M, G, F) in [(64, 300, 5), (16, 128, 3), (256, 1000, 9), (100, 257, 4), (200, 129, 3)]:
    ref, pref, _ = run(M, G, F, N.MAP_SIMT_F64)
    for mode, name in MODES:
        for rd in (True, False):
            ref, pref, _ = run(M, G, F, N.MAP_SIMT_F64, remove_diag=rd)
            got, pgot, dt = run(M, G, F, mode, remove_diag=rd)
            e = np.max(np.abs(got - ref)) / np.max(ref)
            ep = np.max(np.abs(pgot - pref)) / np.max(pref)
            print(f"M={M} G={G} F={F} {name} rd={rd}: map err/max {e:.2e}  per-bin err/max {ep:.2e}  ({dt*1e3:.1f} ms)", flush=True)

print("--- dynamic range / formulations (f16x3 vs mixed)")
for steering in range(4):
    for kw in (dict(), dict(csm_scale=1e-14), dict(csm_scale=1e9, geo=40.0), dict(geo=0.02), dict(bin_ramp=True)):
        ref, pref, _ = run(96, 400, 6, N.MAP_SIMT_F64, steering=steering, **kw)
        line = f"steer={steering} {kw}:"
        for mode, name in MODES[:3:2]:
            got, pgot, dt = run(96, 400, 6, mode, steering=steering, **kw)
            e = np.max(np.abs(got - ref)) / np.max(ref)
            ep = np.max(np.abs(pgot - pref) / np.max(pref, axis=0, keepdims=True))
            line += f"  {name}: map {e:.2e} per-bin(rel. to bin max) {ep:.2e}"
        print(line, flush=True)
