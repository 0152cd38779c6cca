"""Measure the library peaks this path's rooflines need and MEASURED_PEAKS.json lacks:
cuBLAS TF32 and FP64 GEMM throughput (torch.matmul 8192^3 / 4096^3, best of 10, CUDA events)."""
import json, sys, torch

def bench(dtype, n, tf32):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn(n, n, device="cuda", dtype=dtype); b = torch.randn(n, n, device="cuda", dtype=dtype)
    for _ in range(3): a @ b
    best = 1e9
    for _ in range(10):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); a @ b; e.record(); torch.cuda.synchronize()
        best = min(best, s.elapsed_time(e))
    return 2 * n ** 3 / best / 1e9   # TFLOP/s

out = dict(gpu=torch.cuda.get_device_name(0),
           tf32_tflops=bench(torch.float32, 8192, True),
           fp32_tflops=bench(torch.float32, 8192, False),
           fp64_tflops=bench(torch.float64, 4096, False))
print(json.dumps(out))
if len(sys.argv) > 1:
    json.dump(out, open(sys.argv[1], "w"), indent=1)
