"""Measure the roofline denominators MEASURED_PEAKS.json lacks: cuBLAS TF32 and FP64 dense GEMM throughput."""
import json, sys, time
import torch

def gemm_tflops(dtype, n, tf32, seconds=1.5):
    torch.backends.cuda.matmul.allow_tf32 = tf32
    a = torch.randn((n, n), device="cuda", dtype=dtype)
    b = torch.randn((n, n), device="cuda", dtype=dtype)
    for _ in range(3):
        a @ b
    torch.cuda.synchronize()
    best = 0.0
    t_end = time.time() + seconds
    total_ms, total_n = 0.0, 0
    while time.time() < t_end:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            a @ b
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        best = max(best, 2.0 * n ** 3 / ms / 1e9)
        total_ms += ms; total_n += 1
    torch.backends.cuda.matmul.allow_tf32 = False
    return best, 2.0 * n ** 3 / (total_ms / total_n) / 1e9

def measure():
    tb, ts = gemm_tflops(torch.float32, 8192, True)
    db, ds = gemm_tflops(torch.float64, 8192, False)
    return {"tf32_tflops": tb, "tf32_tflops_sustained": ts, "fp64_tflops": db, "fp64_tflops_sustained": ds,
            "how": "torch.matmul 8192^3 (cuBLAS): fp32 inputs with allow_tf32 (TF32 tensor cores) and fp64; best of 5-launch groups (burst) and mean over 1.5 s (sustained)"}

if __name__ == "__main__":
    print(json.dumps(measure()))
