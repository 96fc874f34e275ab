"""Measure the operational FP64 GEMM peak of this B200 via cuBLAS (torch.matmul, float64).

Burst = best of 10 single 8192^3 GEMMs; sustained = back-to-back for ~4 s (same protocol as the
driver's MEASURED_PEAKS.json bf16 entries). Also a batched SYRK-like shape. Prints one JSON line.
"""
import json, os, subprocess, time
import torch

def clocks():
    try:
        out = subprocess.check_output(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                                       "--format=csv,noheader"], text=True).strip()
        return out
    except Exception as e:  # noqa
        return str(e)

def main():
    dev = torch.device("cuda:0")
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    c = torch.empty(n, n, dtype=torch.float64, device=dev)
    for _ in range(3):
        torch.matmul(a, b, out=c)
    torch.cuda.synchronize()
    flops = 2.0 * n ** 3
    best = 1e9
    for _ in range(10):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b, out=c); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    burst = flops / best * 1e-9
    # sustained
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    reps = max(10, int(4000.0 / best))
    e0.record()
    for i in range(reps):
        torch.matmul(a, b, out=c)
        if i == reps // 2:
            mid = None
    e1.record()
    clk_mid = clocks()
    torch.cuda.synchronize()
    sustained = flops * reps / e0.elapsed_time(e1) * 1e-9
    # NT form (A @ B^T), the shape of the Cholesky trailing products
    for _ in range(2):
        torch.matmul(a, b.t(), out=c)
    torch.cuda.synchronize()
    bestnt = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); torch.matmul(a, b.t(), out=c); e1.record(); torch.cuda.synchronize()
        bestnt = min(bestnt, e0.elapsed_time(e1))
    # library Cholesky bar: torch.linalg.cholesky batched 8 x 4096
    k = torch.randn(8, 4096, 4096, dtype=torch.float64, device=dev)
    k = k @ k.transpose(-1, -2) + 4096 * torch.eye(4096, dtype=torch.float64, device=dev)
    torch.linalg.cholesky(k); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); torch.linalg.cholesky(k); e1.record(); torch.cuda.synchronize()
    chol_ms = e0.elapsed_time(e1)
    chol_tf = 8 * 4096 ** 3 / 3.0 / chol_ms * 1e-9
    print(json.dumps({"fp64_tflops": round(burst, 2), "fp64_tflops_sustained": round(sustained, 2),
                      "fp64_nt_tflops": round(flops / bestnt * 1e-9, 2),
                      "lib_cholesky_8x4096_ms": round(chol_ms, 2), "lib_cholesky_tflops": round(chol_tf, 2),
                      "clocks_mid_sustained": clk_mid, "gpu": torch.cuda.get_device_name(0),
                      "mem_gb": round(torch.cuda.get_device_properties(0).total_memory / 2**30, 1),
                      "host_cores": os.cpu_count(), "torch_threads": torch.get_num_threads()}))

if __name__ == "__main__":
    main()
