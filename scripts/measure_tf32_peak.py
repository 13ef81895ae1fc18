"""Measures the dense TF32 tensor-core throughput of this GPU with cuBLAS (torch.matmul, allow_tf32) - the denominator
for the pointwise kernels' tensor-pipe fraction (BASELINE.json north_star asks for it).  The 3xTF32 split issues three
TF32 MMAs per fp32-equivalent product, so a layer's TF32 rate = 3 x its reported fp32-equivalent TFLOP/s."""
import json
import torch

torch.backends.cuda.matmul.allow_tf32 = True
n = 8192
a = torch.randn(n, n, device="cuda")
b = torch.randn(n, n, device="cuda")
for _ in range(3):
    a @ b
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    a @ b
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
t0 = torch.cuda.Event(enable_timing=True)
t1 = torch.cuda.Event(enable_timing=True)
t0.record()
reps = 0
while True:
    for _ in range(10):
        a @ b
    reps += 10
    t1.record()
    torch.cuda.synchronize()
    if t0.elapsed_time(t1) > 3000:
        break
sustained = 2.0 * n ** 3 * reps / (t0.elapsed_time(t1) * 1e-3) / 1e12
print(json.dumps({"tf32_tflops_burst": 2.0 * n ** 3 / (best * 1e-3) / 1e12, "tf32_tflops_sustained": sustained,
                  "how": "torch.matmul fp32 inputs with allow_tf32 (cuBLAS TF32 tensor-core GEMM) 8192^3, best of 10 / 3 s loop",
                  "gpu": torch.cuda.get_device_name(0)}))
