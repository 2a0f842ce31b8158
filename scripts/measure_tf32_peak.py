#!/usr/bin/env python
"""Measure the dense TF32 tensor-core peak the way MEASURED_PEAKS.json measured bf16:
torch.matmul 8192^3 with fp32 operands and TF32 tensor-core math (cuBLAS), best of 10 (burst) and back to back for
4 s (sustained, under the power cap).  Writes profiles/tf32_peak.json; bench.py uses `tf32_tflops_sustained` as the
roofline denominator of the tf32 mode.

    gpurun -- 'python scripts/measure_tf32_peak.py'   (then copy gpurun_out/tf32_peak.json to profiles/)
"""
import json
import os
import time

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
    e0.record(); a @ b; e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.time()
reps = 0
e0.record()
while time.time() - t0 < 4.0:
    for _ in range(20):
        a @ b
    reps += 20
    torch.cuda.synchronize()
e1.record()
torch.cuda.synchronize()
fl = 2.0 * n ** 3
out = {"tf32_tflops": fl / (best * 1e-3) / 1e12, "tf32_tflops_sustained": fl * reps / (e0.elapsed_time(e1) * 1e-3) / 1e12,
       "how": "torch.matmul fp32 8192^3 with allow_tf32 (cuBLAS TF32 tensor cores): best of 10 and back to back for 4 s",
       "gpu_name": torch.cuda.get_device_name(0), "torch": torch.__version__}
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/tf32_peak.json", "w"), indent=1)
print(json.dumps(out))
