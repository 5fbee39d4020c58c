"""FP64 roof cross-check: cuBLAS DGEMM (torch.matmul on float64 = cublasDgemm) next to the library's own DMMA probe
(sr_measure_fp64_peak), same GPU, same process. SURVEY §8d asked for a library number beside the hand-written one.
Usage: python tools/dgemm_peak.py [--n 8192] [--out gpurun_out/fp64_peak_r02.json]"""
import argparse
import json
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from subrecon_b200 import native  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=8192)
ap.add_argument("--out", default="")
a = ap.parse_args()

n = a.n
x = torch.randn(n, n, dtype=torch.float64, device="cuda")
y = torch.randn(n, n, dtype=torch.float64, device="cuda")
z = torch.empty(n, n, dtype=torch.float64, device="cuda")
for _ in range(2):
    torch.matmul(x, y, out=z)
torch.cuda.synchronize()
best, times = 1e30, []
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    torch.matmul(x, y, out=z)
    e1.record()
    torch.cuda.synchronize()
    times.append(e0.elapsed_time(e1))
best = min(times)
# sustained: back to back for ~3 s
reps = max(1, int(3000.0 / best))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    torch.matmul(x, y, out=z)
e1.record()
torch.cuda.synchronize()
sustained_ms = e0.elapsed_time(e1) / reps
del x, y, z
ctx = native.Context(0)
dmma = [ctx.measure_fp64_peak(150.0) for _ in range(3)]
ctx.close()
clk = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw", "--format=csv,noheader,nounits"],
                     capture_output=True, text=True).stdout.strip()
rec = {"gpu": torch.cuda.get_device_name(0),
       "cublas_dgemm": {"n": n, "tflops_best": 2.0 * n ** 3 / (best * 1e-3) * 1e-12, "tflops_sustained_3s": 2.0 * n ** 3 / (sustained_ms * 1e-3) * 1e-12,
                        "ms_best": best, "ms_all": times, "how": "torch.matmul(float64) -> cublasDgemm, CUDA events, 2 warm-up calls"},
       "dmma_probe_tflops": dmma, "dmma_probe_how": "sr_measure_fp64_peak: register-resident mma.sync.m8n8k4.f64 loop, 150 ms, all SMs",
       "nominal_tflops": 148 * 64 * 2 * 1.965e9 * 1e-12, "clocks_after_sm_max_power": clk}
print(json.dumps(rec))
if a.out:
    json.dump(rec, open(a.out, "w"), indent=1)
