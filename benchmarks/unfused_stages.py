"""The unfused stage kernels on their own (A/B evidence for fusing):
  fft_batch_kernel  K2 alone, slices staged HBM -> shared memory by cp.async.bulk (TMA): HBM-bound,
                    131,072 algorithmic bytes per N=4096 slice (64 KiB in + 64 KiB out)
  cf_kernel         K1 alone writing phi to HBM: 16 B out per grid point
against MEASURED_PEAKS.json hbm_gbs.   python benchmarks/unfused_stages.py   (needs a B200)
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from bench import sobol_sets  # noqa: E402
from pde_b200 import characteristic_function, fft_batch  # noqa: E402

peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
hbm = json.load(open(peaks))["hbm_gbs"] if os.path.exists(peaks) else 6650.0


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


out = {"hbm_peak_gbs": hbm}
n_slices, n = 32768, 4096
x = torch.randn(n_slices, n, dtype=torch.complex128, device="cuda:0")  # 2 GiB > L2
ms = timed(lambda: fft_batch(x))
gb = 2 * n_slices * n * 16 / 1e9
out["fft_batch_kernel"] = {"slices": n_slices, "ms": round(ms, 3), "slices_per_s": round(n_slices / ms * 1e3),
                           "achieved_gbs": round(gb / ms * 1e3, 1), "frac_of_hbm_peak": round(gb / ms * 1e3 / hbm, 3)}
del x
P, M = 4096, 8
X = torch.tensor(sobol_sets(P), device="cuda:0")
T = torch.linspace(0.1, 1.0, M, dtype=torch.float64, device="cuda:0")
u = torch.tensor(0.25 * np.arange(n) - 1.75j, device="cuda:0")
ms = timed(lambda: characteristic_function(X, T, u, S0=100.0, r=0.05, q=0.02))
pts = P * M * n
out["cf_kernel"] = {"points": pts, "ms": round(ms, 3), "slices_per_s": round(P * M / ms * 1e3),
                    "achieved_gbs_out": round(pts * 16 / 1e9 / ms * 1e3, 1),
                    "w1_tflops": round(pts * 700 / ms * 1e3 / 1e12, 2)}
print(json.dumps(out))
