"""Objective-only throughput of the fused FFT kernel (BASELINE.json config 4: population / multi-start
search evaluates losses only): config-3 surface, N = 4096.    python benchmarks/objective_rate.py [sets]
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import LB, Q, R, S0, TRUTH, UB, sobol_sets, surface  # noqa: E402
from pde_b200 import BatchPricer  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 131072  # 1M candidates over 8 GPUs
K, T = surface()
pr = BatchPricer("fft").set_surface(K, T, True, None, S0=S0, r=R, q=Q)
mk = np.maximum(pr.price_host(TRUTH[None, :])[0] * (1 + 0.001 * np.random.default_rng(42).normal(size=K.size)), 0.01)
pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
X = torch.tensor(sobol_sets(P), device="cuda:0")
pr.objective(X)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
loss = pr.objective(X)
b.record()
torch.cuda.synchronize()
ms = a.elapsed_time(b)
rate = P * 32 / ms * 1e3
print(json.dumps({"sets": P, "maturities": 32, "strikes": 50, "ms": round(ms, 2), "slices_per_s": round(rate),
                  "w1_tflops": round(rate * 3113560 / 1e12, 2),
                  "sentinel_losses": int((loss >= 1e10).sum().item())}))
