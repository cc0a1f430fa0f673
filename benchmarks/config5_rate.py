"""BASELINE.json config 5 throughput: N = 16384, 128 maturities x 200 strikes, objective only.
    python benchmarks/config5_rate.py [sets]      # needs a B200
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import LB, Q, R, S0, TRUTH, UB, sobol_sets  # noqa: E402
from pde_b200 import BatchPricer  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
out = {}
for eta in (0.25, 0.0625):
    K = np.tile(np.linspace(80.0, 120.0, 200), 128)
    T = np.repeat(np.linspace(0.1, 1.0, 128), 200)
    pr = BatchPricer("fft", n_grid=16384, eta=eta).set_surface(K, T, True, None, S0=S0, r=R, q=Q)
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
    out[f"eta={eta}"] = {"sets": P, "slices": P * 128, "ms": round(ms, 2), "slices_per_s": round(P * 128 / ms * 1e3),
                         "w1_tflops": round(P * 128 / ms * 1e3 * 12618080 / 1e12, 2)}
print(json.dumps(out))
