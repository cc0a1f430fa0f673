"""Throughput of the `refgrid` mode (the reference's own 1023-point quadrature, heston.cpp:94-151) on the
config-3 surface: objective and FD normal equations.     python benchmarks/refgrid_rate.py [sets]   # B200
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import LB, Q, R, S0, TRUTH, UB, sobol_sets, surface  # noqa: E402
from pde_b200 import BatchPricer  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K, T = surface()
pr = BatchPricer("refgrid").set_surface(K, T, True, None, S0=S0, r=R, q=Q)
mk = np.maximum(pr.price_host(TRUTH[None, :])[0] * (1 + 0.001 * np.random.default_rng(42).normal(size=K.size)), 0.01)
pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
X = torch.tensor(sobol_sets(P), device="cuda:0")
out = {"sets": P, "maturities": 32, "strikes": 50}
for name, fn, mult in (("objective", pr.objective, 1), ("normal_equations", pr.normal_equations, 6)):
    fn(X)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    fn(X)
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    # W1 convention for refgrid: 1,137,576 FLOPs per slice
    rate = P * 32 * mult / ms * 1e3
    out[name] = {"ms": round(ms, 2), "slice_evals_per_s": round(rate), "w1_tflops": round(rate * 1137576 / 1e12, 2)}
print(json.dumps(out))
