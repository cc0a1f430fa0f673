"""Rate of the bench step (config 3: normal equations over P sets x 32 maturities x 50 strikes, N = 4096) and of
the objective-only step, for kernel A/B runs:

    PDE_B200_LIB=pde_b200/csrc/libheston_b200_<variant>.so python benchmarks/step_rate.py [--sets P] [--trunc E]
        [--population sobol|slow] [--steps K]

--trunc: admissible absolute price error of the significance cut (default: the library's 2^-80; 0 = exact).
--population slow: the slow-decay cluster (sigma near 2, rho near -0.99, small kappa theta) where no grid point
can be elided.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import LB, UB, N_MAT, Q, R, S0, TRUTH, slow_decay_sets, sobol_sets, surface  # noqa: E402
from pde_b200 import BatchPricer  # noqa: E402


def _probe_read():
    """Phase-cycle table of a -DHB_PROBE build (None for the product library)."""
    import ctypes as C

    from pde_b200 import _lib

    L = _lib.load()
    if not hasattr(L, "hb_probe_read"):
        return None
    buf = (C.c_ulonglong * 16)()
    L.hb_probe_read(buf)
    return list(buf)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sets", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--trunc", type=float, default=None)
    ap.add_argument("--population", default="sobol", choices=["sobol", "slow"])
    ap.add_argument("--tag", default="")
    a = ap.parse_args()
    K, T = surface()
    pr = BatchPricer("fft")
    if a.trunc is not None:
        pr.set_truncation(a.trunc)
    pr.set_surface(K, T, True, None, S0=S0, r=R, q=Q)
    mk = pr.price_host(TRUTH[None, :])[0]
    mk = np.maximum(mk * (1 + 0.001 * np.random.default_rng(42).normal(size=mk.size)), 0.01)
    pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
    X = torch.from_numpy(sobol_sets(a.sets) if a.population == "sobol" else slow_decay_sets(a.sets)).cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    out = {"tag": a.tag, "lib": os.environ.get("PDE_B200_LIB", "default"), "sets": a.sets, "population": a.population,
           "log_cut": pr.log_cut}
    for name, fn, evals in (("neq", pr.normal_equations, 6), ("objective", pr.objective, 1)):
        for _ in range(3):
            res = fn(X)
        torch.cuda.synchronize()
        _probe_read()
        ms = 0.0
        for _ in range(a.steps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = fn(X)
            e1.record()
            torch.cuda.synchronize()
            ms += e0.elapsed_time(e1)
        ms /= a.steps
        probe = _probe_read()
        if probe is not None:
            names = ["K1", "pass1", "barrier_after_pass1", "pass2+pairbar", "pass3", "barrier_after_pass3", "K3",
                     "finalize"]
            tot = float(sum(probe[:8])) or 1.0
            out[name + "_probe_share"] = {n: round(probe[i] / tot, 4) for i, n in enumerate(names)}
        out[name] = {"ms_per_step": ms, "slice_evals_per_s": a.sets * N_MAT * evals / (ms * 1e-3),
                     "checksum": float(torch.nan_to_num(res.double(), nan=0.0, posinf=0.0, neginf=0.0).clamp(-1e6, 1e6).sum())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
