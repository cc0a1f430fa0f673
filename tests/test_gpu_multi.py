"""GPU, >= 2 devices: parameter-set sharding over NCCL (BASELINE.json config 4).  Skipped on a
single-GPU box; the host logic of the same code path runs on gloo in tests/test_sharding.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["PDE_ROOT"])
import numpy as np, torch, torch.distributed as dist
from pde_b200 import BatchPricer
from pde_b200.sharding import init_from_env, ShardedPricer
from pde_b200.calibration import PopulationCalibrator, HestonCalibrator, sobol_population
rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
K = np.tile(np.linspace(85, 115, 11), 4); T = np.repeat(np.linspace(0.25, 1.0, 4), 11)
pr = BatchPricer("fft", device=local).set_surface(K, T, True, None, S0=100.0, r=0.05, q=0.02)
mk = pr.price_host(np.array([[1.8, 0.06, 0.45, -0.6, 0.05]]))[0] * 1.001
pr.set_surface(K, T, True, mk, S0=100.0, r=0.05, q=0.02)
lb = np.array([0.1, 0.01, 0.01, -0.99, 0.01]); ub = np.array([10.0, 1.0, 2.0, 0.99, 1.0])
X = torch.as_tensor(sobol_population(4099, lb, ub, seed=1), device=dev)   # not divisible by world
full = pr.objective(X)                       # every rank prices everything ...
shard = ShardedPricer(pr).objective(X)       # ... and its shard + all-gather
assert torch.equal(full, shard), "sharded losses differ"
neq = ShardedPricer(pr).normal_equations(X[:257])
assert torch.equal(neq, pr.normal_equations(X[:257]))
import pandas as pd
df = pd.DataFrame({"strike": K, "maturity": T, "mid_price": mk, "is_call": True})
res = PopulationCalibrator(mode="fft", device=local).calibrate(df, 100.0, 0.05, 0.02, n_candidates=8192, n_starts=8, lm_iters=15)
out = torch.tensor(res.params.to_array(), device=dev)
ref = out.clone(); dist.broadcast(ref, 0)
assert torch.equal(out, ref), "ranks disagree on the calibrated parameters"
assert res.rmse < 0.05
if rank == 0: print("MULTI_OK", world, res.rmse)
dist.barrier(); dist.destroy_process_group()
'''


def test_sharded_objective_and_population_nccl(tmp_path):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, PDE_ROOT=ROOT)
    world = 2 if n < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)],
                       env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "MULTI_OK" in r.stdout
