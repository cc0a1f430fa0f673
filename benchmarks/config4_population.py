"""BASELINE.json config 4: population calibration with 1,048,576 candidate parameter sets sharded across
the GPUs of one box, NCCL all-gather of the losses, then a batched multi-start LM on the best starts.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        benchmarks/config4_population.py [candidates]
Rank 0 prints one JSON line: the time of one sharded population evaluation (CUDA events, max over ranks,
all-gather included) and of the whole calibration.
"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import LB, Q, R, S0, TRUTH, UB, surface  # noqa: E402
from pde_b200 import BatchPricer  # noqa: E402
from pde_b200.calibration import PopulationCalibrator, sobol_population  # noqa: E402
from pde_b200.sharding import ShardedPricer, init_from_env  # noqa: E402

n_cand = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
rank, world, local = init_from_env("nccl")
dev = torch.device("cuda", local)
K, T = surface()
pr = BatchPricer("fft", device=local).set_surface(K, T, True, None, S0=S0, r=R, q=Q)
mk = np.maximum(pr.price_host(TRUTH[None, :])[0] * (1 + 0.001 * np.random.default_rng(42).normal(size=K.size)), 0.01)
pr.set_surface(K, T, True, mk, S0=S0, r=R, q=Q).set_bounds(LB, UB)
X = torch.as_tensor(sobol_population(n_cand, LB, UB, seed=42), device=dev)
sp = ShardedPricer(pr)
sp.objective(X[: 4096 * world])  # warm-up (NCCL communicator, kernels)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
loss = sp.objective(X)
b.record()
torch.cuda.synchronize()
ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)

import pandas as pd  # noqa: E402

df = pd.DataFrame({"strike": K, "maturity": T, "mid_price": mk, "is_call": True})
t0 = time.perf_counter()
res = PopulationCalibrator(mode="fft", device=local).calibrate(df, S0, R, Q, n_candidates=n_cand, n_starts=32,
                                                               lm_iters=30)
torch.cuda.synchronize()
t_cal = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t_cal, op=dist.ReduceOp.MAX)
if rank == 0:
    print(json.dumps({
        "config": "C4: population calibration, %d candidates x 32 maturities x 50 strikes, N=4096" % n_cand,
        "n_gpus": world, "population_eval_ms": round(float(ms.item()), 2),
        "population_slices_per_s": round(n_cand * 32 / float(ms.item()) * 1e3),
        "allgather_bytes": n_cand * 8,
        "calibration_seconds": round(float(t_cal.item()), 3), "rmse": round(float(res.rmse), 5),
        "params": [round(float(v), 5) for v in res.params.to_array()],
        "best_loss_of_population": float(loss.min().item())}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
