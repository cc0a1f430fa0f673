"""CPU: parameter-set sharding and the result all-gather on the gloo backend,
world_size 2.  The per-shard evaluation is injected; here it is the oracle (tests may use it)."""
import os

import numpy as np
import pytest

from pde_b200.sharding import shard_bounds

S0, R, Q = 100.0, 0.05, 0.02


def test_shard_bounds_partition():
    for P in (0, 1, 7, 8, 65536, 1_000_003):
        for W in (1, 2, 3, 8):
            spans = [shard_bounds(P, r, W) for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == P
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def _worker(rank, world, port, P, q):
    import torch
    import torch.distributed as dist

    from oracle.oracle import MODE_REFGRID, Oracle
    from pde_b200.sharding import sharded_map

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = Oracle()
    K = np.tile(np.linspace(90, 110, 5), 2)
    T = np.repeat([0.5, 1.0], 5)
    mk = orc.price_batch(MODE_REFGRID, [[2.0, 0.04, 0.3, -0.7, 0.04]], K, T, True, S0, R, Q)[0] * 1.01
    rng = np.random.default_rng(0)  # identical candidates on every rank
    X = torch.tensor(np.array([0.1, 0.01, 0.01, -0.99, 0.01]) + rng.random((P, 5)) * np.array([5, 0.5, 1, 1.5, 0.5]))
    calls = []

    def local_objective(Xl):
        calls.append(len(Xl))
        return torch.tensor(orc.objective_batch(MODE_REFGRID, Xl.numpy(), K, T, True, mk, S0, R, Q))

    full = sharded_map(local_objective, X)
    want = torch.tensor(orc.objective_batch(MODE_REFGRID, X.numpy(), K, T, True, mk, S0, R, Q))
    ok = bool(torch.equal(full, want)) and calls == [shard_bounds(P, rank, world)[1] - shard_bounds(P, rank, world)[0]]
    # 2-D results (normal-equation rows) gather as well
    rows = sharded_map(lambda Xl: torch.arange(len(Xl) * 22, dtype=torch.float64).reshape(len(Xl), 22) + rank, X)
    ok = ok and rows.shape == (P, 22)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("P", [7, 16])
def test_sharded_map_gloo_world2(P):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 500) + P
    procs = [ctx.Process(target=_worker, args=(r, 2, port, P, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(results) == [(0, True), (1, True)]


def test_sharded_map_without_process_group_is_identity():
    import torch

    from pde_b200.sharding import sharded_map

    X = torch.arange(10.0).reshape(2, 5)
    assert torch.equal(sharded_map(lambda x: x.sum(1), X), X.sum(1))
