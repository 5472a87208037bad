"""CPU, world_size = 2 over gloo: the trajectory-sharding host logic (cd_dynamax_b200/parallel.py).  The per-shard filter
is the NumPy oracle here (no GPU in this tier); on the GPU box the same code path runs with NCCL and the CUDA kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cd_dynamax_b200 import parallel
from oracle import cd_oracle as o


def test_shard_bounds_cover_the_batch_exactly():
    for n in (0, 1, 2, 7, 8, 9, 65536, 65537):
        for w in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        parallel.shard_bounds(4, 2, 2)


def _problem(N=7, K=12):
    rng = np.random.default_rng(3)
    gaps = 0.01 * rng.uniform(0.5, 1.5, size=(N, K)); gaps[:, 0] = 0.0
    t = np.cumsum(gaps, axis=1)
    y = 8.0 * rng.standard_normal((N, K, 1))
    p = o.NonlinearParams(m0=np.zeros(3), P0=5 * np.eye(3), drift=o.Lorenz63Drift(), L=np.eye(3), Qc=np.eye(3),
                          H=np.array([[1.0, 0.0, 0.0]]), R=np.eye(1))
    return p, y, t


def _worker(rank, world, port, algo, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p, y, t = _problem()
        if algo == "ekf":
            fn = lambda yy, tt, uu, off: o.extended_kalman_filter(p, yy, tt, settings=o.SolverSettings("rk4", 0.0025))["marginal_loglik"]
        else:
            fn = lambda yy, tt, uu, off: o.ensemble_kalman_filter(p, yy, tt, E=16, seed=9, rng_offset=off,
                                                                   settings=o.SolverSettings("euler", 0.0025))["marginal_loglik"]
        total = parallel.sharded_marginal_log_prob(fn, y, t)
        out[rank] = float(total)
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("algo", ["ekf", "enkf"])
def test_sharded_loglik_over_gloo_matches_single_process(algo):
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), algo, out), nprocs=world, join=True)
    p, y, t = _problem()
    if algo == "ekf":
        ref = o.extended_kalman_filter(p, y, t, settings=o.SolverSettings("rk4", 0.0025))["marginal_loglik"].sum()
    else:  # rng_offset makes the EnKF stream independent of the sharding
        ref = o.ensemble_kalman_filter(p, y, t, E=16, seed=9, settings=o.SolverSettings("euler", 0.0025))["marginal_loglik"].sum()
    assert len(out) == world
    for r in range(world):
        assert abs(out[r] - ref) <= 1e-12 * abs(ref), (out[r], ref)


def test_allreduce_rejects_wrong_payload():
    with pytest.raises(ValueError):
        parallel.allreduce_loglik(torch.zeros(2, dtype=torch.float64))
    assert parallel.allreduce_loglik(torch.ones(1, dtype=torch.float64)).item() == 1.0  # no process group: identity
