#!/usr/bin/env python
"""C4 (UKF / EKF, Lorenz-96 n = 40, m = 20) pass time with the BASELINE solver (rk4, dt0 = 0.005) and with the reference's
default hyper-parameters (dopri5, dt0 = 0.01)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd
import scripts.bench_configs as bc
n, m, K, N = 40, 20, 500, int(sys.argv[1]) if len(sys.argv) > 1 else 2048
p = bc.nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **bc.f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(4).standard_normal(n))
t = bc.times(N, K, 0.02, 4); y = 8 + 2 * torch.randn(N, K, m, **bc.f64)
for algo, HP in (("ukf", cd.UKFHyperParams), ("ekf", cd.EKFHyperParams)):
    for solver, dt0 in (("rk4", 0.005), ("dopri5", 0.01), ("heun", 0.005), ("euler", 0.005)):
        hp = HP(diffeqsolve_settings={"solver": solver, "dt0": dt0})
        ms = bc.timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp), reps=2)
        print(json.dumps(dict(algo=algo, solver=solver, dt0=dt0, N=N, ms=round(ms, 2), obs_steps_per_s=round(N * K / ms * 1e3))), flush=True)
