#!/usr/bin/env python
"""Small KF (n=16) / UKF (n=40) / EnKF (n=40, E=1024) launches for ncu: python scripts/profile_generic.py kf|ukf|enkf"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd
sys.argv = [sys.argv[0]] + sys.argv[1:]
import scripts.bench_configs as bc  # noqa
which = sys.argv[1]
dev = bc.dev
if which == "kf":
    n, m, K, N = 16, 4, 100, 1184
    rng = np.random.default_rng(1235)
    F = -0.5 * np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    p = cd.ParamsCDLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=T(np.zeros(n)), cov=T(np.eye(n))),
        dynamics=cd.ParamsCDLGSSMDynamics(weights=T(F), bias=T(np.zeros(n)), input_weights=None,
                                          diffusion_coefficient=T(np.eye(n)), diffusion_cov=T(0.1 * np.eye(n))),
        emissions=cd.ParamsLGSSMEmissions(weights=T(np.eye(n)[:m]), bias=T(np.zeros(m)), input_weights=None, cov=T(0.1 * np.eye(m))))
    t = bc.times(N, K, 0.04, 2); y = torch.randn(N, K, m, **bc.f64)
    hp = cd.KFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.01})
    for _ in range(2):
        cd.cdlgssm_filter(p, y, t[..., None], hp)
elif which == "enkf":
    n, m, K, N, E = 40, 20, 20, 37, 1024
    p = bc.nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **bc.f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(5).standard_normal(n))
    t = bc.times(N, K, 0.02, 5); y = 8 + 2 * torch.randn(N, K, m, **bc.f64)
    hp = cd.EnKFHyperParams(N_particles=E, key=1234, diffeqsolve_settings={"solver": "euler", "dt0": 0.005})
    for _ in range(2):
        cd.cdnlgssm_filter(p, y, t[..., None], hp)
else:
    n, m, K, N = 40, 20, 20, 296
    p = bc.nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **bc.f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(4).standard_normal(n))
    t = bc.times(N, K, 0.02, 4); y = 8 + 2 * torch.randn(N, K, m, **bc.f64)
    hp = cd.UKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.005})
    for _ in range(2):
        cd.cdnlgssm_filter(p, y, t[..., None], hp)
torch.cuda.synchronize()
