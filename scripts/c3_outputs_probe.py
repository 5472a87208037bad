#!/usr/bin/env python
"""C3 pass time as a function of which moment arrays are requested and of the batch size (what the output path costs)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd
import scripts.bench_configs as bc
K = 1000
p = bc.nl_params(3, 1, cd.LearnableLorenz63(sigma=torch.tensor(10.0, **bc.f64), rho=torch.tensor(28.0, **bc.f64), beta=torch.tensor(8 / 3, **bc.f64)),
                 1.0, 1.0, P0=5 * np.eye(3), H=np.array([[1.0, 0, 0]]))
hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
Ns = [int(a) for a in sys.argv[1:]] or [65536]
for N in Ns:
    t = bc.times(N, K, 0.01, 3); y = 8 * torch.randn(N, K, 1, **bc.f64)
    for fields in ([], ["filtered_means", "filtered_covariances"], None):
        ms = bc.timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp, output_fields=fields), reps=5)
        print(json.dumps(dict(N=N, fields=fields if fields is not None else "all", ms=round(ms, 3), traj_per_us=round(N / ms / 1e3, 2))), flush=True)
