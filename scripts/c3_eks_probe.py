#!/usr/bin/env python
"""C3 filter + EKS smoother pass time (N = 65,536, K = 1,000) for the current launch geometry (env switches are read by the library)."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd
import scripts.bench_configs as bc
N, K = 65536, 1000
p = bc.nl_params(3, 1, cd.LearnableLorenz63(sigma=torch.tensor(10.0, **bc.f64), rho=torch.tensor(28.0, **bc.f64), beta=torch.tensor(8 / 3, **bc.f64)),
                 1.0, 1.0, P0=5 * np.eye(3), H=np.array([[1.0, 0, 0]]))
t = bc.times(N, K, 0.01, 3); y = 8 * torch.randn(N, K, 1, **bc.f64)
hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
f = bc.timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp), reps=3)
s = bc.timeit(lambda: cd.cdnlgssm_smoother(p, y, t[..., None], hp), reps=3)
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith("CDK_")}, filter_ms=round(f, 3), filter_plus_eks_ms=round(s, 3), eks_ms=round(s - f, 3))))
