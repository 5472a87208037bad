#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "enkf or EnKF or forecast or sample" 2>&1 | tail -4
timeout 600 python scripts/bench_configs.py c5 2>&1 | cut -c1-330
python - <<'PY'
import os, sys, json, numpy as np, torch
sys.path.insert(0, '.')
import cd_dynamax_b200 as cd
import scripts.bench_configs as bc
n, m, K, E, N = 40, 20, 500, 1024, 1024
p = bc.nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **bc.f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(5).standard_normal(n))
t = bc.times(N, K, 0.02, 5); y = 8 + 2 * torch.randn(N, K, m, **bc.f64)
for light in ("1", "0"):
    os.environ["CDK_ENKF_LIGHT"] = light
    hp = cd.EnKFHyperParams(N_particles=E, key=1234, diffeqsolve_settings={"solver": "heun", "dt0": 0.005})
    ms = bc.timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp), reps=2)
    print(json.dumps(dict(config="C5 with the reference's default SDE solver (heun)", light=light, ms=round(ms, 1), obs_steps_per_s=round(N * K / ms * 1e3))))
PY
