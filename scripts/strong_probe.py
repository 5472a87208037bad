#!/usr/bin/env python
"""Pass time of the C3 filter (CD-EKF Lorenz-63, K = 1,000) as a function of the batch size on ONE GPU: what a shard of the
fixed N = 65,536 problem costs at 2 / 4 / 8 GPUs (N = 32,768 / 16,384 / 8,192).  CDK_LW_WARPS=k forces the warps per CTA
(read once per process), so geometries are compared across runs of this script.

    python scripts/strong_probe.py [N ...]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd  # noqa: E402

dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)
K = 1000
Ns = [int(a) for a in sys.argv[1:]] or [2048, 4096, 8192, 16384, 32768, 65536]
T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
p = cd.ParamsCDNLGSSM(
    initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(T(np.zeros(3))), cov=cd.LearnableMatrix(T(5 * np.eye(3)))),
    dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableLorenz63(sigma=T(10.0), rho=T(28.0), beta=T(8 / 3)),
                                       diffusion_coefficient=cd.LearnableMatrix(T(np.eye(3))),
                                       diffusion_cov=cd.LearnableMatrix(T(np.eye(3)))),
    emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=T([[1.0, 0, 0]]), bias=T(np.zeros(1))),
                                         emission_cov=cd.LearnableMatrix(T(np.eye(1)))))
hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
for N in Ns:
    g = torch.Generator(device=dev); g.manual_seed(N)
    gaps = 0.01 * (0.5 + torch.rand(N, K, generator=g, **f64)); gaps[:, 0] = 0
    t = torch.cumsum(gaps, 1)
    y = 8 * torch.randn(N, K, 1, generator=g, **f64)
    out = {"N": N, "K": K, "CDK_LW_WARPS": os.environ.get("CDK_LW_WARPS", "auto")}
    for name, fn in (("filter_all_outputs", lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp)),
                     ("filter_ll_only", lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp, output_fields=[])),
                     ("filter_plus_eks", lambda: cd.cdnlgssm_smoother(p, y, t[..., None], hp))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e30
        for _ in range(5):
            e0.record(); fn(); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        out[name + "_ms"] = round(best, 4)
    out["speedup_vs_7p5ms_full_batch"] = round(7.5 / out["filter_all_outputs_ms"], 2)
    print(json.dumps(out), flush=True)
