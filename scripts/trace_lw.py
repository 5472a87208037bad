#!/usr/bin/env python
"""Per-warp run times of the Lorenz-63 EKF kernel (cdk_debug_set_trace): how long does each warp of 32 trajectories
live, and how does that depend on how many warps share its SM sub-partition?

    python scripts/trace_lw.py [--n 65536] [--k 1000] [--out gpurun_out/trace_lw.json]
"""
import argparse
import collections
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import cd_dynamax_b200 as cd  # noqa: E402
from cd_dynamax_b200 import _lib as L  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=65536)
ap.add_argument("--k", type=int, default=1000)
ap.add_argument("--out", default="gpurun_out/trace_lw.json")
ap.add_argument("--regular", action="store_true", help="regular grid: every gap has exactly 4 substeps")
ap.add_argument("--no-outputs", action="store_true", help="log-likelihood only (output_fields=[])")
a = ap.parse_args()
dev = torch.device("cuda", 0)
N, K = a.n, a.k
t_np = bench.make_times_rows(0, N, K, bench.CFG["seed"])
if a.regular:
    t_np = np.broadcast_to(0.01 * np.arange(K)[None], (N, K)).copy()
t = torch.as_tensor(t_np, device=dev)
g = torch.Generator(device=dev); g.manual_seed(1)
y = 8.0 * torch.randn(N, K, 1, generator=g, device=dev, dtype=torch.float64)
T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
params = cd.ParamsCDNLGSSM(
    initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(T(np.zeros(3))), cov=cd.LearnableMatrix(T(5 * np.eye(3)))),
    dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableLorenz63(sigma=T(10.0), rho=T(28.0), beta=T(8.0 / 3.0)),
                                       diffusion_coefficient=cd.LearnableMatrix(T(np.eye(3))),
                                       diffusion_cov=cd.LearnableMatrix(T(np.eye(3)))),
    emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=T([[1.0, 0, 0]]), bias=T(np.zeros(1))),
                                         emission_cov=cd.LearnableMatrix(T(np.eye(1)))))
hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
lib = L.lib()
kw = dict(output_fields=[]) if a.no_outputs else {}
nw = (N + 31) // 32
buf = torch.zeros(nw * 4, dtype=torch.int64, device=dev)
for _ in range(2):
    cd.cdnlgssm_filter(params, y, t[..., None], hp, **kw)
torch.cuda.synchronize()
L.check(lib.cdk_debug_set_trace(ctypes.c_void_p(buf.data_ptr())), "set_trace")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
cd.cdnlgssm_filter(params, y, t[..., None], hp, **kw)
e1.record()
torch.cuda.synchronize()
L.check(lib.cdk_debug_set_trace(None), "clear_trace")
r = buf.cpu().numpy().reshape(nw, 4)
t0, t1, smid, wid = r[:, 0], r[:, 1], r[:, 2], r[:, 3]
base = t0.min()
dur = (t1 - t0) / 1e6  # ms
end = (t1 - base) / 1e6
start = (t0 - base) / 1e6
# warps per (SM, sub-partition = hardware warp slot % 4)
cnt = collections.Counter(zip(smid.tolist(), (wid % 4).tolist()))
per_sm = collections.Counter(smid.tolist())
occ = np.array([cnt[(s, w % 4)] for s, w in zip(smid.tolist(), wid.tolist())])
res = {"kernel_ms_events": e0.elapsed_time(e1), "warps": int(nw), "span_ms": float(end.max()),
       "start_ms": {"min": float(start.min()), "max": float(start.max())},
       "dur_ms": {"min": float(dur.min()), "mean": float(dur.mean()), "max": float(dur.max())},
       "warps_per_sm_hist": dict(collections.Counter(per_sm.values())),
       "warps_per_subpartition_hist": dict(collections.Counter(cnt.values())),
       "dur_by_subpartition_occupancy": {int(o): {"n": int((occ == o).sum()), "mean": float(dur[occ == o].mean()),
                                                   "max": float(dur[occ == o].max())} for o in sorted(set(occ.tolist()))},
       "end_ms_percentiles": {str(p): float(np.percentile(end, p)) for p in (1, 10, 25, 50, 75, 90, 99, 100)},
       "mean_warps_alive_fraction": float(dur.sum() / (end.max() * nw))}
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
print(json.dumps(res))
