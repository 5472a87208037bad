#!/usr/bin/env python
"""BASELINE config 2 at FULL size through the chunked streaming path: CD-KF, n = 16, m = 4, N = 262,144, K = 500, host-resident
emissions and time stamps (5.2 GB pinned), 285 GB of moments produced and consumed on the device chunk by chunk (sum of
the log-likelihoods + a checksum of every moment array).  End-to-end obs-steps/s including the host->device copies.

    python scripts/c2_full.py [--n 262144] [--chunk 16384] [--smoother] [--solver rk4|dopri5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd  # noqa: E402
from cd_dynamax_b200 import streaming  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=262144)
ap.add_argument("--chunk", type=int, default=16384)
ap.add_argument("--solver", default="rk4")
ap.add_argument("--smoother", action="store_true")
a = ap.parse_args()
dev = torch.device("cuda", 0)
n, m, K, N = 16, 4, 500, a.n
rng = np.random.default_rng(1235)
F = -0.5 * np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
p = cd.ParamsCDLGSSM(
    initial=cd.ParamsLGSSMInitial(mean=T(np.zeros(n)), cov=T(np.eye(n))),
    dynamics=cd.ParamsCDLGSSMDynamics(weights=T(F), bias=T(np.zeros(n)), input_weights=None,
                                      diffusion_coefficient=T(np.eye(n)), diffusion_cov=T(0.1 * np.eye(n))),
    emissions=cd.ParamsLGSSMEmissions(weights=T(np.eye(n)[:m]), bias=T(np.zeros(m)), input_weights=None, cov=T(0.1 * np.eye(m))))
hp = cd.KFHyperParams(diffeqsolve_settings={"solver": a.solver, "dt0": 0.01})
t0 = time.time()
y_host = torch.empty((N, K, m), dtype=torch.float64).pin_memory()
t_host = torch.empty((N, K, 1), dtype=torch.float64).pin_memory()
g = torch.Generator(device=dev); g.manual_seed(2)
for lo in range(0, N, 32768):  # synthetic inputs generated on the device block by block, parked in pinned host memory
    hi = min(N, lo + 32768)
    gaps = 0.04 * (0.5 + torch.rand(hi - lo, K, generator=g, dtype=torch.float64, device=dev)); gaps[:, 0] = 0
    t_host[lo:hi, :, 0].copy_(torch.cumsum(gaps, 1))
    y_host[lo:hi].copy_(torch.randn(hi - lo, K, m, generator=g, dtype=torch.float64, device=dev))
torch.cuda.synchronize()
setup_s = time.time() - t0
acc = {"ll": torch.zeros((), dtype=torch.float64, device=dev), "chk": torch.zeros((), dtype=torch.float64, device=dev), "bad": 0}


def consume(post, lo, hi):
    acc["ll"] += post.marginal_loglik.sum()
    fields = ("smoothed_means", "smoothed_covariances") if a.smoother else ("filtered_means", "filtered_covariances", "predicted_means", "predicted_covariances")
    for f in fields:
        acc["chk"] += getattr(post, f).sum()


fn = (lambda y, t: cd.cdlgssm_smoother(p, y, t, hp)) if a.smoother else (lambda y, t: cd.cdlgssm_filter(p, y, t, hp))
streaming.filter_in_chunks(fn, y_host[: 2 * a.chunk], t_host[: 2 * a.chunk], a.chunk, consume)  # warm-up
acc["ll"].zero_(); acc["chk"].zero_()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
streaming.filter_in_chunks(fn, y_host, t_host, a.chunk, consume)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
out_bytes = N * K * (2 * (n + n * n)) * 8 if not a.smoother else N * K * ((n + n * n) * 2 + n * n) * 8
print(json.dumps({"config": f"C2 full: CD-KF n=16 m=4 N={N} K={K} ({'filter + type-1 smoother' if a.smoother else 'filter'}, {a.solver})",
                  "chunk": a.chunk, "ms": ms, "obs_steps_per_s_end_to_end": N * K / ms * 1e3,
                  "h2d_bytes": int(y_host.numel() * 8 + t_host.numel() * 8), "moment_bytes_produced_and_consumed_on_device": int(out_bytes),
                  "ll_sum": float(acc["ll"].item()), "moment_checksum": float(acc["chk"].item()), "finite": bool(torch.isfinite(acc["chk"]).item()),
                  "setup_s": round(setup_s, 1), "peak_device_GB": round(torch.cuda.max_memory_allocated() / 1e9, 2)}), flush=True)
