#!/usr/bin/env python
"""Throughput of every BASELINE config on cuda:0 (kernel timed with CUDA events, inputs resident).  Not the driver's
bench (that is /bench.py, config 3 only): this feeds the per-config table in DESIGN.md.

    python scripts/bench_configs.py [c1 c2 c3 c4 c5 ...] [--scale S]   (S < 1 shrinks N for quick runs)
"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cd_dynamax_b200 as cd  # noqa: E402
from cd_dynamax_b200 import _lib as L  # noqa: E402
from cd_dynamax_b200.continuous_discrete_linear_gaussian_ssm.inference import _filter_device  # noqa: E402
from cd_dynamax_b200.continuous_discrete_nonlinear_gaussian_ssm._common import run_filter  # noqa: E402

dev = torch.device("cuda", 0)
f64 = dict(dtype=torch.float64, device=dev)


def times(N, K, mean_gap, seed):
    g = torch.Generator(device=dev); g.manual_seed(seed)
    gaps = mean_gap * (0.5 + torch.rand(N, K, generator=g, **f64))
    gaps[:, 0] = 0.0
    return torch.cumsum(gaps, dim=1)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e30
    for _ in range(reps):
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def nl_params(n, m, drift, Qc, R, m0=None, P0=None, H=None):
    T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    H = np.eye(n)[:: n // m][:m] if H is None else H
    return cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(T(np.zeros(n) if m0 is None else m0)),
                                      cov=cd.LearnableMatrix(T(np.eye(n) if P0 is None else P0))),
        dynamics=cd.ParamsCDNLGSSMDynamics(drift=drift, diffusion_coefficient=cd.LearnableMatrix(T(np.eye(n))),
                                           diffusion_cov=cd.LearnableMatrix(T(Qc * np.eye(n)))),
        emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=T(H), bias=T(np.zeros(m))),
                                             emission_cov=cd.LearnableMatrix(T(R * np.eye(m)))))


def c1(scale):
    n, m, N, K = 4, 2, 1, 200
    F = np.zeros((n, n)); F[0, 2] = F[1, 3] = 1.0
    p = cd.ParamsCDLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=np.array([8.0, 10.0, 1.0, 0.0]), cov=0.1 * np.eye(n)),
        dynamics=cd.ParamsCDLGSSMDynamics(weights=F, bias=np.zeros(n), input_weights=None,
                                          diffusion_coefficient=1e-3 * np.eye(n), diffusion_cov=np.eye(n)),
        emissions=cd.ParamsLGSSMEmissions(weights=np.eye(m, n), bias=np.zeros(m), input_weights=None, cov=0.5 * np.eye(m)))
    t = times(N, K, 0.05, 1)
    y = torch.randn(N, K, m, **f64) + 8
    hp = cd.KFHyperParams(dt_final=1.0, diffeqsolve_settings={"solver": "rk4", "dt0": 0.0125})
    ms_f = timeit(lambda: cd.cdlgssm_filter(p, y, t[..., None], hp), reps=10, warm=3)
    ms_s = timeit(lambda: cd.cdlgssm_smoother(p, y, t[..., None], hp), reps=10, warm=3)
    return dict(config="C1 tracking KF n=4 m=2 N=1 K=200", filter_ms=ms_f, smoother_ms=ms_s, note="latency-bound (single trajectory)")


def c2(scale):
    n, m, K = 16, 4, 500
    N = max(148, int(8192 * scale))
    rng = np.random.default_rng(1235)
    F = -0.5 * np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64), device=dev)
    p = cd.ParamsCDLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=T(np.zeros(n)), cov=T(np.eye(n))),
        dynamics=cd.ParamsCDLGSSMDynamics(weights=T(F), bias=T(np.zeros(n)), input_weights=None,
                                          diffusion_coefficient=T(np.eye(n)), diffusion_cov=T(0.1 * np.eye(n))),
        emissions=cd.ParamsLGSSMEmissions(weights=T(np.eye(n)[:m]), bias=T(np.zeros(m)), input_weights=None, cov=T(0.1 * np.eye(m))))
    t = times(N, K, 0.04, 2)
    y = torch.randn(N, K, m, **f64)
    hp = cd.KFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.01})
    ms_f = timeit(lambda: cd.cdlgssm_filter(p, y, t[..., None], hp))
    ms_s = timeit(lambda: cd.cdlgssm_smoother(p, y, t[..., None], hp))
    hp_def = cd.KFHyperParams()  # the reference's defaults: Dopri5, dt0 = 0.01 (diffrax_utils.py:50, :121-124)
    ms_fd = timeit(lambda: cd.cdlgssm_filter(p, y, t[..., None], hp_def))
    ms_sd = timeit(lambda: cd.cdlgssm_smoother(p, y, t[..., None], hp_def))
    q = 4.5
    fl_f = (74752 * q + 23915) * N * K
    fl_s = fl_f + 52500 * N * K  # the type-1 smoother reads the filter's (A, Q) back instead of re-integrating them
    return dict(config=f"C2 KF n=16 m=4 N={N} (of 262,144) K=500", filter_ms=ms_f, filter_obs_steps_per_s=N * K / ms_f * 1e3,
                filter_tflops_survey=fl_f / ms_f / 1e9, smoother_ms=ms_s, smoother_obs_steps_per_s=N * K / ms_s * 1e3,
                smoother_tflops_executed=fl_s / ms_s / 1e9,
                default_hyperparams_dopri5=dict(filter_ms=ms_fd, filter_obs_steps_per_s=N * K / ms_fd * 1e3,
                                                smoother_ms=ms_sd, smoother_obs_steps_per_s=N * K / ms_sd * 1e3))


def c3(scale):
    N, K = max(224, int(65536 * scale)), 1000
    p = nl_params(3, 1, cd.LearnableLorenz63(sigma=torch.tensor(10.0, **f64), rho=torch.tensor(28.0, **f64), beta=torch.tensor(8 / 3, **f64)),
                  1.0, 1.0, P0=5 * np.eye(3), H=np.array([[1.0, 0, 0]]))
    t = times(N, K, 0.01, 3)
    y = 8 * torch.randn(N, K, 1, **f64)
    out = {}
    for name, fields in (("all outputs", None), ("ll only", [])):
        hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
        ms = timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp, output_fields=fields))
        out[name] = dict(ms=ms, obs_steps_per_s=N * K / ms * 1e3, tflops_survey=(496 * 4.5 + 107) * N * K / ms / 1e9)
    hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
    ms = timeit(lambda: cd.cdnlgssm_smoother(p, y, t[..., None], hp), reps=2)
    out["filter+EKS smoother"] = dict(ms=ms, obs_steps_per_s=N * K / ms * 1e3)
    y32, t32 = y.float(), t.float()
    ms = timeit(lambda: cd.cdnlgssm_filter(p, y32, t32[..., None], hp))
    out["fp32 variant, all outputs"] = dict(ms=ms, obs_steps_per_s=N * K / ms * 1e3)
    ms = timeit(lambda: cd.ekf_marginal_log_prob_and_grad(p, y, t[..., None], hp), reps=2)
    out["ll + d ll / d (sigma, rho, beta), forward mode"] = dict(ms=ms, obs_steps_per_s=N * K / ms * 1e3)
    ms = timeit(lambda: cd.ekf_marginal_log_prob_and_grad(p, y, t[..., None], hp, wrt="all"), reps=2)
    out["ll + gradient w.r.t. ALL parameters (23 columns), reverse mode"] = dict(ms=ms, obs_steps_per_s=N * K / ms * 1e3)
    return dict(config=f"C3 EKF Lorenz-63 N={N} K=1000", **out)


def c4(scale):
    n, m, K = 40, 20, 500
    N = max(148, int(8192 * scale))
    p = nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(4).standard_normal(n))
    t = times(N, K, 0.02, 4)
    y = 8 + 2 * torch.randn(N, K, m, **f64)
    hp = cd.UKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.005})
    os.environ["CDK_UKF_SIGMA_POINTS"] = "0"
    ms = timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp), reps=2)
    ms_ll = timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp, output_fields=[]), reps=2)
    os.environ["CDK_UKF_SIGMA_POINTS"] = "1"
    Ns = min(N, 2048)
    ms_sig = timeit(lambda: cd.cdnlgssm_filter(p, y[:Ns], t[:Ns, :, None], hp), reps=1)
    os.environ["CDK_UKF_SIGMA_POINTS"] = "0"
    # executed flops of the closed-form kernel per observation-step: 18 RK stages x (n^2 entries x ~14 flop) + update ~0.2 M
    return dict(config=f"C4 UKF Lorenz-96 n=40 m=20 N={N} (of 8,192) K=500", ms=ms, obs_steps_per_s=N * K / ms * 1e3,
                ll_only_ms=ms_ll, ll_only_obs_steps_per_s=N * K / ms_ll * 1e3,
                note="closed-form unscented predict (default); tflops_survey counts the sigma-point algorithm's flops and is NOT "
                     "what this kernel executes",
                tflops_survey_equivalent=5.98e6 * N * K / ms / 1e9,
                hbm_gbs_outputs=26240.0 * N * K / ms / 1e6,
                sigma_point_kernel=dict(N=Ns, ms=ms_sig, obs_steps_per_s=Ns * K / ms_sig * 1e3,
                                        tflops_survey=5.98e6 * Ns * K / ms_sig / 1e9))


def c5(scale):
    n, m, K, E = 40, 20, 500, 1024
    N = max(37, int(1024 * scale))
    p = nl_params(n, m, cd.LearnableLorenz96(forcing=torch.tensor(8.0, **f64)), 0.1, 1.0, m0=8 + 0.5 * np.random.default_rng(5).standard_normal(n))
    t = times(N, K, 0.02, 5)
    y = 8 + 2 * torch.randn(N, K, m, **f64)
    hp = cd.EnKFHyperParams(N_particles=E, key=1234, diffeqsolve_settings={"solver": "euler", "dt0": 0.005})
    ms = timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp), reps=2)
    ms_ll = timeit(lambda: cd.cdnlgssm_filter(p, y, t[..., None], hp, output_fields=[]), reps=2)
    return dict(config=f"C5 EnKF Lorenz-96 n=40 m=20 E=1024 N={N} (of 1,024) K=500", ms=ms, obs_steps_per_s=N * K / ms * 1e3,
                tflops_survey=(0.4e6 * 4.5 + 13.2e6) * N * K / ms / 1e9, ll_only_ms=ms_ll, ll_only_obs_steps_per_s=N * K / ms_ll * 1e3)


if __name__ == "__main__":
    args = [a for a in sys.argv[1:] if a in ("c1", "c2", "c3", "c4", "c5")]
    scale = float(sys.argv[sys.argv.index("--scale") + 1]) if "--scale" in sys.argv else 1.0
    for name in (args or ["c1", "c2", "c3", "c4", "c5"]):
        t0 = time.time()
        try:
            r = globals()[name](scale)
        except Exception as e:  # keep going: one config must not hide the others
            r = dict(config=name, error=repr(e))
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
