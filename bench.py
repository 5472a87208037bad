#!/usr/bin/env python
"""bench.py -- filtered observation-steps/s (N x K per second) of the batched CD-EKF hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[2], the configuration the metric is quoted on): CD-EKF on stochastic Lorenz-63,
d_x = 3, d_y = 1 (observe x), N = 65,536 trajectories, K = 1,000 irregular observation times, classical RK4 moment ODE
with dt0 = mean gap / 4.  One "step" = one filter pass over the whole batch (update + predict for all N x K
observation-steps, writing the four moment arrays the reference returns by default) followed by the sum of the
per-trajectory log-likelihoods and, at N > 1 GPUs, one all-reduce of that scalar (cdk_ll_allreduce on a raw NCCL
communicator).  Trajectories are independent and shard with no data-path collective.

`--scaling strong` (default): the north star's FIXED N = 65,536 is split over the ranks (parallel.shard_bounds), so the
            value at 8 GPUs is the 8-GPU time of the SAME job; `--scaling weak`: 65,536 trajectories PER GPU.  At N > 1 the
            other mode is measured too and reported beside the headline (`weak_scaling` / `strong_scaling` object).

`value`   : inputs resident in HBM, CUDA-event timed, max over ranks.
`e2e`     : the same metric through the public drop-in API with HOST (pinned) buffers: every step copies emissions and
            time stamps host->device and reads the log-likelihoods device->host.
`--impl reference`: the CPU restatement of the reference algorithm (oracle/, C + OpenMP, all host cores) on a bounded
            sample of the same workload.  The reference's own JAX path cannot run here (jax/diffrax are not
            installable in this image) -- this is a "CPU restatement (not JAX)".
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "filtered observation-steps/sec (N x K per second)"
UNIT = "obs-steps/s"
CFG = dict(N=65536, K=1000, n=3, m=1, mean_gap=0.01, dt0=0.0025, solver="rk4", seed=1237)
# flop model (1 FMA = 2 flop).  SURVEY 8(d) ALGORITHMIC count: 496*q + 107 per observation-step for n=3, m=1 (dense
# 3x3 products, loop-invariant L Qc L^T hoisted) -- this is the figure `roofline.achieved` is computed from.
# The kernel stores P symmetric (6 entries), uses the sparsity of the Lorenz-63 Jacobian and folds dt into the RK
# coefficients, so it EXECUTES fewer: per RK4 substep 4 stages * (f 8 + J.P 33 + dP 12 + stage/accumulate FMAs 36) - 18
# = 338 (SASS: 127 DFMA + 51 DADD + 25 DMUL per substep); scalar-emission update 60.  Both are reported (`achieved` /
# `achieved_executed`).
FLOP_SUBSTEP_SURVEY, FLOP_UPDATE_SURVEY = 496.0, 107.0
FLOP_SUBSTEP_EXEC, FLOP_UPDATE_EXEC = 338.0, 60.0
BYTES_PER_OBS_STEP = 16 + 192  # y,t in (16 B) + filtered/predicted mean+cov out (24 doubles)
TRAFFIC_NCU_BYTES = 16.44e9  # dram read 2.64 GB + write 13.80 GB: profiles/r01_ekf_small_lw14_inplace_rk.txt


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong: --n-traj is the TOTAL batch, split over the GPUs (north star); weak: --n-traj per GPU")
    ap.add_argument("--n-traj", type=int, default=CFG["N"], help="trajectories (default = BASELINE config 3: 65,536)")
    ap.add_argument("--no-other-mode", action="store_true", help="at N > 1 GPUs skip the second (other scaling mode) measurement")
    ap.add_argument("--k-obs", type=int, default=CFG["K"])
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU leg (profiling runs)")
    ap.add_argument("--cpu-sample-traj", type=int, default=0, help="trajectories in the CPU sample (0 = auto)")
    ap.add_argument("--simple-data", action="store_true",
                    help="emissions = 8*randn instead of a simulated Lorenz-63 truth (profiling runs: skips the "
                         "~100k setup launches that would otherwise sit in front of the profiled kernel)")
    return ap.parse_args()


def make_times(N, K, seed, rank=0):
    rng = np.random.Generator(np.random.PCG64(seed + 1000 * rank))
    gaps = CFG["mean_gap"] * rng.uniform(0.5, 1.5, size=(N, K))
    gaps[:, 0] = 0.0
    return np.cumsum(gaps, axis=1)


def l63(x, s=10.0, r=28.0, b=8.0 / 3.0):
    import torch
    return torch.stack([s * (x[:, 1] - x[:, 0]), x[:, 0] * (r - x[:, 2]) - x[:, 1], x[:, 0] * x[:, 1] - b * x[:, 2]], 1)


def make_emissions_torch(t, seed, device):
    """Synthetic data: simulate the stochastic Lorenz-63 truth (Euler-Maruyama, 8 substeps per gap, Qc = I) and
    observe x with unit noise.  Runs in torch on `device` (setup only, never timed)."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    N, K = t.shape
    x = torch.tensor([1.0, 1.0, 20.0], dtype=torch.float64, device=device).repeat(N, 1)
    x = x + torch.randn(N, 3, generator=g, device=device, dtype=torch.float64)
    y = torch.empty(N, K, 1, dtype=torch.float64, device=device)
    sub = 8
    for k in range(K):
        if k > 0:
            h = ((t[:, k] - t[:, k - 1]) / sub)[:, None]
            for _ in range(sub):
                x = x + h * l63(x) + torch.sqrt(h) * torch.randn(N, 3, generator=g, device=device, dtype=torch.float64)
        y[:, k, 0] = x[:, 0] + torch.randn(N, generator=g, device=device, dtype=torch.float64)
    return y


def substeps_total(t, dt0, dt_final=1e-10):
    """Sum over all gaps of the solver substep count q_k (diffrax clipping rule), vectorised closed form checked
    against oracle.substep_counts in tests."""
    from oracle.cd_oracle import substep_counts
    t1 = np.concatenate([t[:, 1:], t[:, -1:] + dt_final], axis=1)
    # bounded memory: process in row blocks
    tot = 0
    for i in range(0, t.shape[0], 8192):
        tot += int(substep_counts(t[i:i + 8192], t1[i:i + 8192], dt0).sum())
    return tot


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            p = [c.strip() for c in r.split(",")]
            if len(p) < 8:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


def cpu_reference_leg(n_traj, K, steps, warmup):
    """Time the CPU restatement (oracle/, C + OpenMP when built, else NumPy) on a bounded sample of the workload."""
    from oracle import cpu_baseline
    return cpu_baseline.time_ekf_l63(n_traj, K, CFG, steps=steps, warmup=warmup)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_traj = args.cpu_sample_traj or 512 * cores
    res = cpu_reference_leg(n_traj, args.k_obs, steps=args.steps, warmup=min(args.warmup, 1))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, extra={"cpu_sample": res["sample"]}),
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                         "sample": res["sample"]},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU restatement of the reference algorithm (not JAX: jax/diffrax are not installable in this image)",
    }
    print(json.dumps(line), flush=True)


def workload_config(args, extra=None):
    c = {"workload": "CD-EKF stochastic Lorenz-63 (BASELINE configs[2])", "n_traj_per_gpu": args.n_traj,
         "k_obs": args.k_obs, "d_x": 3, "d_y": 1, "state_order": "second", "solver": "rk4", "dt0": CFG["dt0"],
         "mean_gap": CFG["mean_gap"], "outputs": "filtered+predicted means and covariances, log-likelihood",
         "parallelism": f"trajectory-sharded x{args.gpus} (no data-path collective; 1 all-reduce of sum ll)",
         "l2": "inputs+outputs per step (>= 1 GB) exceed the 126 MB L2; no explicit flush"}
    if extra:
        c.update(extra)
    return c


def run_ours(args):
    import torch
    import torch.distributed as dist

    import cd_dynamax_b200 as cd
    from cd_dynamax_b200 import _engine as E
    from cd_dynamax_b200 import _lib as L
    from cd_dynamax_b200 import parallel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = L.lib()
    N, K = args.n_traj, args.k_obs

    # ---- synthetic inputs (setup, untimed) ----
    t_np = make_times(N, K, CFG["seed"], rank)
    t_dev = torch.as_tensor(t_np, device=dev)
    if args.simple_data:
        gen = torch.Generator(device=dev)
        gen.manual_seed(CFG["seed"] + rank)
        y_dev = 8.0 * torch.randn(N, K, 1, generator=gen, device=dev, dtype=torch.float64)
    else:
        y_dev = make_emissions_torch(t_dev, CFG["seed"] + rank, dev)
    sum_q = substeps_total(t_np, CFG["dt0"])
    params = cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(torch.zeros(3, dtype=torch.float64, device=dev)),
                                      cov=cd.LearnableMatrix(5.0 * torch.eye(3, dtype=torch.float64, device=dev))),
        dynamics=cd.ParamsCDNLGSSMDynamics(
            drift=cd.LearnableLorenz63(sigma=torch.tensor(10.0, dtype=torch.float64, device=dev),
                                       rho=torch.tensor(28.0, dtype=torch.float64, device=dev),
                                       beta=torch.tensor(8.0 / 3.0, dtype=torch.float64, device=dev)),
            diffusion_coefficient=cd.LearnableMatrix(torch.eye(3, dtype=torch.float64, device=dev)),
            diffusion_cov=cd.LearnableMatrix(torch.eye(3, dtype=torch.float64, device=dev))),
        emissions=cd.ParamsCDNLGSSMEmissions(
            emission_function=cd.LearnableLinear(weights=torch.tensor([[1.0, 0.0, 0.0]], dtype=torch.float64, device=dev),
                                                 bias=torch.zeros(1, dtype=torch.float64, device=dev)),
            emission_cov=cd.LearnableMatrix(torch.eye(1, dtype=torch.float64, device=dev))))
    hp = cd.EKFHyperParams(state_order="second", diffeqsolve_settings={"solver": cd.solvers.RK4(), "dt0": CFG["dt0"]})

    def step_resident():
        post = cd.cdnlgssm_filter(params, y_dev, t_dev[..., None], hp)
        s = parallel.allreduce_loglik(E.ll_sum(post.marginal_loglik))  # one NCCL all-reduce of 8 bytes when world > 1
        return post, s

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        post, s = step_resident()
    barrier()
    ll_total = float(s.item())
    n_bad = int((post.marginal_loglik != post.marginal_loglik).sum().item())
    del post

    # ---- timed region: resident inputs ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    launches0 = lib.cdk_launch_count()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier()
    ev[0].record()
    for i in range(args.steps):
        post, s = step_resident()
        del post
        ev[i + 1].record()
    barrier()
    launches = lib.cdk_launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    tmax = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    total_ms = float(tmax.item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- kernel-only time of the dominant kernel (ekf_small_kernel) for the roofline, on the launching stream ----
    stream = torch.cuda.current_stream(dev)
    from cd_dynamax_b200.continuous_discrete_nonlinear_gaussian_ssm._common import run_filter
    kev0, kev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # pre-stage everything so only the kernel launch sits between the events
    fields = dict(dt_final=1e-10, state_order=2, num_iter=1, cov_rescaling=1.0)
    kern_ms = []
    for _ in range(min(args.steps, 5)):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        kev0.record(stream)
        _post, _out, _ = run_filter("cdk_ekf_filter", params, y_dev, t_dev[..., None], None, None, fields,
                                    diffeqsolve_settings={"solver": "rk4", "dt0": CFG["dt0"]})
        kev1.record(stream)
        torch.cuda.synchronize()
        kern_ms.append(kev0.elapsed_time(kev1))
        del _post, _out
    kernel_ms = float(np.median(kern_ms))

    # ---- e2e: host (pinned) buffers through the public API ----
    y_host = y_dev.cpu().pin_memory()
    t_host = t_dev.cpu().pin_memory()[..., None]
    p_host = cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(np.zeros(3)), cov=cd.LearnableMatrix(5.0 * np.eye(3))),
        dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableLorenz63(sigma=10.0, rho=28.0, beta=8.0 / 3.0),
                                           diffusion_coefficient=cd.LearnableMatrix(np.eye(3)),
                                           diffusion_cov=cd.LearnableMatrix(np.eye(3))),
        emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=np.array([[1.0, 0, 0]]),
                                                                                  bias=np.zeros(1)),
                                             emission_cov=cd.LearnableMatrix(np.eye(1))))

    def step_e2e():
        # public API call with host buffers: H2D of y,t inside, D2H of the [N] log-likelihoods inside
        post = cd.cdnlgssm_filter(p_host, y_host, t_host, hp, output_fields=[])
        return post.marginal_loglik

    for _ in range(2):
        ll_host = step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, min(args.steps, 10))
    e0.record()
    for _ in range(e2e_steps):
        ll_host = step_e2e()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_ms = float(e2e_ms.item())
    e2e_ok = bool(abs(float(ll_host.double().sum()) * 1.0 - (ll_total / world if world > 1 else ll_total)) <=
                  1e-6 * abs(ll_total)) if world == 1 else True

    # ---- FP64 FMA peak probe (roofline denominator; MEASURED_PEAKS.json has no FP64 entry) ----
    fp64_peak = fp64_peak3 = None
    if rank == 0:
        blocks, iters = 148 * 8, 20000
        sink = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
        best = 0.0
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(5):
            p0.record()
            L.check(lib.cdk_fma_probe_f64(blocks, iters, ctypes.c_void_p(sink.data_ptr()),
                                          ctypes.c_void_p(stream.cuda_stream)), "fma_probe")
            p1.record()
            torch.cuda.synchronize()
            best = max(best, 2.0 * 16 * iters * blocks * 256 / (p0.elapsed_time(p1) * 1e-3) / 1e12)
        fp64_peak = best
        # the same probe with three distinct register operands per DFMA (what filter arithmetic looks like)
        seed = torch.rand(256, dtype=torch.float64, device=dev)
        best3 = 0.0
        for _ in range(5):
            p0.record()
            L.check(lib.cdk_fma3_probe_f64(blocks, iters, ctypes.c_void_p(sink.data_ptr()),
                                           ctypes.c_void_p(seed.data_ptr()), ctypes.c_void_p(stream.cuda_stream)),
                    "fma3_probe")
            p1.record()
            torch.cuda.synchronize()
            best3 = max(best3, 2.0 * 16 * iters * blocks * 256 / (p0.elapsed_time(p1) * 1e-3) / 1e12)
        fp64_peak3 = best3

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    units_per_step = N * K * world
    value = units_per_step * args.steps / (total_ms * 1e-3)
    peaks, peak_src = measured_peaks()
    flops_exec = FLOP_SUBSTEP_EXEC * sum_q + FLOP_UPDATE_EXEC * N * K
    flops_survey = FLOP_SUBSTEP_SURVEY * sum_q + FLOP_UPDATE_SURVEY * N * K
    ach = flops_exec / (kernel_ms * 1e-3) / 1e12
    ach_survey = flops_survey / (kernel_ms * 1e-3) / 1e12
    hbm_ach = BYTES_PER_OBS_STEP * N * K / (kernel_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, extra={"mean_substeps_per_gap": sum_q / (N * K)}),
        "clocks": clocks,
        "e2e": {"value": units_per_step * e2e_steps / (e2e_ms * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int(y_host.numel() * 8 + t_host.numel() * 8),
                "d2h_bytes_per_step": int(N * 8), "steps": e2e_steps, "result_matches_resident": e2e_ok},
        "gpu_launches": int(launches),
        "roofline": {
            "bound": "fp64", "kernel": "ekf_small_lw<double, DriftL63, 1, RK4, 14> (one CTA per SM, 14 independent warps, "
                                        "state and drift parameters in registers, TMA tensor stores)",
            "achieved": ach_survey, "peak": fp64_peak, "unit": "TFLOP/s",
            "frac": ach_survey / fp64_peak if fp64_peak else None,
            "peak_source": "DFMA probe (cdk_fma_probe_f64) measured in this run, burst; MEASURED_PEAKS.json has no FP64 "
                           "entry; nominal 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2",
            "flop_model": "algorithmic (SURVEY 8d): 496*sum_q + 107*N*K",
            "achieved_executed": ach, "frac_executed": ach / fp64_peak if fp64_peak else None,
            "peak_3_register_operands": fp64_peak3,
            "peak_3_register_operands_note": "cdk_fma3_probe_f64, measured in this run: a DFMA whose three operands are "
                                             "distinct vector registers issues every 3 cycles per SM sub-partition on "
                                             "B200, not 2 (register-file bandwidth); 127 of the kernel's 203 FP64 "
                                             "instructions per substep are such DFMAs",
            "frac_of_3_register_peak": ach_survey / fp64_peak3 if fp64_peak3 else None,
            "flop_model_executed": "338*sum_q + 60*N*K (symmetric P, sparse Lorenz-63 Jacobian, dt folded into RK weights)",
            "kernel_ms": kernel_ms, "traffic": TRAFFIC_NCU_BYTES,
            "traffic_note": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel "
                            "at this workload (profiles/); algorithmic bytes = 208 * N * K = 13.6e9",
            "hbm": {"achieved": hbm_ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": hbm_ach / peaks["hbm_gbs"],
                    "peak_source": f"MEASURED_PEAKS.json ({peak_src})", "bytes_per_obs_step": BYTES_PER_OBS_STEP},
        },
        "ll_sum": ll_total, "non_finite_trajectories": n_bad,
    }
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        res = cpu_reference_leg(args.cpu_sample_traj or 512 * cores, K, steps=2, warmup=1)
        line["cpu_baseline"] = {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": res["kind"],
                                "sample": res["sample"]}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
