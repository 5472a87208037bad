"""Generate golden vectors by running the REFERENCE'S OWN hot-path code (``/root/reference/src``), unmodified, on the
NumPy stand-ins in ``refshim.py`` for jax / diffrax / tensorflow-probability (absent from this image, see
SURVEY.md F2-F3).  Run in the build container only:

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

The reference functions are single-trajectory; each case runs them on a few trajectories and stacks the outputs along
a leading N axis (what ``jax.vmap`` would produce, ``src/ssm_temissions.py:555-567``).
Reference entry points exercised:
  cdlgssm_filter / cdlgssm_smoother            src/continuous_discrete_linear_gaussian_ssm/inference.py:555-823
  cdnlgssm_filter / cdnlgssm_smoother          src/continuous_discrete_nonlinear_gaussian_ssm/models.py:658-764
  (-> extended_kalman_filter / _smoother, unscented_kalman_filter)
"""
import os
import sys
from typing import NamedTuple

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

refshim.install()
jnp = refshim.jnp
import diffrax as dfx  # noqa: E402  (the shim)
from continuous_discrete_linear_gaussian_ssm.inference import (  # noqa: E402
    KFHyperParams, ParamsCDLGSSM, ParamsCDLGSSMDynamics, cdlgssm_filter, cdlgssm_smoother)
from continuous_discrete_nonlinear_gaussian_ssm.cdnlgssm_utils import (  # noqa: E402
    LearnableLinear, LearnableLorenz63, LearnableMatrix, LearnableVector, ParamsCDNLGSSM, ParamsCDNLGSSMDynamics,
    ParamsCDNLGSSMEmissions)
from continuous_discrete_nonlinear_gaussian_ssm.inference_ekf import EKFHyperParams  # noqa: E402
from continuous_discrete_nonlinear_gaussian_ssm.inference_ukf import UKFHyperParams  # noqa: E402
from continuous_discrete_nonlinear_gaussian_ssm.models import cdnlgssm_filter, cdnlgssm_smoother  # noqa: E402
from dynamax.linear_gaussian_ssm.inference import ParamsLGSSMEmissions, ParamsLGSSMInitial  # noqa: E402

SOLVERS = {"euler": dfx.Euler, "heun": dfx.Heun, "midpoint": dfx.Midpoint, "ralston": dfx.Ralston,
           "bosh3": dfx.Bosh3, "rk4": dfx.Rk4, "dopri5": dfx.Dopri5}


def settings(solver, dt0):
    if solver is None:
        return {}
    return {"solver": SOLVERS[solver](), "dt0": dt0}


class LearnableLorenz96(NamedTuple):
    """Not in the reference: the drift BASELINE configs 4-5 name, written as a LearnableFunction."""
    forcing: float

    def f(self, x, u=None, t=None):
        return (jnp.roll(x, -1) - jnp.roll(x, 2)) * jnp.roll(x, 1) - x + self.forcing


class LearnableQuadratic(NamedTuple):
    """f_i = a_i + sum_j B_ij x_j + sum_jk C_ijk x_j x_k -- has a NON-zero 'second order' term, to pin the
    reference's trace(H_t @ P) axis quirk (inference_ekf.py:111-114)."""
    a: np.ndarray
    B: np.ndarray
    C: np.ndarray

    def f(self, x, u=None, t=None):
        return self.a + self.B @ x + jnp.einsum("ijk,j,k->i", self.C, x, x)


def irregular_times(rng, N, K, mean_gap):
    gaps = mean_gap * rng.uniform(0.5, 1.5, size=(N, K))
    gaps[:, 0] = 0.0
    return np.cumsum(gaps, axis=1)


def stack(dicts):
    return {k: np.stack([d[k] for d in dicts], axis=0) for k in dicts[0]}


def post_to_dict(post):
    out = {}
    for k in post._fields:
        v = getattr(post, k)
        if v is not None:
            out[k] = np.asarray(v)
    return out


# ---------------------------------------------------------------------------------------------------------------------
def linear_case(name, seed, n, m, N, K, mean_gap, solver, dt0, dt_final, d_u=0, bias=True, regular=False, tracking=False,
                diag_R=False):
    rng = np.random.default_rng(seed)
    if tracking:  # cdlgssm_tracking.ipynb:136-225 parameters (BASELINE config 1)
        F = np.zeros((4, 4)); F[0, 2] = F[1, 3] = 1.0
        L, Qc = 1e-3 * np.eye(4), np.eye(4)
        H = np.array([[1.0, 0, 0, 0], [0, 1.0, 0, 0]])
        R = 0.5 * np.eye(2)
        m0, P0 = np.array([8.0, 10.0, 1.0, 0.0]), 0.1 * np.eye(4)
        b, d = np.zeros(4), np.zeros(2)
    else:
        G = rng.standard_normal((n, n))
        F = -0.5 * np.eye(n) + 0.3 * G / np.sqrt(n)
        L = np.eye(n) + 0.1 * rng.standard_normal((n, n))
        q = rng.standard_normal((n, n)); Qc = 0.1 * (q @ q.T / n + np.eye(n))
        H = rng.standard_normal((m, n))
        r = rng.standard_normal((m, m)); R = 0.1 * (r @ r.T / m + np.eye(m))
        m0 = rng.standard_normal(n)
        p0 = rng.standard_normal((n, n)); P0 = p0 @ p0.T / n + 0.5 * np.eye(n)
        b = 0.1 * rng.standard_normal(n) if bias else np.zeros(n)
        d = 0.1 * rng.standard_normal(m) if bias else np.zeros(m)
    if diag_R:  # 1-D emissions.cov: the Woodbury branch of _condition_on (cd_linear/inference.py:240-254)
        R = 0.1 * (1.0 + rng.uniform(size=m))
    B = rng.standard_normal((n, d_u)) if d_u else np.zeros((n, 0))
    D = rng.standard_normal((m, d_u)) if d_u else np.zeros((m, 0))
    t = np.tile(np.arange(K, dtype=np.float64), (N, 1)) if regular else irregular_times(rng, N, K, mean_gap)
    y = rng.standard_normal((N, K, m)) + (np.array([8.0, 10.0]) if tracking else 0.0)
    u = rng.standard_normal((N, K, d_u)) if d_u else None
    params = ParamsCDLGSSM(
        initial=ParamsLGSSMInitial(mean=m0, cov=P0),
        dynamics=ParamsCDLGSSMDynamics(weights=F, bias=b, input_weights=B, diffusion_coefficient=L, diffusion_cov=Qc),
        emissions=ParamsLGSSMEmissions(weights=H, bias=d, input_weights=D, cov=R))
    hp = KFHyperParams(dt_final=dt_final, diffeqsolve_settings=settings(solver, dt0))
    f_out, s1_out, s2_out = [], [], []
    for i in range(N):
        ui = None if u is None else u[i]
        f_out.append(post_to_dict(cdlgssm_filter(params, y[i], t[i][:, None], hp, ui)))
        s1 = post_to_dict(cdlgssm_smoother(params, y[i], t[i][:, None], hp, ui, smoother_type="cd_smoother_1"))
        s1_out.append({k: s1[k] for k in ("smoothed_means", "smoothed_covariances", "smoothed_cross_covariances")})
        if d_u == 0:
            s2 = post_to_dict(cdlgssm_smoother(params, y[i], t[i][:, None], hp, ui, smoother_type="cd_smoother_2"))
            s2_out.append({k: s2[k] for k in ("smoothed_means", "smoothed_covariances")})
    save = dict(m0=m0, P0=P0, F=F, L=L, Qc=Qc, H=H, R=R, b=b, d=d, B=B, D=D, y=y, t=t,
                solver=np.array(solver or "dopri5"), dt0=np.array(dt0 if solver else 0.01), dt_final=np.array(dt_final))
    if u is not None:
        save["u"] = u
    save.update({"filt_" + k: v for k, v in stack(f_out).items()})
    save.update({"s1_" + k: v for k, v in stack(s1_out).items()})
    if s2_out:
        save.update({"s2_" + k: v for k, v in stack(s2_out).items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(name, "ll", save["filt_marginal_loglik"])


def nonlinear_case(name, seed, drift_kind, n, m, N, K, mean_gap, solver, dt0, dt_final=1e-10, algo="ekf",
                   state_order="second", num_iter=1, smoother=False, cov_rescaling=1.0):
    rng = np.random.default_rng(seed)
    extra = {}
    if drift_kind == "lorenz63":
        theta = np.array([10.0, 28.0, 8.0 / 3.0])
        drift = LearnableLorenz63(sigma=theta[0], rho=theta[1], beta=theta[2])
        x_scale, m0 = 5.0, np.array([1.0, 1.0, 20.0])
    elif drift_kind == "lorenz96":
        theta = np.array([8.0])
        drift = LearnableLorenz96(forcing=8.0)
        x_scale, m0 = 2.0, 8.0 + 0.5 * rng.standard_normal(n)
    elif drift_kind == "linear":
        G = rng.standard_normal((n, n))
        W = -0.5 * np.eye(n) + 0.3 * G / np.sqrt(n)
        bias = 0.2 * rng.standard_normal(n)
        drift = LearnableLinear(weights=W, bias=bias)
        theta = np.concatenate([W.ravel(), bias])
        x_scale, m0 = 1.0, rng.standard_normal(n)
    elif drift_kind == "quadratic":
        a = 0.3 * rng.standard_normal(n)
        Bm = -0.8 * np.eye(n) + 0.2 * rng.standard_normal((n, n))
        C = 0.05 * rng.standard_normal((n, n, n))
        drift = LearnableQuadratic(a=a, B=Bm, C=C)
        theta = np.concatenate([a, Bm.ravel(), C.ravel()])
        x_scale, m0 = 1.0, 0.3 * rng.standard_normal(n)
    else:
        raise ValueError(drift_kind)
    L = np.eye(n) + 0.05 * rng.standard_normal((n, n))
    Qc = 0.2 * np.eye(n) + 0.02 * np.ones((n, n))
    H = np.zeros((m, n)); H[np.arange(m), (np.arange(m) * max(n // m, 1)) % n] = 1.0
    H = H + 0.05 * rng.standard_normal((m, n))
    d = 0.1 * rng.standard_normal(m)
    r = rng.standard_normal((m, m)); R = 0.5 * (r @ r.T / m + np.eye(m))
    p0 = rng.standard_normal((n, n)); P0 = 0.5 * (p0 @ p0.T / n + np.eye(n))
    t = irregular_times(rng, N, K, mean_gap)
    y = x_scale * rng.standard_normal((N, K, m)) + (H @ m0 + d)
    params = ParamsCDNLGSSM(
        initial=ParamsLGSSMInitial(mean=LearnableVector(params=m0), cov=LearnableMatrix(params=P0)),
        dynamics=ParamsCDNLGSSMDynamics(drift=drift, diffusion_coefficient=LearnableMatrix(params=L),
                                        diffusion_cov=LearnableMatrix(params=Qc), approx_order=1.0),
        emissions=ParamsCDNLGSSMEmissions(emission_function=LearnableLinear(weights=H, bias=d),
                                          emission_cov=LearnableMatrix(params=R)))
    if algo == "ekf":
        hp = EKFHyperParams(dt_final=dt_final, state_order=state_order, cov_rescaling=cov_rescaling,
                            diffeqsolve_settings=settings(solver, dt0))
    else:
        hp = UKFHyperParams(dt_final=dt_final, diffeqsolve_settings=settings(solver, dt0))
    fields = ["filtered_means", "filtered_covariances", "predicted_means", "predicted_covariances"]
    f_out, c_out, s_out = [], [], []
    for i in range(N):
        f_out.append(post_to_dict(cdnlgssm_filter(params, y[i], t[i][:, None], hp, None, num_iter, fields)))
        # "marginal_loglik" in output_fields -> the per-step cumulative array replaces the scalar
        cum = cdnlgssm_filter(params, y[i], t[i][:, None], hp, None, num_iter, ["marginal_loglik"])
        c_out.append({"marginal_loglik_cumulative": np.asarray(cum.marginal_loglik)})
        if smoother:
            s = post_to_dict(cdnlgssm_smoother(params, y[i], t[i][:, None], hp, None, num_iter))
            s_out.append({k: s[k] for k in ("smoothed_means", "smoothed_covariances")})
    save = dict(m0=m0, P0=P0, theta=theta, L=L, Qc=Qc, H=H, R=R, d=d, y=y, t=t, drift=np.array(drift_kind),
                algo=np.array(algo), state_order=np.array(state_order), num_iter=np.array(num_iter),
                cov_rescaling=np.array(cov_rescaling),
                solver=np.array(solver or "dopri5"), dt0=np.array(dt0 if solver else 0.01), dt_final=np.array(dt_final))
    save.update({"filt_" + k: v for k, v in stack(f_out).items()})
    save.update({"filt_" + k: v for k, v in stack(c_out).items()})
    if s_out:
        save.update({"smooth_" + k: v for k, v in stack(s_out).items()})
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **save)
    print(name, "ll", save["filt_marginal_loglik"])


def pushforward_constants():
    """The two constants the reference hard-codes (cdlgssm_test_filter_TRegular.py:61-62) reproduced by running the
    reference's compute_pushforward on the shims in float32 and float64."""
    from continuous_discrete_linear_gaussian_ssm.inference import compute_pushforward
    out = {}
    for dt in (np.float32, np.float64):
        refshim.set_default_dtype(dt)
        params = ParamsCDLGSSM(
            initial=ParamsLGSSMInitial(mean=np.zeros(1, dt), cov=np.eye(1, dtype=dt)),
            dynamics=ParamsCDLGSSMDynamics(weights=np.array([[-0.1]], dt), bias=None, input_weights=None,
                                           diffusion_coefficient=np.eye(1, dtype=dt),
                                           diffusion_cov=np.array([[0.125]], dt)),
            emissions=ParamsLGSSMEmissions(weights=None, bias=None, input_weights=None, cov=None))
        A, Q = compute_pushforward(params, dt(0.0), dt(1.0))
        out[f"A_{np.dtype(dt).name}"], out[f"Q_{np.dtype(dt).name}"] = A, Q
    refshim.set_default_dtype(np.float64)
    out["A_reference_constant"] = np.float32(0.9048373699188232421875)
    out["Q_reference_constant"] = np.float32(0.11329327523708343505859375)
    np.savez_compressed(os.path.join(HERE, "pushforward_constants.npz"), **out)
    print("pushforward", {k: repr(np.asarray(v).ravel()[0]) for k, v in out.items()})


if __name__ == "__main__":
    pushforward_constants()
    # --- linear CD-KF + smoothers
    linear_case("kf_tracking_c1", 1234, 4, 2, N=2, K=40, mean_gap=0.05, solver=None, dt0=None, dt_final=1.0, tracking=True)
    linear_case("kf_random_rk4", 11, 3, 2, N=3, K=30, mean_gap=0.07, solver="rk4", dt0=0.02, dt_final=1e-10)
    linear_case("kf_inputs_heun", 12, 3, 2, N=2, K=25, mean_gap=0.05, solver="heun", dt0=0.01, dt_final=0.3, d_u=2)
    linear_case("kf_regular_dopri5", 13, 2, 6, N=2, K=8, mean_gap=1.0, solver=None, dt0=None, dt_final=1.0, regular=True, bias=False)
    linear_case("kf_n16_rk4", 14, 16, 4, N=2, K=12, mean_gap=0.04, solver="rk4", dt0=0.01, dt_final=1e-10)
    linear_case("kf_diagR_rk4", 15, 5, 3, N=3, K=25, mean_gap=0.05, solver="rk4", dt0=0.0125, dt_final=0.02, diag_R=True)
    linear_case("kf_diagR_inputs_dopri5", 16, 4, 2, N=2, K=15, mean_gap=0.05, solver=None, dt0=None, dt_final=1e-10, d_u=1,
                diag_R=True)
    # --- CD-EKF / EKS
    nonlinear_case("ekf_l63_second_rk4", 21, "lorenz63", 3, 1, N=3, K=60, mean_gap=0.01, solver="rk4", dt0=0.0025, smoother=True)
    nonlinear_case("ekf_l63_first_dopri5", 22, "lorenz63", 3, 2, N=2, K=30, mean_gap=0.03, solver=None, dt0=None, state_order="first", smoother=True)
    nonlinear_case("ekf_l63_zeroth_euler", 23, "lorenz63", 3, 1, N=2, K=30, mean_gap=0.01, solver="euler", dt0=0.002, state_order="zeroth", cov_rescaling=0.7)
    nonlinear_case("ekf_l63_iter2_bosh3", 24, "lorenz63", 3, 3, N=2, K=30, mean_gap=0.01, solver="bosh3", dt0=0.004, num_iter=2, dt_final=0.01)
    nonlinear_case("ekf_linear_rk4", 25, "linear", 4, 2, N=2, K=30, mean_gap=0.05, solver="rk4", dt0=0.0125, smoother=True)
    nonlinear_case("ekf_quadratic_second_rk4", 26, "quadratic", 3, 2, N=2, K=40, mean_gap=0.02, solver="rk4", dt0=0.005)
    nonlinear_case("ekf_l96_rk4", 27, "lorenz96", 8, 4, N=2, K=30, mean_gap=0.02, solver="rk4", dt0=0.005, smoother=True)
    # --- CD-UKF
    nonlinear_case("ukf_l63_rk4", 31, "lorenz63", 3, 1, N=3, K=50, mean_gap=0.01, solver="rk4", dt0=0.0025, algo="ukf")
    nonlinear_case("ukf_linear_dopri5", 32, "linear", 4, 2, N=2, K=20, mean_gap=0.05, solver=None, dt0=None, algo="ukf")
    nonlinear_case("ukf_l96_rk4", 33, "lorenz96", 10, 5, N=2, K=30, mean_gap=0.02, solver="rk4", dt0=0.005, algo="ukf")
