"""GPU: a USER-DEFINED drift compiled into a variant of libcdk.so at first use
(cd_dynamax_b200.build.build_user_drift: nvcc on the box, cached by content hash) -- a Van der Pol oscillator with a cubic
restoring force (not a polynomial of degree <= 2, so the UKF takes literal sigma points) -- through the EKF (first order and
the reference's 'second' order with its Hessian-trace quirk), the EKS, the UKF, the EnKF, a forecast and the path sampler,
against the oracle with the same drift written in NumPy."""
import os

import numpy as np
import pytest

from oracle import cd_oracle as o
from tests.helpers import max_rel_err, record, scaled_err
from tests.test_gpu_parity import FIELDS, TOL, api, check_moments

pytestmark = pytest.mark.gpu

VDP_CODE = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "cd_dynamax_b200", "examples",
                             "vdp_drift.cuh")).read()


class VdpDrift:
    """The same drift for the oracle."""

    def __init__(self, mu, eps):
        self.mu, self.eps = mu, eps

    def f(self, x):
        x0, x1 = x[..., 0], x[..., 1]
        return np.stack([x1, self.mu * (1 - x0 * x0) * x1 - x0 - self.eps * x0 ** 3], axis=-1)

    def jac(self, x):
        x0, x1 = x[..., 0], x[..., 1]
        J = np.zeros(x.shape + (2,))
        J[..., 0, 1] = 1.0
        J[..., 1, 0] = -2 * self.mu * x0 * x1 - 1 - 3 * self.eps * x0 * x0
        J[..., 1, 1] = self.mu * (1 - x0 * x0)
        return J

    def grad_div(self, x):
        g = np.zeros_like(x)
        g[..., 0] = -2 * self.mu * x[..., 0]
        return g


def _setup(N=9, K=40, seed=0):
    cd = api()
    rng = np.random.default_rng(seed)
    theta = np.array([1.3, 0.4])
    g = dict(m0=np.array([1.0, 0.5]), P0=0.3 * np.eye(2), L=np.eye(2) + 0.1 * rng.standard_normal((2, 2)),
             Qc=0.2 * np.eye(2) + 0.02, H=np.array([[1.0, 0.2]]), R=0.1 * np.eye(1), d=np.array([0.05]))
    p = cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(g["m0"]), cov=cd.LearnableMatrix(g["P0"])),
        dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableUserDrift(device_code=VDP_CODE, theta=theta),
                                           diffusion_coefficient=cd.LearnableMatrix(g["L"]),
                                           diffusion_cov=cd.LearnableMatrix(g["Qc"])),
        emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=g["H"], bias=g["d"]),
                                             emission_cov=cd.LearnableMatrix(g["R"])))
    po = o.NonlinearParams(m0=g["m0"], P0=g["P0"], drift=VdpDrift(*theta), L=g["L"], Qc=g["Qc"], H=g["H"], R=g["R"], d=g["d"])
    t = np.cumsum(0.05 * rng.uniform(0.5, 1.5, (N, K)), axis=1)
    y = 1.5 * rng.standard_normal((N, K, 1))
    return cd, p, po, t, y


@pytest.mark.parametrize("order", ["first", "second"])
def test_user_drift_ekf_and_eks(order):
    cd, p, po, t, y = _setup()
    st = {"solver": "rk4", "dt0": 0.0125}
    hp = cd.EKFHyperParams(state_order=order, dt_final=0.02, diffeqsolve_settings=st)
    f = cd.cdnlgssm_filter(p, y, t[..., None], hp)
    r = o.extended_kalman_filter(po, y, t, dt_final=0.02, state_order=order, settings=o.SolverSettings("rk4", 0.0125))
    assert max_rel_err(f.marginal_loglik, r["marginal_loglik"]) < TOL
    check_moments(f, r, f"user_drift_ekf_{order}")
    if order == "first":
        s = cd.cdnlgssm_smoother(p, y, t[..., None], hp)
        rs = o.extended_kalman_smoother(po, y, t, dt_final=0.02, state_order=order, settings=o.SolverSettings("rk4", 0.0125))
        for fld in ("smoothed_means", "smoothed_covariances"):
            assert scaled_err(getattr(s, fld), rs[fld]) < 1e-8, fld


def test_user_drift_ukf_enkf_forecast_and_sampler():
    cd, p, po, t, y = _setup(N=5, K=25, seed=1)
    st = {"solver": "rk4", "dt0": 0.0125}
    f = cd.cdnlgssm_filter(p, y, t[..., None], cd.UKFHyperParams(diffeqsolve_settings=st))
    r = o.unscented_kalman_filter(po, y, t, settings=o.SolverSettings("rk4", 0.0125))
    assert max_rel_err(f.marginal_loglik, r["marginal_loglik"]) < TOL
    check_moments(f, r, "user_drift_ukf")
    fe = cd.cdnlgssm_filter(p, y, t[..., None], cd.EnKFHyperParams(N_particles=80, key=3, diffeqsolve_settings={"dt0": 0.0125}))
    re = o.ensemble_kalman_filter(po, y, t, E=80, seed=3, settings=o.SolverSettings("heun", 0.0125))
    assert max_rel_err(fe.marginal_loglik, re["marginal_loglik"]) < 1e-8
    for fld in FIELDS:
        assert scaled_err(getattr(fe, fld), re[fld]) < 1e-8, fld
    fc = cd.cdnlgssm_forecast(p, (po.m0, po.P0), 0.0, t[..., None], cd.EKFHyperParams(state_order="first", diffeqsolve_settings=st))
    T = np.concatenate([np.zeros((t.shape[0], 1)), t], axis=1)
    rf = o.extended_kalman_filter(po, np.zeros_like(y), T, state_order="first", settings=o.SolverSettings("rk4", 0.0125), forecast=True)
    assert scaled_err(fc.forecasted_state_covariances, rf["predicted_covariances"]) < 1e-10
    xs, ys = cd.cdnlgssm_path_sample(p, 11, t.shape[1], t[..., None], diffeqsolve_settings={"dt0": 0.0125})
    rx, ry = o.sample_paths(po, t, seed=11, settings=o.SolverSettings("heun", 0.0125))
    assert scaled_err(xs, rx) < 1e-8 and scaled_err(ys, ry) < 1e-8
    record("user_drift_sampler:states", scaled_err(xs, rx))


def test_stock_library_rejects_the_user_drift_id():
    import ctypes
    from cd_dynamax_b200 import _lib as L
    lib = L.lib()
    assert lib.cdk_has_user_drift() == 0
    d = L.new_desc()
    d.N, d.K, d.n, d.m, d.drift_id, d.n_theta = 2, 3, 2, 1, L.DRIFT_USER, 2
    ins = (ctypes.c_void_p * L.NUM_IN)()
    outs = (ctypes.c_void_p * L.NUM_OUT)()
    assert lib.cdk_ekf_filter_f64(ctypes.byref(d), ins, outs, None) == -4  # CDK_E_UNSUPPORTED
