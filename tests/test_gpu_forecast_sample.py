"""GPU: the callers either side of the filters against the oracle:
cdnlgssm_forecast (EKF / UKF / EnKF _predict scanned with no updates; point estimates -> one SDE sample path),
cdnlgssm_path_sample / sample_batch(transition_type="path"), cdnlgssm_emissions.
Deterministic forecasts: the filters' gates (log-likelihood-free; moments 1e-9 / 1e-10).  Sampled paths: 1e-8 against the
oracle on the shared Philox stream (libm vs CUDA log / sincos in the last ulp of every normal, as for the EnKF)."""
import numpy as np
import pytest

from oracle import cd_oracle as o
from tests.helpers import gate_err, moment_norm_err, record, scaled_err
from tests.test_gpu_parity import api, nonlinear_params_api

pytestmark = pytest.mark.gpu

L63 = dict(m0=np.array([1.0, 1.0, 20.0]), P0=2.0 * np.eye(3), drift="lorenz63", theta=np.array([10.0, 28.0, 8.0 / 3.0]),
           L=np.eye(3) + 0.1 * np.arange(9).reshape(3, 3) / 9, Qc=0.5 * np.eye(3) + 0.05, H=np.array([[1.0, 0.3, -0.2]]),
           R=0.7 * np.eye(1), d=np.array([0.1]))


def _l96(n=8, m=4, seed=0):
    rng = np.random.default_rng(seed)
    return dict(m0=8.0 + rng.standard_normal(n), P0=np.eye(n), drift="lorenz96", theta=np.array([8.0]), L=np.eye(n),
                Qc=0.1 * np.eye(n), H=np.eye(n)[::2][:m], R=np.eye(m), d=np.zeros(m))


def _oracle_params(g):
    n = g["m0"].shape[-1]
    drift = o.Lorenz63Drift(*g["theta"]) if g["drift"] == "lorenz63" else o.Lorenz96Drift(g["theta"][0])
    return o.NonlinearParams(m0=g["m0"], P0=g["P0"], drift=drift, L=g["L"], Qc=g["Qc"], H=g["H"], R=g["R"], d=g["d"])


def _times(N, K, gap, seed):
    rng = np.random.default_rng(seed)
    return 0.3 + np.cumsum(gap * rng.uniform(0.5, 1.5, (N, K)), axis=1)


@pytest.mark.parametrize("model,algo,solver,dt0", [("l63", "ekf", "rk4", 0.0025), ("l63", "ekf", "dopri5", 0.01),
                                                    ("l96", "ekf", "rk4", 0.005), ("l63", "ukf", "rk4", 0.0025),
                                                    ("l96", "ukf", "rk4", 0.005), ("l96", "ukf_sigma", "rk4", 0.005)])
def test_forecast_of_a_gaussian_vs_oracle(model, algo, solver, dt0, monkeypatch):
    cd = api()
    monkeypatch.setenv("CDK_UKF_SIGMA_POINTS", "1" if algo == "ukf_sigma" else "0")
    g = L63 if model == "l63" else _l96()
    N, K = 37, 15
    rng = np.random.default_rng(3)
    m0 = g["m0"][None] + 0.3 * rng.standard_normal((N, g["m0"].shape[0]))  # one initial Gaussian per trajectory
    tf = _times(N, K, 0.01 if model == "l63" else 0.02, 5)
    p = nonlinear_params_api(g)
    st = {"solver": solver, "dt0": dt0}
    hp = cd.EKFHyperParams(diffeqsolve_settings=st) if algo == "ekf" else cd.UKFHyperParams(diffeqsolve_settings=st)
    fc = cd.cdnlgssm_forecast(p, cd.MultivariateNormalFullCovariance(m0, g["P0"]), 0.3, tf[..., None], hp)
    assert fc.forecasted_state_means.shape == (N, K, m0.shape[1]) and fc.forecasted_state_path is None
    po = _oracle_params(dict(g, m0=m0))
    T = np.concatenate([np.full((N, 1), 0.3), tf], axis=1)
    ydummy = np.zeros((N, K, g["H"].shape[0]))
    fn = o.extended_kalman_filter if algo == "ekf" else o.unscented_kalman_filter
    r = fn(po, ydummy, T, settings=o.SolverSettings(solver, dt0), forecast=True)
    # An unobserved chaotic forecast lets the covariance grow until an RK STAGE covariance of a trajectory or two stops being
    # positive definite: the reference (and the oracle, and the literal sigma-point kernel -- compared NaN for NaN below)
    # then return NaN from chol(P_stage).  The closed-form unscented predict never factors P, so it keeps returning the
    # moments of the same ODE there: compared on the trajectories the reference completes (documented in DESIGN.md).
    ok = np.isfinite(r["predicted_means"]).all(axis=(1, 2))
    assert ok.sum() >= N - 3
    if algo != "ukf":
        ok[:] = True
    for fld, ref, core in (("forecasted_state_means", "predicted_means", 1), ("forecasted_state_covariances", "predicted_covariances", 2)):
        e, en = gate_err(getattr(fc, fld)[ok], r[ref][ok]), moment_norm_err(getattr(fc, fld)[ok], r[ref][ok], core)
        record(f"forecast_{model}_{algo}_{solver}:{fld}:gate", e)
        assert e < 1e-9 and en < 1e-10, (fld, e, en)
    # unbatched call: [K, 1] times, one Gaussian; output_fields selects
    f1 = cd.cdnlgssm_forecast(p, (m0[0], g["P0"]), np.array([[0.3]]), tf[0][:, None], hp,
                              output_fields=["forecasted_state_covariances"])
    assert f1.forecasted_state_means is None and f1.forecasted_state_covariances.shape == (K,) + g["P0"].shape
    assert scaled_err(f1.forecasted_state_covariances, fc.forecasted_state_covariances[0]) < 1e-12


def test_forecast_enkf_vs_oracle():
    cd = api()
    g = _l96()
    N, K, E = 3, 8, 96
    tf = _times(N, K, 0.02, 6)
    hp = cd.EnKFHyperParams(N_particles=E, key=4321, diffeqsolve_settings={"solver": "euler", "dt0": 0.005})
    fc = cd.cdnlgssm_forecast(nonlinear_params_api(g), (np.repeat(g["m0"][None], N, 0), g["P0"]), 0.3, tf[..., None], hp)
    T = np.concatenate([np.full((N, 1), 0.3), tf], axis=1)
    r = o.ensemble_kalman_filter(_oracle_params(g), np.zeros((N, K, 4)), T, E=E, seed=4321,
                                 settings=o.SolverSettings("euler", 0.005), forecast=True)
    assert scaled_err(fc.forecasted_state_means, r["predicted_means"]) < 1e-8
    assert scaled_err(fc.forecasted_state_covariances, r["predicted_covariances"]) < 1e-8


@pytest.mark.parametrize("model,solver", [("l63", "heun"), ("l63", "euler"), ("l96", "heun")])
def test_path_sampler_vs_oracle(model, solver):
    cd = api()
    g = L63 if model == "l63" else _l96()
    N, K = 301, 25
    t = _times(N, K, 0.01 if model == "l63" else 0.02, 8)
    dt0 = 0.0025 if model == "l63" else 0.005
    p = nonlinear_params_api(g)
    xs, ys = cd.cdnlgssm_path_sample(p, 99, K, t[..., None], diffeqsolve_settings={"solver": solver, "dt0": dt0})
    assert xs.shape == (N, K, g["m0"].shape[0]) and ys.shape == (N, K, g["H"].shape[0])
    rx, ry = o.sample_paths(_oracle_params(g), t, seed=99, settings=o.SolverSettings(solver, dt0))
    ex, ey = scaled_err(xs, rx), scaled_err(ys, ry)
    record(f"sampler_{model}_{solver}:states", ex)
    assert ex < 1e-8 and ey < 1e-8, (ex, ey)
    # sample_batch with a shared time grid = the vmapped reference call; default solver for an SDE is Heun
    model_obj = cd.ContDiscreteNonlinearGaussianSSM(state_dim=g["m0"].shape[0], emission_dim=g["H"].shape[0],
                                                     diffeqsolve_settings={"dt0": dt0})
    xb, yb = model_obj.sample_batch(p, 7, 50, K, t[0][:, None], transition_type="path")
    rb = o.sample_paths(_oracle_params(g), np.repeat(t[:1], 50, 0), seed=7, settings=o.SolverSettings("heun", dt0))
    assert xb.shape == (50, K, g["m0"].shape[0]) and scaled_err(xb, rb[0]) < 1e-8 and scaled_err(yb, rb[1]) < 1e-8
    x1, y1 = model_obj.sample(p, 7, K, t[0][:, None], transition_type="path")
    assert x1.shape == (K, g["m0"].shape[0]) and np.array_equal(x1, xb[0])
    with pytest.raises(NotImplementedError):
        model_obj.sample(p, 7, K, t[0][:, None])  # the Gaussian-transition sampler is not provided


def test_forecast_of_a_point_estimate_is_a_sample_path():
    cd = api()
    g = L63
    N, K = 20, 12
    tf = _times(N, K, 0.01, 11)
    x0 = g["m0"][None] + np.random.default_rng(2).standard_normal((N, 3))
    fc = cd.cdnlgssm_forecast(nonlinear_params_api(g), x0, 0.3, tf[..., None], key=5,
                              diffeqsolve_settings={"solver": "heun", "dt0": 0.0025})
    assert fc.forecasted_state_means is None and fc.forecasted_state_path.shape == (N, K, 3)
    T = np.concatenate([np.full((N, 1), 0.3), tf], axis=1)
    rx, ry = o.sample_paths(_oracle_params(dict(g, m0=x0)), T, seed=5, settings=o.SolverSettings("heun", 0.0025), fixed_init=True)
    assert scaled_err(fc.forecasted_state_path, rx) < 1e-8 and scaled_err(fc.forecasted_emission_path, ry) < 1e-8


def test_sampler_statistics_of_a_linear_sde():
    """Distributional check (what the reference's own sampler tests can assert): for dx = -a x dt + s dW the sampled paths
    have mean m0 e^{-a t} and variance P0 e^{-2 a t} + s^2 (1 - e^{-2 a t}) / (2 a); emissions add R."""
    cd = api()
    a_, s_, N, K = 0.8, 0.5, 40000, 6
    p = cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(np.array([2.0])), cov=cd.LearnableMatrix(np.array([[0.3]]))),
        dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableLinear(weights=np.array([[-a_]]), bias=np.zeros(1)),
                                           diffusion_coefficient=cd.LearnableMatrix(np.array([[s_]])),
                                           diffusion_cov=cd.LearnableMatrix(np.eye(1))),
        emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=np.eye(1), bias=np.zeros(1)),
                                             emission_cov=cd.LearnableMatrix(np.array([[0.2]]))))
    t = np.linspace(0.0, 1.0, K)
    xs, ys = cd.cdnlgssm_path_sample(p, 123, K, t[:, None], diffeqsolve_settings={"dt0": 0.005}, num_sequences=N)
    mean = 2.0 * np.exp(-a_ * t)
    var = 0.3 * np.exp(-2 * a_ * t) + s_ ** 2 * (1 - np.exp(-2 * a_ * t)) / (2 * a_)
    assert np.allclose(xs[..., 0].mean(0), mean, atol=4 * np.sqrt(var.max() / N) + 2e-3)
    assert np.allclose(xs[..., 0].var(0), var, rtol=0.04)
    assert np.allclose(ys[..., 0].var(0), var + 0.2, rtol=0.04)


def test_emission_moments_vs_oracle():
    cd = api()
    g = _l96()
    rng = np.random.default_rng(1)
    N, K, n, m = 5, 9, 8, 4
    sm = rng.standard_normal((N, K, n))
    A = rng.standard_normal((N, K, n, n))
    sp = A @ np.swapaxes(A, -1, -2) + np.eye(n)
    p = nonlinear_params_api(g)
    em, ec = cd.cdnlgssm_emissions(p, np.zeros((K, 1)), sm, sp, hyperparams=cd.EKFHyperParams())
    rm, rc = o.emission_moments(_oracle_params(g), sm, sp)
    assert em.shape == (N, K, m) and ec.shape == (N, K, m, m)
    assert scaled_err(em, rm) < 1e-13 and scaled_err(ec, rc) < 1e-13
    em2, ec2 = cd.cdnlgssm_emissions(p, np.zeros((K, 1)), sm[0])  # point estimates: the model's own (H m + d, R)
    rm2, rc2 = o.emission_moments(_oracle_params(g), sm[0])
    assert em2.shape == (K, m) and scaled_err(em2, rm2) < 1e-13 and scaled_err(ec2, rc2) < 1e-13


@pytest.mark.parametrize("algo", ["ekf", "ukf"])
def test_long_forecast_l96_n40_stays_symmetric(algo):
    """150 gaps with no update on Lorenz-96 n = 40 (t = 3, ~5 Lyapunov times): the register-resident moment ODE reads P_{c,r} as
    P_{r,c}, i.e. it integrates dP = J P + P J^T, whose antisymmetric part is as unstable as the symmetric one -- every gap
    therefore starts from sym(P).  The forecast covariance must stay symmetric to rounding and follow the oracle's (EKF: gate
    relaxed to 1e-7 -- the unobserved chaotic covariance grows by e^{2 lambda t} and so does the distance between two roundings)."""
    from tests.test_gpu_baseline_shapes import _l96_case
    cd = api()
    N, K = 2, 150
    g, po, t, y = _l96_case(N=N, K=K, seed=8)
    g = dict(g, P0=0.01 * np.eye(40))
    po = o.NonlinearParams(m0=g["m0"], P0=g["P0"], drift=o.Lorenz96Drift(8.0), L=g["L"], Qc=g["Qc"], H=g["H"], R=g["R"], d=g["d"])
    tf = 0.3 + t + 0.02
    st = {"solver": "rk4", "dt0": 0.005}
    hp = cd.EKFHyperParams(diffeqsolve_settings=st) if algo == "ekf" else cd.UKFHyperParams(diffeqsolve_settings=st)
    m0 = np.repeat(g["m0"][None], N, axis=0)
    fc = cd.cdnlgssm_forecast(nonlinear_params_api(g), cd.MultivariateNormalFullCovariance(m0, g["P0"]), 0.3, tf[..., None], hp)
    P = np.asarray(fc.forecasted_state_covariances)
    assert np.isfinite(P).all()
    asym = np.abs(P - np.swapaxes(P, -1, -2)).max(axis=(-1, -2)) / np.abs(P).max(axis=(-1, -2))
    record(f"long_forecast_{algo}:asymmetry", asym.max())
    assert asym.max() < 1e-12, asym.max()
    if algo == "ekf":
        T = np.concatenate([np.full((N, 1), 0.3), tf], axis=1)
        r = o.extended_kalman_filter(po, np.zeros((N, K, 20)), T, settings=o.SolverSettings("rk4", 0.005), forecast=True)
        e = scaled_err(P, r["predicted_covariances"])
        record("long_forecast_ekf:covariances:scaled", e)
        assert e < 1e-7, e
