"""GPU: tests written the way the reference's own test scripts are, through the drop-in API.

  * src/test_scripts/cdlgssm_test_filter_TRegular.py:61-62,245 -- on a regular grid the CD Kalman filter equals the
    discrete Kalman filter run with the exact pushforward (A, Q); the reference asserts rtol 1e-5 in fp32.
  * src/test_scripts/cdnlgssm_test_filter_linear_TRegular.py:324,424,434-470 -- on a linear model the CD-EKF (every
    state_order), the CD-UKF and the CD-KF coincide, and the EKS coincides with the type-2 linear smoother.
  * call conventions the notebooks rely on: t_emissions=None, integer time stamps, hyper-parameters None,
    model-class entry points, output_fields, the loud failures for what a CUDA kernel cannot take.
"""
import numpy as np
import pytest
import scipy.linalg

pytestmark = pytest.mark.gpu


def api():
    import cd_dynamax_b200 as cd
    return cd


def _linear_model(n=4, m=2, seed=0):
    rng = np.random.default_rng(seed)
    F = -0.4 * np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    Lm = np.eye(n)
    Qc = 0.2 * np.eye(n) + 0.02
    H = rng.standard_normal((m, n))
    R = 0.3 * np.eye(m)
    return dict(F=F, L=Lm, Qc=Qc, H=H, R=R, m0=rng.standard_normal(n), P0=np.eye(n))


def _lin_params(cd, g, bias=None):
    n, m = g["F"].shape[0], g["H"].shape[0]
    return cd.ParamsCDLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=g["m0"], cov=g["P0"]),
        dynamics=cd.ParamsCDLGSSMDynamics(weights=g["F"], bias=np.zeros(n) if bias is None else bias, input_weights=None,
                                          diffusion_coefficient=g["L"], diffusion_cov=g["Qc"]),
        emissions=cd.ParamsLGSSMEmissions(weights=g["H"], bias=np.zeros(m), input_weights=None, cov=g["R"]))


def _nl_params(cd, g):
    n, m = g["F"].shape[0], g["H"].shape[0]
    return cd.ParamsCDNLGSSM(
        initial=cd.ParamsLGSSMInitial(mean=cd.LearnableVector(g["m0"]), cov=cd.LearnableMatrix(g["P0"])),
        dynamics=cd.ParamsCDNLGSSMDynamics(drift=cd.LearnableLinear(weights=g["F"], bias=np.zeros(n)),
                                           diffusion_coefficient=cd.LearnableMatrix(g["L"]),
                                           diffusion_cov=cd.LearnableMatrix(g["Qc"])),
        emissions=cd.ParamsCDNLGSSMEmissions(emission_function=cd.LearnableLinear(weights=g["H"], bias=np.zeros(m)),
                                             emission_cov=cd.LearnableMatrix(g["R"])))


def _van_loan(F, LQL, dt):
    """Exact (A, Q) of dx = F x dt + L dW over dt (matrix-fraction / Van Loan block exponential)."""
    n = F.shape[0]
    M = np.block([[-F, LQL], [np.zeros((n, n)), F.T]]) * dt
    E = scipy.linalg.expm(M)
    A = E[n:, n:].T
    return A, A @ E[:n, n:]


def _discrete_kf(g, y, A, Q):
    """Textbook discrete Kalman filter with the prior as the prediction at t_0 (dynamax lgssm_filter convention)."""
    m, P = g["m0"].copy(), g["P0"].copy()
    H, R = g["H"], g["R"]
    ll, fm, fP = 0.0, [], []
    for k in range(y.shape[0]):
        S = H @ P @ H.T + R
        r = y[k] - H @ m
        ll += -0.5 * (r @ np.linalg.solve(S, r) + np.linalg.slogdet(S)[1] + len(r) * np.log(2 * np.pi))
        Kg = np.linalg.solve(S, H @ P).T
        m, P = m + Kg @ r, P - Kg @ S @ Kg.T
        fm.append(m.copy()), fP.append(P.copy())
        m, P = A @ m, A @ P @ A.T + Q
    return ll, np.array(fm), np.array(fP)


def test_cdkf_on_regular_grid_equals_discrete_kf():
    cd = api()
    g = _linear_model()
    K, dt = 40, 0.25
    rng = np.random.default_rng(1)
    y = rng.standard_normal((K, 2))
    t = dt * np.arange(K)
    A, Q = _van_loan(g["F"], g["L"] @ g["Qc"] @ g["L"].T, dt)
    ll, fm, fP = _discrete_kf(g, y, A, Q)
    hp = cd.KFHyperParams(dt_final=dt, diffeqsolve_settings={"solver": "dopri5", "dt0": 0.01})
    f = cd.cdlgssm_filter(_lin_params(cd, g), y, t[:, None], hp)
    # Dopri5 with dt0 = 0.01 integrates the pushforward to ~1e-12; the reference's own bar is rtol 1e-5
    assert abs(f.marginal_loglik - ll) < 1e-8 * abs(ll)
    np.testing.assert_allclose(f.filtered_means, fm, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(f.filtered_covariances, fP, rtol=1e-8, atol=1e-10)


def test_ekf_ukf_kf_coincide_on_a_linear_model():
    cd = api()
    g = _linear_model(seed=3)
    N, K = 5, 30
    rng = np.random.default_rng(2)
    gaps = 0.1 * rng.uniform(0.5, 1.5, size=(N, K))
    gaps[:, 0] = 0.0
    t = np.cumsum(gaps, axis=1)
    y = rng.standard_normal((N, K, 2))
    st = {"solver": "rk4", "dt0": 0.01}
    kf = cd.cdlgssm_filter(_lin_params(cd, g), y, t[..., None], cd.KFHyperParams(diffeqsolve_settings=st))
    p = _nl_params(cd, g)
    fields = ("marginal_loglik", "filtered_means", "filtered_covariances", "predicted_means", "predicted_covariances")
    for order in ("first", "second"):
        ekf = cd.cdnlgssm_filter(p, y, t[..., None], cd.EKFHyperParams(state_order=order, diffeqsolve_settings=st))
        for fld in fields:  # the moment ODE of a linear drift IS the pushforward composed with the discrete predict
            np.testing.assert_allclose(getattr(ekf, fld), getattr(kf, fld), rtol=1e-7, atol=1e-9, err_msg=fld)
    ukf = cd.cdnlgssm_filter(p, y, t[..., None], cd.UKFHyperParams(diffeqsolve_settings=st))
    for fld in fields:
        np.testing.assert_allclose(getattr(ukf, fld), getattr(kf, fld), rtol=1e-7, atol=1e-9, err_msg=fld)
    # the EKS on a linear drift is the type-2 (backward ODE) linear smoother -- which always integrates with the default
    # solver (cd_linear/inference.py:688), so run the EKS with the defaults too
    hp_def = cd.EKFHyperParams()
    eks = cd.cdnlgssm_smoother(p, y, t[..., None], hp_def)
    ks2 = cd.cdlgssm_smoother(_lin_params(cd, g), y, t[..., None], cd.KFHyperParams(), smoother_type="cd_smoother_2")
    np.testing.assert_allclose(eks.smoothed_means, ks2.smoothed_means, rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(eks.smoothed_covariances, ks2.smoothed_covariances, rtol=1e-7, atol=1e-9)
    # (type 1 and type 2 do NOT agree beyond O(gap): the reference's backward ODE freezes m_f, P_f at t_k over the whole
    # gap, cd_linear/inference.py:664-684 -- both are pinned separately against the reference's own output in
    # test_gpu_parity.py::test_kf_filter_and_smoothers_vs_reference_golden)


def test_call_conventions_of_the_notebooks():
    cd = api()
    g = _linear_model(seed=5)
    K = 12
    rng = np.random.default_rng(3)
    y = rng.standard_normal((K, 2))
    p = _lin_params(cd, g)
    # t_emissions=None means unit spacing arange(K) (cd_linear/inference.py:590-593); integer stamps are cast
    a = cd.cdlgssm_filter(p, y)
    b = cd.cdlgssm_filter(p, y, np.arange(K, dtype=np.int32)[:, None])
    c = cd.cdlgssm_filter(p, y, np.arange(K, dtype=np.float64)[:, None], None)  # hyper-parameters None = defaults
    for fld in ("marginal_loglik", "filtered_means", "predicted_covariances"):
        assert np.array_equal(getattr(a, fld), getattr(b, fld)) and np.array_equal(getattr(a, fld), getattr(c, fld))
    assert a.filtered_means.shape == (K, 4) and a.filtered_covariances.shape == (K, 4, 4) and np.ndim(a.marginal_loglik) == 0
    # model-class entry points (cd_linear/models.py:336-365)
    model = cd.ContDiscreteLinearGaussianSSM(state_dim=4, emission_dim=2)
    assert model.marginal_log_prob(p, y) == a.marginal_loglik
    sm = model.smoother(p, y)
    assert sm.smoothed_means.shape == (K, 4) and sm.smoothed_cross_covariances.shape == (K - 1, 4, 4)
    assert np.array_equal(model.filter(p, y).filtered_means, a.filtered_means)
    # nonlinear dispatcher: output_fields selects what comes back (inference_ekf.py:308-315)
    pn = _nl_params(cd, g)
    full = cd.cdnlgssm_filter(pn, y)
    part = cd.cdnlgssm_filter(pn, y, output_fields=["filtered_means"])
    assert part.filtered_covariances is None and np.array_equal(part.filtered_means, full.filtered_means)
    cum = cd.cdnlgssm_filter(pn, y, output_fields=["marginal_loglik"])
    assert cum.marginal_loglik.shape == (K,) and abs(cum.marginal_loglik[-1] - full.marginal_loglik) < 1e-12 * abs(full.marginal_loglik)
    nm = cd.ContDiscreteNonlinearGaussianSSM(state_dim=4, emission_dim=2)
    assert nm.marginal_log_prob(pn, y) == full.marginal_loglik
    # batched emissions with a SHARED time grid (t_emissions=None, one [K,1] column, or a [1,K,1] array) broadcast it
    Yb = rng.standard_normal((5, K, 2))
    tg = np.cumsum(rng.uniform(0.5, 1.5, K))
    fb0 = cd.cdlgssm_filter(p, Yb)
    fb1 = cd.cdlgssm_filter(p, Yb, np.arange(K, dtype=np.float64)[:, None])
    assert np.array_equal(fb0.filtered_means, fb1.filtered_means)
    fb2 = cd.cdlgssm_filter(p, Yb, tg[:, None])
    fb3 = cd.cdlgssm_filter(p, Yb, np.repeat(tg[None, :, None], 5, 0))
    fb4 = cd.cdnlgssm_filter(pn, Yb, tg[None, :, None])
    assert np.array_equal(fb2.predicted_covariances, fb3.predicted_covariances)
    assert np.array_equal(fb4.filtered_means[3], cd.cdnlgssm_filter(pn, Yb[3], tg[:, None]).filtered_means)
    # the nonlinear registry models take no inputs: zeros are accepted, anything else is refused (never dropped silently)
    assert np.array_equal(cd.cdnlgssm_filter(pn, y, inputs=np.zeros((K, 1))).filtered_means, full.filtered_means)
    with pytest.raises(NotImplementedError):
        cd.cdnlgssm_filter(pn, y, inputs=np.ones((K, 1)))
    # an empty batch is legal and returns empty arrays
    e = cd.cdlgssm_filter(p, np.zeros((0, K, 2)), np.zeros((0, K, 1)))
    assert e.filtered_means.shape == (0, K, 4) and e.marginal_loglik.shape == (0,)


def test_diagonal_emission_covariance_takes_the_woodbury_branch():
    """A 1-D emissions.cov (cd_linear/inference.py:240-254).  The Woodbury update is algebraically the diag(R) update, so
    filtered and smoothed moments must agree with the full-matrix call (up to where the 1e-9 boost enters); the
    log-likelihood does NOT (the reference broadcasts the vector over H P H^T, :613) except for a scalar emission.  The
    literal numbers are pinned by the kf_diagR_* goldens (reference code on the NumPy shims)."""
    cd = api()
    g = _linear_model(seed=9)
    rng = np.random.default_rng(2)
    K = 20
    y = rng.standard_normal((3, K, 2))
    t = np.cumsum(rng.uniform(0.02, 0.08, (3, K)), axis=1)
    Rd = np.array([0.3, 0.45])
    p_full, p_diag = _lin_params(cd, dict(g, R=np.diag(Rd))), _lin_params(cd, dict(g, R=Rd))
    hp = cd.KFHyperParams(dt_final=0.05, diffeqsolve_settings={"solver": "rk4", "dt0": 0.01})
    a, b = cd.cdlgssm_smoother(p_full, y, t[..., None], hp), cd.cdlgssm_smoother(p_diag, y, t[..., None], hp)
    # the 1e-9 boost sits on I + U^T X here and on S there: agreement to ~1e-8, not to rounding
    np.testing.assert_allclose(b.filtered_means, a.filtered_means, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b.filtered_covariances, a.filtered_covariances, rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(b.smoothed_covariances, a.smoothed_covariances, rtol=1e-6, atol=1e-7)
    assert not np.allclose(b.marginal_loglik, a.marginal_loglik, rtol=1e-6)  # the reference's broadcast, reproduced
    g1 = dict(g, H=g["H"][:1])
    a1 = cd.cdlgssm_filter(_lin_params(cd, dict(g1, R=np.array([[0.3]]))), y[..., :1], t[..., None], hp)
    b1 = cd.cdlgssm_filter(_lin_params(cd, dict(g1, R=np.array([0.3]))), y[..., :1], t[..., None], hp)
    np.testing.assert_allclose(b1.marginal_loglik, a1.marginal_loglik, rtol=1e-7)
    np.testing.assert_allclose(b1.predicted_covariances, a1.predicted_covariances, rtol=1e-6, atol=1e-7)


def test_loud_failures():
    cd = api()
    g = _linear_model(seed=6)
    y = np.zeros((6, 2))
    p, pn = _lin_params(cd, g), _nl_params(cd, g)
    with pytest.raises(ValueError):
        cd.cdlgssm_smoother(p, y, smoother_type="cd_smoother_3")
    with pytest.raises(ValueError):
        cd.cdnlgssm_filter(pn, y, output_fields=["smoothed_means"])
    with pytest.raises(NotImplementedError):  # adaptive step-size control is not in the fixed-step registry
        cd.cdnlgssm_filter(pn, y, hyperparams=cd.EKFHyperParams(diffeqsolve_settings={"solver": "tsit5"}))
    with pytest.raises(NotImplementedError):  # a Python callable cannot run inside a CUDA kernel
        bad = pn._replace(dynamics=pn.dynamics._replace(drift=lambda x, u, t: -x))
        cd.cdnlgssm_filter(bad, y)
    with pytest.raises(ValueError):
        cd.cdnlgssm_smoother(pn, y, hyperparams=cd.UKFHyperParams())  # "UKS not implemented yet", as upstream
    with pytest.raises(NotImplementedError):
        cd.ContDiscreteLinearGaussianSSM(state_dim=4, emission_dim=2).fit_sgd(p, None, y)
