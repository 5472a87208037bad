"""CPU: the NumPy oracle against golden vectors produced by the reference's own code (tests/golden/make_golden.py)
and against the constants the reference hard-codes."""
import numpy as np
import pytest

from oracle import cd_oracle as o
from tests.helpers import (golden_cases, linear_params, load_golden, max_rel_err, nonlinear_params, scaled_err,
                           settings_of)

TOL = 1e-9  # north-star fp64 parity bound
FIELDS = ["filtered_means", "filtered_covariances", "predicted_means", "predicted_covariances"]


def test_reference_hardcoded_pushforward_constants():
    """cdlgssm_test_filter_TRegular.py:61-62: Dopri5, dt0=0.01, Delta=1, F=-0.1, LQcL^T=0.125, float32."""
    for dt, tol in ((np.float32, 4 * np.finfo(np.float32).eps), (np.float64, 1e-14)):
        F, LQL = np.array([[[-0.1]]], dt), np.array([[[0.125]]], dt)
        A, Q, _ = o.compute_pushforward(F, LQL, np.zeros(1, dt), np.ones(1, dt), o.SolverSettings())
        ref_A = np.float32(0.9048373699188232421875) if dt is np.float32 else np.exp(-0.1)
        ref_Q = np.float32(0.11329327523708343505859375) if dt is np.float32 else 0.125 * (1 - np.exp(-0.2)) / 0.2
        assert abs(float(A[0, 0, 0]) - float(ref_A)) <= tol * float(ref_A)
        assert abs(float(Q[0, 0, 0]) - float(ref_Q)) <= tol * float(ref_Q)
    assert o.substep_counts(np.zeros(1), np.ones(1), 0.01)[0] == 100
    g = load_golden("pushforward_constants")
    assert abs(g["A_float64"].ravel()[0] - np.exp(-0.1)) < 1e-14


@pytest.mark.parametrize("name", golden_cases("kf_"))
def test_linear_filter_and_smoothers(name):
    g = load_golden(name)
    p = linear_params(g)
    u = g.get("u")
    for stype in (1, 2):
        if stype == 2 and "s2_smoothed_means" not in g:
            continue
        r = o.cdlgssm_smoother(p, g["y"], g["t"], float(g["dt_final"]), settings_of(g), u, smoother_type=stype)
        if stype == 1:
            assert max_rel_err(r["marginal_loglik"], g["filt_marginal_loglik"]) < TOL
            for f in FIELDS:
                assert scaled_err(r[f], g["filt_" + f]) < TOL, f
            assert scaled_err(r["smoothed_cross_covariances"], g["s1_smoothed_cross_covariances"]) < TOL
        else:
            assert np.isnan(r["smoothed_cross_covariances"]).all() or r["smoothed_cross_covariances"].size == 0
        for f in ("smoothed_means", "smoothed_covariances"):
            # the RTS recursion amplifies rounding by cond(P_pred); 1e-8 on the scaled error is the honest bound
            assert scaled_err(r[f], g[f"s{stype}_" + f]) < 1e-8, (stype, f)


@pytest.mark.parametrize("name", golden_cases("ekf_"))
def test_ekf_and_eks(name):
    g = load_golden(name)
    p = nonlinear_params(g)
    r = o.extended_kalman_filter(p, g["y"], g["t"], float(g["dt_final"]), str(g["state_order"]), settings_of(g),
                                 int(g["num_iter"]), float(g["cov_rescaling"]))
    assert max_rel_err(r["marginal_loglik"], g["filt_marginal_loglik"]) < TOL
    assert max_rel_err(r["marginal_loglik_cumulative"], g["filt_marginal_loglik_cumulative"]) < TOL
    for f in FIELDS:
        assert scaled_err(r[f], g["filt_" + f]) < TOL, f
    if "smooth_smoothed_means" in g:
        s = o.extended_kalman_smoother(p, g["y"], g["t"], float(g["dt_final"]), str(g["state_order"]), settings_of(g))
        for f in ("smoothed_means", "smoothed_covariances"):
            assert scaled_err(s[f], g["smooth_" + f]) < 1e-8, f


@pytest.mark.parametrize("name", golden_cases("ukf_"))
def test_ukf(name):
    g = load_golden(name)
    p = nonlinear_params(g)
    r = o.unscented_kalman_filter(p, g["y"], g["t"], float(g["dt_final"]), settings=settings_of(g))
    assert max_rel_err(r["marginal_loglik"], g["filt_marginal_loglik"]) < TOL
    for f in FIELDS:
        assert scaled_err(r[f], g["filt_" + f]) < TOL, f


def test_enkf_oracle_matches_kalman_filter_in_distribution():
    """Pins the EnKF restatement the way the reference's own test does (cdnlgssm_test_filter_linear_TRegular.py:434-470):
    on a linear model its moments approach the CD-KF's as the ensemble grows."""
    rng = np.random.default_rng(11)
    n, m, N, K = 3, 2, 2, 15
    F = -0.5 * np.eye(n) + 0.2 * rng.standard_normal((n, n))
    H = rng.standard_normal((m, n))
    gaps = 0.05 * rng.uniform(0.5, 1.5, size=(N, K)); gaps[:, 0] = 0.0
    t = np.cumsum(gaps, axis=1)
    y = rng.standard_normal((N, K, m))
    lin = o.LinearParams(m0=np.zeros(n), P0=np.eye(n), F=F, L=np.eye(n), Qc=0.3 * np.eye(n), H=H, R=0.5 * np.eye(m))
    kf = o.cdlgssm_filter(lin, y, t, settings=o.SolverSettings("dopri5", 0.01))
    nl = o.NonlinearParams(m0=np.zeros(n), P0=np.eye(n), drift=o.LinearDrift(F, np.zeros(n)), L=np.eye(n), Qc=0.3 * np.eye(n),
                           H=H, R=0.5 * np.eye(m))
    errs = []
    for E in (256, 4096):
        en = o.ensemble_kalman_filter(nl, y, t, E=E, seed=5, settings=o.SolverSettings("euler", 0.0025))
        errs.append(np.max(np.abs(en["filtered_means"] - kf["filtered_means"])))
    assert errs[1] < 0.08 and errs[1] < errs[0]


def test_c_oracle_matches_numpy_oracle():
    """The C restatement (oracle/cd_oracle_c.c: the CPU baseline of bench.py and the checker of the K = 1,000 GPU parity
    tests) against the NumPy oracle that the golden vectors pin: same algorithm, operations in a different order.
    Also the evidence for the absolute floor of the element-wise gate (tests/helpers.gate_err): WITHOUT it (floor 1e-12 of
    the scale on the denominator) two CPU restatements of one algorithm already disagree at ~1e-9 on zero-crossing entries
    while agreeing to ~1e-14 of the scale."""
    from oracle import cpu_baseline as cb
    from tests.helpers import gate_err, moment_norm_err, scaled_err
    rng = np.random.default_rng(1237)
    N, K = 48, 400
    gaps = 0.01 * rng.uniform(0.5, 1.5, size=(N, K))
    gaps[:, 0] = 0.0
    t = np.cumsum(gaps, axis=1)
    y = 8.0 * rng.standard_normal((N, K, 1))
    po = o.NonlinearParams(m0=np.zeros(3), P0=5 * np.eye(3), drift=o.Lorenz63Drift(10.0, 28.0, 8.0 / 3.0), L=np.eye(3),
                           Qc=np.eye(3), H=np.array([[1.0, 0.0, 0.0]]), R=np.eye(1), d=np.zeros(1))
    r = o.extended_kalman_filter(po, y, t, settings=o.SolverSettings("rk4", 0.0025))
    c = cb.filter_c("ekf", y, t, po.m0, po.P0, np.array([10.0, 28.0, 8.0 / 3.0]), po.L, po.Qc, po.H, po.d, po.R,
                    drift_id=1, solver="rk4", dt0=0.0025)
    assert max_rel_err(c["marginal_loglik"], r["marginal_loglik"]) < 1e-12
    for fld, core in (("filtered_means", 1), ("filtered_covariances", 2), ("predicted_means", 1), ("predicted_covariances", 2)):
        assert gate_err(c[fld], r[fld]) < 1e-10 and moment_norm_err(c[fld], r[fld], core) < 1e-11, fld
        assert scaled_err(c[fld], r[fld]) < 1e-12, fld
    # linear CD-KF, n = 16, m = 4 (BASELINE config 2 model), RK4 and the reference-default Dopri5
    n, m, N, K = 16, 4, 6, 60
    F = -0.5 * np.eye(n) + 0.3 * rng.standard_normal((n, n)) / np.sqrt(n)
    lp = o.LinearParams(m0=np.zeros(n), P0=np.eye(n), F=F, L=np.eye(n), Qc=0.1 * np.eye(n), H=np.eye(n)[:m],
                        R=0.1 * np.eye(m), b=0.05 * rng.standard_normal(n), d=np.zeros(m))
    gaps = 0.04 * rng.uniform(0.5, 1.5, size=(N, K))
    gaps[:, 0] = 0.0
    t = np.cumsum(gaps, axis=1)
    y = rng.standard_normal((N, K, m))
    for solver in ("rk4", "dopri5"):
        r = o.cdlgssm_filter(lp, y, t, settings=o.SolverSettings(solver, 0.01))
        c = cb.filter_c("kf", y, t, lp.m0, lp.P0, F, lp.L, lp.Qc, lp.H, lp.d, lp.R, bias=lp.b, solver=solver, dt0=0.01)
        assert max_rel_err(c["marginal_loglik"], r["marginal_loglik"]) < 1e-12
        for fld, core in (("filtered_means", 1), ("filtered_covariances", 2), ("predicted_means", 1), ("predicted_covariances", 2)):
            assert gate_err(c[fld], r[fld]) < 1e-10 and moment_norm_err(c[fld], r[fld], core) < 1e-11, (solver, fld)


def test_normal_deviate_stream_is_pinned():
    """Known-answer pins of the counter-based deviate stream (Philox4x32-10 + the float32 Box-Muller made of correctly
    rounded operations only, oracle/cd_oracle.py:box_muller_f32): exact bit patterns, so a change of the stream -- which the
    GPU must mirror bit for bit (tests/test_gpu_parity.py::test_normal_deviates_bit_identical_to_oracle) -- cannot pass
    silently; plus the edge words (radius word 0 and 2^32 - 1, all quadrants) and the first two moments."""
    i = np.arange(6, dtype=np.uint32)
    z = np.stack(o.philox_normal_quad(i, np.uint32(77), np.uint32(5), np.uint32((2 << 28) | (3 << 8)) + i, 0x1234567887654321), axis=1)
    want = np.array([[1053230178, 1029320931, 3206743056, 1058688914], [1067050446, 1060128608, 1068043539, 1066804496],
                     [3211958930, 1055106219, 1076386205, 3211180195], [1065972825, 3208228336, 1059546727, 1065486141],
                     [3214596317, 3208984820, 1039668577, 1048666161], [1063250250, 3192822815, 3209181155, 1028190270]], np.uint32)
    assert z.dtype == np.float64 and np.array_equal(z.astype(np.float32).astype(np.float64), z)  # exactly fp32-representable
    assert np.array_equal(z.astype(np.float32).view(np.uint32), want)
    ra = np.array([0, 1, 2**32 - 1, 2**31, 123456789], np.uint32)
    rb = np.array([0, 2**32 - 1, 2**30, 2**31, 987654321], np.uint32)
    a, b = o.box_muller_f32(ra, rb)
    assert np.array_equal(a.view(np.uint32), np.array([1087926343, 1087581517, 0, 3214325087, 1051416574], np.uint32))
    assert np.array_equal(b.view(np.uint32), np.array([895848794, 3042966698, 2147483648, 3021986233, 1076439692], np.uint32))
    rng = np.random.default_rng(0)
    w = rng.integers(0, 2**32, size=(2, 400_000), dtype=np.uint64).astype(np.uint32)
    zz = np.concatenate(o.box_muller_f32(w[0], w[1])).astype(np.float64)
    assert abs(zz.mean()) < 4e-3 and abs(zz.var() - 1.0) < 4e-3 and abs((zz**4).mean() - 3.0) < 3e-2 and np.abs(zz).max() < 6.77
