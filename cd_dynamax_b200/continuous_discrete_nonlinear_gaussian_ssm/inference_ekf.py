"""Drop-in for src/continuous_discrete_nonlinear_gaussian_ssm/inference_ekf.py: EKFHyperParams :34-44,
extended_kalman_filter :202-326, iterated_extended_kalman_filter :328-361, extended_kalman_smoother :450-539,
iterated_extended_kalman_smoother :541-593."""
from typing import List, NamedTuple, Optional

from .. import _engine as E
from .. import _lib as L
from ..continuous_discrete_linear_gaussian_ssm.inference import _sq
from ..types import PosteriorGSSMFiltered, PosteriorGSSMSmoothed
from ._common import DEFAULT_FIELDS, run_filter


class EKFHyperParams(NamedTuple):
    dt_final: float = 1e-10
    state_order: str = "second"
    emission_order: str = "first"
    smooth_order: str = "first"
    cov_rescaling: float = 1.0
    diffeqsolve_settings: dict = {}


def _ekf_fields(hp: EKFHyperParams, num_iter: int):
    if hp.state_order not in L.ORDERS:
        raise ValueError("EKF hyperparams.state_order = {} not implemented yet".format(hp.state_order))  # :118
    return dict(dt_final=float(hp.dt_final), state_order=L.ORDERS[hp.state_order], num_iter=int(num_iter),
                cov_rescaling=float(hp.cov_rescaling))


def extended_kalman_filter(params, emissions, t_emissions=None, hyperparams: EKFHyperParams = EKFHyperParams(),
                           inputs=None, num_iter: int = 1,
                           output_fields: Optional[List[str]] = DEFAULT_FIELDS) -> PosteriorGSSMFiltered:
    post, _, _ = run_filter("cdk_ekf_filter", params, emissions, t_emissions, inputs, output_fields,
                            _ekf_fields(hyperparams, num_iter), diffeqsolve_settings=hyperparams.diffeqsolve_settings)
    return post


def iterated_extended_kalman_filter(params, emissions, t_emissions=None, hyperparams: EKFHyperParams = EKFHyperParams(),
                                    inputs=None, num_iter: int = 2,
                                    output_fields: Optional[List[str]] = DEFAULT_FIELDS) -> PosteriorGSSMFiltered:
    return extended_kalman_filter(params, emissions, t_emissions, hyperparams, inputs, num_iter, output_fields)


def extended_kalman_smoother(params, emissions, hyperparams: EKFHyperParams = EKFHyperParams(), t_emissions=None,
                             filtered_posterior: Optional[PosteriorGSSMFiltered] = None,
                             inputs=None) -> PosteriorGSSMSmoothed:
    """Forward EKF (num_iter = 1, as upstream :503-509) then the backward ODE with the Jacobian frozen at the filtered
    mean (:363-448)."""
    if hyperparams.smooth_order != "first":
        raise ValueError("EKF hyperparams.smooth_order = {} not implemented yet".format(hyperparams.smooth_order))
    post, out, (ins, fields, N, K, n, m, dt, batched, kind) = run_filter(
        "cdk_ekf_filter", params, emissions, t_emissions, inputs, ["filtered_means", "filtered_covariances"],
        _ekf_fields(hyperparams, 1), diffeqsolve_settings=hyperparams.diffeqsolve_settings, keep_on_device=True)
    fm, fp = out[L.OUT_FM], out[L.OUT_FP]
    if filtered_posterior is not None:
        # upstream uses a caller-supplied filtered posterior verbatim (:497-512)
        dev = fm.device
        fm = E.to_dev(filtered_posterior.filtered_means, dt, dev).reshape(N, K, n)
        fp = E.to_dev(filtered_posterior.filtered_covariances, dt, dev).reshape(N, K, n, n)
    ins = dict(ins)
    ins[L.IN_FM], ins[L.IN_FP] = fm, fp
    sm = E.run("cdk_ekf_smooth", dt, N, K, n, m, ins, (L.OUT_SM, L.OUT_SP), fields, status=out[L.OUT_STATUS])
    g = lambda t: E.from_dev(_sq(t, batched), kind)
    ll = filtered_posterior.marginal_loglik if filtered_posterior is not None else g(out[L.OUT_LL])
    return PosteriorGSSMSmoothed(marginal_loglik=ll, filtered_means=g(fm), filtered_covariances=g(fp),
                                 smoothed_means=g(sm[L.OUT_SM]), smoothed_covariances=g(sm[L.OUT_SP]))


def iterated_extended_kalman_smoother(params, emissions, hyperparams: EKFHyperParams = EKFHyperParams(),
                                      t_emissions=None, num_iter: int = 2, inputs=None) -> PosteriorGSSMSmoothed:
    """Upstream silently runs a single smoothing pass (:577-586); so does this."""
    return extended_kalman_smoother(params, emissions, hyperparams, t_emissions, None, inputs)


def ekf_marginal_log_prob_and_grad(params, emissions, t_emissions=None, hyperparams: EKFHyperParams = EKFHyperParams(),
                                   inputs=None):
    """Marginal log-likelihood of the CD-EKF and its gradient with respect to the drift parameters:
    -> (marginal_loglik, {"sigma": .., "rho": .., "beta": ..}), batched like the filter ([N] each for [N,K,m] emissions).

    This is what `jax.value_and_grad(lambda p: model.marginal_log_prob(p, ...))` yields upstream for the drift leaves
    (src/utils/optimize_utils.py:102, src/ssm_temissions.py:550-568) -- here as a forward-mode derivative of exactly the
    discrete filter `cdnlgssm_filter` runs (cdk_ekf_grad_f64), so that a `jax.custom_vjp` around the filter can be fed
    (INTEGRATION.md).  First step of SURVEY section 8f rank 1: LearnableLorenz63 drift, scalar emission, num_iter = 1;
    anything else raises NotImplementedError."""
    from .cdnlgssm_utils import drift_to_theta
    from ._common import nonlinear_inputs, _val
    from ..continuous_discrete_linear_gaussian_ssm.inference import _shape, prepare_data
    if type(params.dynamics.drift).__name__ != "LearnableLorenz63":
        raise NotImplementedError("gradients are implemented for the LearnableLorenz63 drift only (SURVEY 8f rank 1)")
    kind = E.kind_of(emissions)
    Y, T, U, batched = prepare_data(emissions, t_emissions, None)
    N, K, m = _shape(Y)
    if m != 1:
        raise NotImplementedError("gradients are implemented for scalar emissions only")
    n = 3
    dt = E.pick_dtype(emissions)
    if dt != "f64":
        raise NotImplementedError("gradients are fp64 only")
    ins, drift_fields = nonlinear_inputs(params, Y, T, n, m)
    fields = dict(E.parse_settings(hyperparams.diffeqsolve_settings), **drift_fields, **_ekf_fields(hyperparams, 1))
    try:
        out = E.run("cdk_ekf_grad", dt, N, K, n, m, ins, (L.OUT_LL, L.OUT_GRAD), fields, host_out=(kind != "cuda"))
    except L.CdkError as e:
        raise NotImplementedError(str(e)) from e
    g = lambda t: E.from_dev(_sq(t, batched), kind)
    grad = out[L.OUT_GRAD]
    return g(out[L.OUT_LL]), {"sigma": g(grad[:, 0]), "rho": g(grad[:, 1]), "beta": g(grad[:, 2])}
