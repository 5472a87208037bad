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
