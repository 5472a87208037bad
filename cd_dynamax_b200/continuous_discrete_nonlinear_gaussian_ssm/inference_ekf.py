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


_GRAD_GROUPS = {"drift": 1, "diffusion_coefficient": 2, "diffusion_cov": 2, "emission_cov": 4, "emission_bias": 8,
                "emission_weights": 16, "initial_mean": 32, "initial_cov": 64}


def ekf_marginal_log_prob_and_grad(params, emissions, t_emissions=None, hyperparams: EKFHyperParams = EKFHyperParams(),
                                   inputs=None, wrt=("drift",)):
    """Marginal log-likelihood of the CD-EKF and its gradient with respect to the model parameters.

    -> (marginal_loglik, grads): `grads` maps "sigma", "rho", "beta" (wrt "drift"), "diffusion_coefficient" [3,3],
    "diffusion_cov" [3,3], "emission_cov" [1,1], "emission_bias" [1], "emission_weights" [1,3], "initial_mean" [3],
    "initial_cov" [3,3] to arrays, with a leading N for batched emissions.  `wrt` = any of those group names, or "all".

    This is what `jax.value_and_grad(lambda p: model.marginal_log_prob(p, ...))` yields upstream
    (src/utils/optimize_utils.py:102, src/ssm_temissions.py:550-568) -- here as the derivative of exactly the discrete
    filter `cdnlgssm_filter` runs (cdk_ekf_grad_f64): REVERSE mode for more than the drift group (a forward filter pass
    + one backward launch for all 23 columns, ~4 filter passes), forward mode otherwise (one launch per column), so that
    a `jax.custom_vjp` around the filter can be fed.  Symmetric matrices (diffusion_cov, initial_cov) are differentiated
    along symmetric directions: the result is the symmetric part (G + G^T)/2 of the entry-wise gradient G, which is what
    any PSD parameterisation consumes; diffusion_coefficient's gradient is exact.  First step of SURVEY section 8f rank 1:
    LearnableLorenz63 drift, scalar emission, num_iter = 1, fp64; anything else raises NotImplementedError."""
    import torch
    from ._common import nonlinear_inputs, _val
    from ..continuous_discrete_linear_gaussian_ssm.inference import _shape, prepare_data
    if type(params.dynamics.drift).__name__ != "LearnableLorenz63":
        raise NotImplementedError("gradients are implemented for the LearnableLorenz63 drift only (SURVEY 8f rank 1)")
    wrt = tuple(_GRAD_GROUPS) if wrt == "all" else ((wrt,) if isinstance(wrt, str) else tuple(wrt))
    unknown = [w for w in wrt if w not in _GRAD_GROUPS]
    if unknown:
        raise ValueError(f"unknown gradient groups {unknown}; choose from {sorted(_GRAD_GROUPS)}")
    groups = 0
    for w in wrt:
        groups |= _GRAD_GROUPS[w]
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
    # more than the three drift columns: REVERSE mode (one backward launch for all 23 columns behind a forward filter pass
    # whose moments go to device scratch, N*K*192 bytes) when that scratch is affordable; else forward mode, one launch
    # per column.  CDK_GRAD_MODE=forward|reverse overrides.
    import os
    mode = os.environ.get("CDK_GRAD_MODE", "auto")
    free, _ = torch.cuda.mem_get_info()
    reverse = mode == "reverse" or (mode == "auto" and groups != 1 and N * K * 192 <= free // 2)
    fields["grad_groups"] = groups | (L.GRAD_REVERSE if reverse else 0)
    dev_ins = {}
    try:
        out = E.run("cdk_ekf_grad", dt, N, K, n, m, ins, (L.OUT_LL, L.OUT_GRAD), fields, dev_inputs=dev_ins)
    except L.CdkError as e:
        raise NotImplementedError(str(e)) from e
    G = out[L.OUT_GRAD]  # [N, 23] on the device: theta 3 | LQL 6 | R | d | H 3 | m0 3 | P0 6
    g = lambda t: E.from_dev(_sq(t, batched), kind)

    def sym(cols):  # packed symmetric-direction derivatives -> symmetric gradient matrix Gs (off-diagonals halved)
        M = torch.zeros((N, 3, 3), dtype=G.dtype, device=G.device)
        iu = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
        for c, (i, j) in enumerate(iu):
            v = cols[:, c] if i == j else 0.5 * cols[:, c]
            M[:, i, j] = v
            M[:, j, i] = v
        return M

    grads = {}
    if groups & 1:
        grads.update(sigma=g(G[:, 0]), rho=g(G[:, 1]), beta=g(G[:, 2]))
    if groups & 2:
        Gs = sym(G[:, 3:9])
        Lm, Qc = dev_ins[L.IN_L].to(G.dtype), dev_ins[L.IN_QC].to(G.dtype)
        Lm = Lm.expand(N, 3, 3) if Lm.dim() == 2 else Lm
        Qc = Qc.expand(N, 3, 3) if Qc.dim() == 2 else Qc
        if "diffusion_cov" in wrt:
            grads["diffusion_cov"] = g(Lm.transpose(1, 2) @ Gs @ Lm)  # d(L Qc L^T) = L dQc L^T
        if "diffusion_coefficient" in wrt:
            grads["diffusion_coefficient"] = g(2.0 * Gs @ Lm @ Qc)  # d(L Qc L^T) = dL Qc L^T + L Qc dL^T
    if groups & 4:
        grads["emission_cov"] = g(G[:, 9].reshape(N, 1, 1))
    if groups & 8:
        grads["emission_bias"] = g(G[:, 10].reshape(N, 1))
    if groups & 16:
        grads["emission_weights"] = g(G[:, 11:14].reshape(N, 1, 3))
    if groups & 32:
        grads["initial_mean"] = g(G[:, 14:17])
    if groups & 64:
        grads["initial_cov"] = g(sym(G[:, 17:23]))
    return g(out[L.OUT_LL]), grads
