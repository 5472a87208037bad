"""Drop-in for src/continuous_discrete_linear_gaussian_ssm/inference.py (filter :555-632, smoother :694-823):
same names, argument order and result tuples; the arithmetic runs in libcdk.so on the GPU.

Batched form (what `jax.vmap` produces in the reference, src/ssm_temissions.py:555-567): emissions [N,K,m],
t_emissions [N,K,1]; any parameter may carry a leading N as well.
"""
from typing import Any, NamedTuple, Optional

import numpy as np
import torch

from .. import _engine as E
from .. import _lib as L
from ..types import ParamsLGSSMEmissions, ParamsLGSSMInitial, PosteriorGSSMFiltered, PosteriorGSSMSmoothed


class KFHyperParams(NamedTuple):
    """cd_linear/inference.py:34-38"""
    dt_final: float = 1e-10
    diffeqsolve_settings: dict = {}


class ParamsCDLGSSMDynamics(NamedTuple):
    """cd_linear/inference.py:57-89"""
    weights: Any
    bias: Any
    input_weights: Any
    diffusion_coefficient: Any
    diffusion_cov: Any


class ParamsCDLGSSM(NamedTuple):
    """cd_linear/inference.py:92-102"""
    initial: ParamsLGSSMInitial
    dynamics: ParamsCDLGSSMDynamics
    emissions: ParamsLGSSMEmissions


def make_cdlgssm_params(initial_mean, initial_cov, dynamics_weights, dynamics_diffusion_coeff, dynamics_diffusion_cov,
                        emissions_weights, emissions_cov, dynamics_bias=None, dynamics_input_weights=None,
                        emissions_bias=None, emissions_input_weights=None):
    """cd_linear/inference.py:146-182"""
    return ParamsCDLGSSM(
        initial=ParamsLGSSMInitial(mean=initial_mean, cov=initial_cov),
        dynamics=ParamsCDLGSSMDynamics(weights=dynamics_weights, bias=dynamics_bias,
                                       input_weights=dynamics_input_weights,
                                       diffusion_coefficient=dynamics_diffusion_coeff,
                                       diffusion_cov=dynamics_diffusion_cov),
        emissions=ParamsLGSSMEmissions(weights=emissions_weights, bias=emissions_bias,
                                       input_weights=emissions_input_weights, cov=emissions_cov))


def _shape(x):
    return tuple(x.shape) if hasattr(x, "shape") else np.shape(x)


def _check_constant(x, name, core):
    if callable(x):
        raise NotImplementedError(f"{name}: callable (time-varying) parameters are not supported "
                                  "(cd_linear/inference.py:42-51 _get_params); pass constant arrays")
    return x


def prepare_data(emissions, t_emissions, inputs):
    """-> (Y [N,K,m], T [N,K], U [N,K,d_u] or None, batched?)"""
    ys = _shape(emissions)
    batched = len(ys) == 3
    if len(ys) not in (2, 3):
        raise ValueError(f"emissions must be [K,m] or [N,K,m], got {ys}")
    K = ys[-2]
    Y = emissions if batched else emissions[None]
    if t_emissions is None:
        # unit spacing (cd_linear/inference.py:590-593)
        T = np.arange(K, dtype=np.float64)[None]
    else:
        ts = _shape(t_emissions)
        if ts[-1] == 1 and len(ts) >= 2 and ts[-2] == K:
            T = t_emissions[..., 0]  # the reference's [K,1] column
        elif ts[-1] == K:
            T = t_emissions
        else:
            raise ValueError(f"t_emissions must be [K,1] (or [N,K,1]); got {ts} for K={K}")
        if len(_shape(T)) == 1:
            T = T[None]
    U = None
    if inputs is not None and _shape(inputs)[-1] > 0:
        U = inputs if len(_shape(inputs)) == 3 else inputs[None]
    return Y, T, U, batched


def _diag_r(params) -> bool:
    """A 1-D `emissions.cov` [m] (or [N, m] when batched over parameter samples together with 3-D weights) is the diagonal
    of R: the reference's Woodbury update (cd_linear/inference.py:240-254)."""
    rs, hs = _shape(params.emissions.cov), _shape(params.emissions.weights)
    return len(rs) == 1 or (len(rs) == 2 and len(hs) == 3 and rs[-1] != rs[-2])


def _linear_inputs(params: ParamsCDLGSSM, Y, T, U, n, m, d_u):
    dyn, em = params.dynamics, params.emissions
    for nm, v in (("dynamics.weights", dyn.weights), ("dynamics.diffusion_coefficient", dyn.diffusion_coefficient),
                  ("dynamics.diffusion_cov", dyn.diffusion_cov), ("emissions.weights", em.weights),
                  ("emissions.cov", em.cov)):
        assert v is not None, f"params.{nm} is required (cd_linear/inference.py:266-272)"
        _check_constant(v, nm, 2)
    zeros = lambda *s: np.zeros(s)
    ins = {
        L.IN_Y: Y, L.IN_T: T, L.IN_M0: params.initial.mean, L.IN_P0: params.initial.cov, L.IN_F: dyn.weights,
        L.IN_B: dyn.bias if dyn.bias is not None else zeros(n), L.IN_L: dyn.diffusion_coefficient,
        L.IN_QC: dyn.diffusion_cov, L.IN_H: em.weights, L.IN_D: em.bias if em.bias is not None else zeros(m),
        L.IN_R: em.cov,
    }
    if d_u > 0:
        ins[L.IN_U] = U
        ins[L.IN_BU] = dyn.input_weights if dyn.input_weights is not None else zeros(n, d_u)
        ins[L.IN_DU] = em.input_weights if em.input_weights is not None else zeros(m, d_u)
    return ins


def _filter_device(params, emissions, t_emissions, filter_hyperparams, inputs, want, host_out=False, flags=0):
    hp = filter_hyperparams if filter_hyperparams is not None else KFHyperParams()  # None crashes upstream (:585)
    Y, T, U, batched = prepare_data(emissions, t_emissions, inputs)
    N, K, m = _shape(Y)
    n = _shape(params.emissions.weights)[-1]
    d_u = _shape(U)[-1] if U is not None else 0
    dt = E.pick_dtype(emissions)
    ins = _linear_inputs(params, Y, T, U, n, m, d_u)
    fields = dict(E.parse_settings(hp.diffeqsolve_settings), dt_final=float(hp.dt_final), d_u=d_u)
    diag = _diag_r(params)
    if diag:
        fields["flags"] = L.FLAG_DIAG_R
    elif flags and dt == "f64" and K > 1:
        # keep (A_k, Q_k) of every gap for the type-1 smoother when the cache is affordable (a quarter of the free HBM)
        free, _ = torch.cuda.mem_get_info()
        if N * (K - 1) * 2 * n * n * 8 <= free // 4:
            fields["flags"] = flags
    dev_ins = {}
    out = E.run("cdk_kf_filter", dt, N, K, n, m, ins, want, fields, theta_core_ndim=2, host_out=host_out,
                dev_inputs=dev_ins, core_ndim_override={L.IN_R: 1} if diag else None)
    return out, {**ins, **dev_ins}, fields, (N, K, n, m, dt, batched)


def _sq(t, batched):
    return t if (batched or t is None) else t[0]


def cdlgssm_filter(params: ParamsCDLGSSM, emissions, t_emissions=None,
                   filter_hyperparams: Optional[KFHyperParams] = KFHyperParams(), inputs=None) -> PosteriorGSSMFiltered:
    """Continuous-discrete Kalman filter (cd_linear/inference.py:555-632)."""
    kind = E.kind_of(emissions)
    want = (L.OUT_LL, L.OUT_FM, L.OUT_FP, L.OUT_PM, L.OUT_PP)
    out, _, _, (N, K, n, m, dt, batched) = _filter_device(params, emissions, t_emissions, filter_hyperparams, inputs,
                                                          want, host_out=(kind != "cuda"))
    g = lambda s: E.from_dev(_sq(out[s], batched), kind)
    return PosteriorGSSMFiltered(marginal_loglik=g(L.OUT_LL), filtered_means=g(L.OUT_FM),
                                 filtered_covariances=g(L.OUT_FP), predicted_means=g(L.OUT_PM),
                                 predicted_covariances=g(L.OUT_PP))


def cdlgssm_smoother(params: ParamsCDLGSSM, emissions, t_emissions=None,
                     filter_hyperparams: Optional[KFHyperParams] = None, inputs=None,
                     smoother_type: Optional[str] = "cd_smoother_1") -> PosteriorGSSMSmoothed:
    """Forward filter + backward smoother (cd_linear/inference.py:694-823): 'cd_smoother_1' = Sarkka Alg. 3.17,
    'cd_smoother_2' = Alg. 3.18 (backward ODE; cross-covariances are NaN as upstream, :792)."""
    if smoother_type not in ("cd_smoother_1", "cd_smoother_2"):
        raise ValueError("CD Kalman Smoother type = {} not implemented yet".format(smoother_type))  # :798-799
    kind = E.kind_of(emissions)
    want = (L.OUT_LL, L.OUT_FM, L.OUT_FP)
    out, ins, fields, (N, K, n, m, dt, batched) = _filter_device(
        params, emissions, t_emissions, filter_hyperparams, inputs, want,
        flags=L.FLAG_KEEP_PUSHFORWARD if smoother_type == "cd_smoother_1" else 0)
    ins = dict(ins)
    ins[L.IN_FM], ins[L.IN_FP] = out[L.OUT_FM], out[L.OUT_FP]
    fields = dict(fields, smoother_type=1 if smoother_type == "cd_smoother_1" else 2)
    if L.OUT_SCRATCH not in out and "flags" in fields:
        # the filter kept nothing (not the warp kernel): the smoother re-integrates the pushforward
        fields["flags"] = int(fields["flags"]) & ~L.FLAG_KEEP_PUSHFORWARD
    sm = E.run("cdk_kf_smooth", dt, N, K, n, m, ins, (L.OUT_SM, L.OUT_SP, L.OUT_SCROSS), fields, theta_core_ndim=2,
               status=out[L.OUT_STATUS], scratch=out.get(L.OUT_SCRATCH),
               core_ndim_override={L.IN_R: 1} if _diag_r(params) else None)
    g = lambda t: E.from_dev(_sq(t, batched), kind)
    return PosteriorGSSMSmoothed(marginal_loglik=g(out[L.OUT_LL]), filtered_means=g(out[L.OUT_FM]),
                                 filtered_covariances=g(out[L.OUT_FP]), smoothed_means=g(sm[L.OUT_SM]),
                                 smoothed_covariances=g(sm[L.OUT_SP]),
                                 smoothed_cross_covariances=g(sm[L.OUT_SCROSS]))
