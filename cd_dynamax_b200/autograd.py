"""Differentiable total log-likelihood for torch callers: the shape `jax.custom_vjp` gives the filter in the reference
stack -- forward = the CUDA filter, backward = the hand-written reverse-mode kernel
(cdk_ekf_grad_f64 with CDK_GRAD_REVERSE), no autodiff through the solver.

    ll = ekf_marginal_log_prob(params, emissions, t_emissions, hyperparams)   # [N] (or scalar), differentiable w.r.t. the
    ll.sum().backward()                                                       # torch leaves of `params`

What `fit_sgd` minimises upstream is `-vmap(marginal_log_prob)(...).sum()` (src/ssm_temissions.py:550-568,
src/utils/optimize_utils.py:102).  Lorenz-63 drift, scalar linear emission, fp64 (the coverage of cdk_ekf_grad_f64)."""
import torch

from .continuous_discrete_nonlinear_gaussian_ssm.inference_ekf import EKFHyperParams, ekf_marginal_log_prob_and_grad

_LEAVES = ("sigma", "rho", "beta", "diffusion_coefficient", "diffusion_cov", "emission_cov", "emission_bias",
           "emission_weights", "initial_mean", "initial_cov")


def _val(x):
    return x.params if hasattr(x, "params") and hasattr(x, "f") else x


def _leaves(params):
    d, e = params.dynamics, params.emissions
    return dict(sigma=d.drift.sigma, rho=d.drift.rho, beta=d.drift.beta, diffusion_coefficient=_val(d.diffusion_coefficient),
                diffusion_cov=_val(d.diffusion_cov), emission_cov=_val(e.emission_cov), emission_bias=e.emission_function.bias,
                emission_weights=e.emission_function.weights, initial_mean=_val(params.initial.mean),
                initial_cov=_val(params.initial.cov))


class _EKFLogLik(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, emissions, t_emissions, hyperparams, *leaf_tensors):
        ll, grads = ekf_marginal_log_prob_and_grad(params, emissions, t_emissions, hyperparams, wrt="all")
        ll = torch.as_tensor(ll)
        ctx.grads = [torch.as_tensor(grads[name]) for name in _LEAVES]
        ctx.shapes = [tuple(t.shape) for t in leaf_tensors]
        ctx.batched = ll.dim() == 1
        return ll

    @staticmethod
    def backward(ctx, grad_ll):
        out = []
        for g, shape in zip(ctx.grads, ctx.shapes):
            g = g.to(grad_ll.device)
            if ctx.batched:
                w = grad_ll.reshape((-1,) + (1,) * (g.dim() - 1))
                g = g * w
                if len(shape) < g.dim():  # a parameter shared by the batch: sum over trajectories
                    g = g.sum(0)
            else:
                g = g * grad_ll
            out.append(g.reshape(shape))
        return (None, None, None, None) + tuple(out)


def ekf_marginal_log_prob(params, emissions, t_emissions=None, hyperparams: EKFHyperParams = EKFHyperParams()):
    """Per-trajectory marginal log-likelihood of the CD-EKF as a differentiable torch tensor.  Every leaf of `params`
    that is a torch tensor with requires_grad receives its gradient on `.backward()`."""
    leaves = _leaves(params)
    tensors = [leaves[n] if isinstance(leaves[n], torch.Tensor) else torch.as_tensor(leaves[n]) for n in _LEAVES]
    return _EKFLogLik.apply(params, emissions, t_emissions, hyperparams, *tensors)
