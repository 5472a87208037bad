"""Drop-in shell for ContDiscreteLinearGaussianSSM (src/continuous_discrete_linear_gaussian_ssm/models.py):
initialize :110-243, marginal_log_prob :336-345, filter :347-355, smoother :357-365.  Sampling, EM and the fit_*
drivers are out of scope."""
from typing import Optional

import numpy as np

from ..types import ParameterProperties, ParamsLGSSMEmissions, ParamsLGSSMInitial
from .inference import (KFHyperParams, ParamsCDLGSSM, ParamsCDLGSSMDynamics, cdlgssm_filter, cdlgssm_smoother)


class ContDiscreteLinearGaussianSSM:
    def __init__(self, state_dim: int, emission_dim: int, input_dim: int = 0, has_dynamics_bias: bool = True,
                 has_emissions_bias: bool = True, diffeqsolve_settings: dict = {}):
        self.state_dim = state_dim
        self.emission_dim = emission_dim
        self.input_dim = input_dim
        self.has_dynamics_bias = has_dynamics_bias
        self.has_emissions_bias = has_emissions_bias
        self._diffeqsolve_settings = diffeqsolve_settings

    @property
    def emission_shape(self):
        return (self.emission_dim,)

    @property
    def inputs_shape(self):
        return (self.input_dim,) if self.input_dim > 0 else None

    @property
    def diffeqsolve_settings(self):
        return self._diffeqsolve_settings

    def initialize(self, key=0, initial_mean: dict = None, initial_cov: dict = None, dynamics_weights: dict = None,
                   dynamics_bias: dict = None, dynamics_input_weights: dict = None,
                   dynamics_diffusion_coefficient: dict = None, dynamics_diffusion_cov: dict = None,
                   dynamics_approx_order: Optional[float] = 2., emission_weights: dict = None,
                   emission_bias: dict = None, emission_input_weights: dict = None, emission_cov: dict = None):
        """Same defaults as upstream (:146-201); `key` seeds NumPy instead of jax.random for the default H."""
        n, m, du = self.state_dim, self.emission_dim, self.input_dim
        fixed = ParameterProperties(trainable=False)
        seed = int(np.asarray(key).ravel()[-1]) if not isinstance(key, int) else key
        dflt = lambda x, p: x if x is not None else {"params": p, "props": fixed}
        initial_mean = dflt(initial_mean, np.zeros(n))
        initial_cov = dflt(initial_cov, np.eye(n))
        dynamics_weights = dflt(dynamics_weights, -0.1 * np.eye(n))
        dynamics_input_weights = dflt(dynamics_input_weights, np.zeros((n, du)))
        dynamics_bias = dflt(dynamics_bias, np.zeros(n) if self.has_dynamics_bias else None)
        dynamics_diffusion_coefficient = dflt(dynamics_diffusion_coefficient, 0.1 * np.eye(n))
        dynamics_diffusion_cov = dflt(dynamics_diffusion_cov, 0.1 * np.eye(n))
        emission_weights = dflt(emission_weights, np.random.default_rng(seed).standard_normal((m, n)))
        emission_input_weights = dflt(emission_input_weights, np.zeros((m, du)))
        emission_bias = dflt(emission_bias, np.zeros(m) if self.has_emissions_bias else None)
        emission_cov = dflt(emission_cov, 0.1 * np.eye(m))
        out = {}
        for k in ("params", "props"):
            out[k] = ParamsCDLGSSM(
                initial=ParamsLGSSMInitial(mean=initial_mean[k], cov=initial_cov[k]),
                dynamics=ParamsCDLGSSMDynamics(weights=dynamics_weights[k], input_weights=dynamics_input_weights[k],
                                               bias=dynamics_bias[k],
                                               diffusion_coefficient=dynamics_diffusion_coefficient[k],
                                               diffusion_cov=dynamics_diffusion_cov[k]),
                emissions=ParamsLGSSMEmissions(weights=emission_weights[k], input_weights=emission_input_weights[k],
                                               bias=emission_bias[k], cov=emission_cov[k]))
        return out["params"], out["props"]

    def marginal_log_prob(self, params, emissions, t_emissions=None, filter_hyperparams: Optional[KFHyperParams] = None,
                          inputs=None):
        return cdlgssm_filter(params, emissions, t_emissions, filter_hyperparams, inputs).marginal_loglik

    def filter(self, params, emissions, t_emissions=None, filter_hyperparams: Optional[KFHyperParams] = None,
               inputs=None):
        return cdlgssm_filter(params, emissions, t_emissions, filter_hyperparams, inputs)

    def smoother(self, params, emissions, t_emissions=None, filter_hyperparams: Optional[KFHyperParams] = None,
                 inputs=None):
        return cdlgssm_smoother(params, emissions, t_emissions, filter_hyperparams, inputs)

    def _unsupported(self, *a, **k):
        raise NotImplementedError("outside the hot path this package replaces (sampling / EM / fit_*); "
                                  "use the reference implementation for these")

    sample = posterior_sample = posterior_predictive = e_step = m_step = fit_em = fit_sgd = fit_mcmc = _unsupported
