"""Drop-in for the dispatchers and the model shell in src/continuous_discrete_nonlinear_gaussian_ssm/models.py:
cdnlgssm_filter :658-718, cdnlgssm_smoother :720-764, ContDiscreteNonlinearGaussianSSM.initialize :172-290 and
.marginal_log_prob :393-408.  Forecasts, emission moments and path sampling live in forecast.py (models.py:525-656, :767-1047)."""
from typing import List, Optional

import numpy as np

from ..types import ParameterProperties, ParamsLGSSMInitial, PosteriorGSSMFiltered
from ._common import DEFAULT_FIELDS
from .cdnlgssm_utils import (LearnableLinear, LearnableMatrix, LearnableVector, ParamsCDNLGSSM, ParamsCDNLGSSMDynamics,
                             ParamsCDNLGSSMEmissions)
from .inference_ekf import EKFHyperParams, iterated_extended_kalman_filter, iterated_extended_kalman_smoother
from .inference_enkf import EnKFHyperParams, ensemble_kalman_filter
from .inference_ukf import UKFHyperParams, unscented_kalman_filter
from .forecast import cdnlgssm_path_sample


def cdnlgssm_filter(params, emissions, t_emissions=None, hyperparams=EKFHyperParams(), inputs=None,
                    num_iter: Optional[int] = 1,
                    output_fields: Optional[List[str]] = DEFAULT_FIELDS) -> PosteriorGSSMFiltered:
    """The hyper-parameter TYPE selects the algorithm (models.py:689-716)."""
    if isinstance(hyperparams, EKFHyperParams):
        return iterated_extended_kalman_filter(params=params, emissions=emissions, t_emissions=t_emissions,
                                               hyperparams=hyperparams, inputs=inputs, num_iter=num_iter,
                                               output_fields=output_fields)
    if isinstance(hyperparams, EnKFHyperParams):
        return ensemble_kalman_filter(params=params, emissions=emissions, t_emissions=t_emissions,
                                      hyperparams=hyperparams, inputs=inputs, output_fields=output_fields)
    if isinstance(hyperparams, UKFHyperParams):
        return unscented_kalman_filter(params=params, emissions=emissions, t_emissions=t_emissions,
                                       hyperparams=hyperparams, inputs=inputs, output_fields=output_fields)
    raise TypeError(f"hyperparams must be EKFHyperParams, UKFHyperParams or EnKFHyperParams, got {type(hyperparams)}")


def cdnlgssm_smoother(params, emissions, t_emissions=None, hyperparams=EKFHyperParams(), inputs=None,
                      num_iter: Optional[int] = 1):
    if isinstance(hyperparams, EKFHyperParams):
        return iterated_extended_kalman_smoother(params=params, emissions=emissions, t_emissions=t_emissions,
                                                 hyperparams=hyperparams, inputs=inputs, num_iter=num_iter)
    if isinstance(hyperparams, EnKFHyperParams):
        raise ValueError("EnKS not implemented yet")  # models.py:759-760
    if isinstance(hyperparams, UKFHyperParams):
        raise ValueError("UKS not implemented yet")  # models.py:761-762
    raise TypeError(f"unknown hyperparams type {type(hyperparams)}")


class ContDiscreteNonlinearGaussianSSM:
    def __init__(self, state_dim: int, emission_dim: int, input_dim: int = 0, diffeqsolve_settings: dict = {}):
        self.state_dim = state_dim
        self.emission_dim = emission_dim
        self.input_dim = 0  # as upstream (models.py:162)
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

    def initialize(self, key=0, initial_mean: dict = None, initial_cov: dict = None, dynamics_drift: dict = None,
                   dynamics_diffusion_coefficient: dict = None, dynamics_diffusion_cov: dict = None,
                   dynamics_approx_order: Optional[float] = 2., emission_function: dict = None,
                   emission_cov: dict = None):
        """Same defaults as upstream (models.py:188-258)."""
        n, m = self.state_dim, self.emission_dim
        fixed = ParameterProperties(trainable=False)
        seed = int(np.asarray(key).ravel()[-1]) if not isinstance(key, int) else key
        dflt = lambda x, x0: x if x is not None else x0
        initial_mean = dflt(initial_mean, {"params": np.zeros(n), "props": fixed})
        initial_cov = dflt(initial_cov, {"params": np.eye(n), "props": fixed})
        dynamics_drift = dflt(dynamics_drift, {
            "params": LearnableLinear(weights=-0.1 * np.eye(n), bias=np.zeros(n)),
            "props": LearnableLinear(weights=fixed, bias=fixed)})
        dynamics_diffusion_coefficient = dflt(dynamics_diffusion_coefficient, {
            "params": LearnableMatrix(params=0.1 * np.eye(n)), "props": LearnableMatrix(params=fixed)})
        dynamics_diffusion_cov = dflt(dynamics_diffusion_cov, {
            "params": LearnableMatrix(params=0.1 * np.eye(n)), "props": LearnableMatrix(params=fixed)})
        approx = {"params": dflt(dynamics_approx_order, 2.), "props": fixed}
        emission_function = dflt(emission_function, {
            "params": LearnableLinear(weights=np.random.default_rng(seed).standard_normal((m, n)), bias=np.zeros(m)),
            "props": LearnableLinear(weights=fixed, bias=fixed)})
        emission_cov = dflt(emission_cov, {
            "params": LearnableMatrix(params=0.1 * np.eye(m)), "props": LearnableMatrix(params=fixed)})
        out = {}
        for k in ("params", "props"):
            out[k] = ParamsCDNLGSSM(
                initial=ParamsLGSSMInitial(mean=initial_mean[k], cov=initial_cov[k]),
                dynamics=ParamsCDNLGSSMDynamics(drift=dynamics_drift[k],
                                                diffusion_coefficient=dynamics_diffusion_coefficient[k],
                                                diffusion_cov=dynamics_diffusion_cov[k], approx_order=approx[k]),
                emissions=ParamsCDNLGSSMEmissions(emission_function=emission_function[k],
                                                  emission_cov=emission_cov[k]))
        return out["params"], out["props"]

    def marginal_log_prob(self, params, emissions, t_emissions=None, filter_hyperparams=EKFHyperParams(), inputs=None):
        return cdnlgssm_filter(params=params, emissions=emissions, t_emissions=t_emissions,
                               hyperparams=filter_hyperparams, inputs=inputs).marginal_loglik

    def filter(self, *a, **k):
        raise NotImplementedError  # upstream base class raises too (ssm_temissions.py:344-386)

    smoother = filter

    def sample(self, params, key, num_timesteps, t_emissions=None, inputs=None, transition_type="distribution"):
        """SSM.sample (src/ssm_temissions.py:227-330) for transition_type="path": one SDE sample path and its emissions
        (cdnlgssm_path_sample, models.py:525-656).  The Gaussian-transition sampler ("distribution") is not provided."""
        if transition_type != "path":
            raise NotImplementedError('only transition_type="path" (the SDE sampler) is provided; the Gaussian-transition '
                                      'sampler ("distribution") stays with the reference')
        return cdnlgssm_path_sample(params, key, num_timesteps, t_emissions, inputs, self._diffeqsolve_settings)

    def sample_batch(self, params, key, num_sequences, num_timesteps, t_emissions=None, inputs=None,
                     transition_type="distribution"):
        """SSM.sample_batch (src/ssm_temissions.py:187-225): all `num_sequences` paths in one kernel launch."""
        if transition_type != "path":
            raise NotImplementedError('only transition_type="path" (the SDE sampler) is provided')
        return cdnlgssm_path_sample(params, key, num_timesteps, t_emissions, inputs, self._diffeqsolve_settings,
                                    num_sequences=num_sequences)

    def _unsupported(self, *a, **k):
        raise NotImplementedError("outside the hot path this package replaces (fit_*); use the reference implementation")

    fit_sgd = fit_mcmc = fit_em = _unsupported
