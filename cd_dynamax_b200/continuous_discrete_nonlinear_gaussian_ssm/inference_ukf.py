"""Drop-in for src/continuous_discrete_nonlinear_gaussian_ssm/inference_ukf.py: UKFHyperParams :25-34,
unscented_kalman_filter :206-308 (the smoother does not exist upstream: raise NotImplementedError :332).

The unscented predict runs in CLOSED FORM by default: every drift of the registry is a polynomial of degree <= 2 and the
emission is linear, and for those the 2n + 1 sigma-point sums of inference_ukf.py:130-152 are exactly
f(m) + tr(Hess f P) / 2 and J P (csrc/cdk_generic.cu, ODE_UKFC) -- same moments up to rounding, no Cholesky factor of P at
every RK stage.  `CDK_UKF_SIGMA_POINTS=1` in the environment (read per call) selects the literal sigma-point kernel."""
import math
import os
from typing import List, NamedTuple, Optional

from ..types import PosteriorGSSMFiltered
from ._common import DEFAULT_FIELDS, run_filter


class UKFHyperParams(NamedTuple):
    dt_final: float = 1e-10
    alpha: float = math.sqrt(3)
    beta: int = 2
    kappa: int = 1
    diffeqsolve_settings: dict = {}


def unscented_kalman_filter(params, emissions, t_emissions=None, hyperparams: UKFHyperParams = UKFHyperParams(),
                            inputs=None, output_fields: Optional[List[str]] = DEFAULT_FIELDS) -> PosteriorGSSMFiltered:
    from .. import _lib as L
    fields = dict(dt_final=float(hyperparams.dt_final), alpha=float(hyperparams.alpha), beta=float(hyperparams.beta),
                  kappa=float(hyperparams.kappa))
    if os.environ.get("CDK_UKF_SIGMA_POINTS", "0") == "1":
        fields["flags"] = L.FLAG_UKF_SIGMA_POINTS
    post, _, _ = run_filter("cdk_ukf_filter", params, emissions, t_emissions, inputs, output_fields, fields,
                            diffeqsolve_settings=hyperparams.diffeqsolve_settings)
    return post


def unscented_kalman_smoother(*args, **kwargs):
    raise NotImplementedError("UKS not implemented yet")  # as upstream (inference_ukf.py:332)
