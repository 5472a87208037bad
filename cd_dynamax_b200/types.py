"""Result / parameter containers with the reference's names and fields
(dynamax/linear_gaussian_ssm/inference.py:19-32, :66-96, :112-143)."""
from typing import Any, NamedTuple, Optional


class ParamsLGSSMInitial(NamedTuple):
    mean: Any
    cov: Any


class ParamsLGSSMEmissions(NamedTuple):
    weights: Any
    bias: Any
    input_weights: Any
    cov: Any


class PosteriorGSSMFiltered(NamedTuple):
    marginal_loglik: Any
    filtered_means: Optional[Any] = None
    filtered_covariances: Optional[Any] = None
    predicted_means: Optional[Any] = None
    predicted_covariances: Optional[Any] = None


class PosteriorGSSMSmoothed(NamedTuple):
    marginal_loglik: Any
    filtered_means: Any
    filtered_covariances: Any
    smoothed_means: Any
    smoothed_covariances: Any
    smoothed_cross_covariances: Optional[Any] = None


class ParameterProperties(NamedTuple):
    """Placeholder for dynamax.parameters.ParameterProperties (dynamax/parameters.py:24-50): the drop-in receives
    already-constrained arrays, so properties are carried through `initialize` untouched."""
    trainable: bool = True
    constrainer: Any = None
