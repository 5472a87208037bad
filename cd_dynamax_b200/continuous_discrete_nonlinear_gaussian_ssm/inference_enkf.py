"""Drop-in for src/continuous_discrete_nonlinear_gaussian_ssm/inference_enkf.py: EnKFHyperParams :28-37,
ensemble_kalman_filter :151-276.

Randomness: the reference draws from jax.random (threefry) and a diffrax VirtualBrownianTree, which cannot be
reproduced outside JAX; this implementation uses a counter-based Philox4x32-10 stream keyed by `key` (an int seed, or
a 2-word PRNGKey-like array).  Parity with the reference is therefore distributional, exactly as in its own test
(src/test_scripts/cdnlgssm_test_filter_linear_TRegular.py:434-470).  The SDE step defaults to the reference's Heun scheme
(diffrax_utils.py:121-127: dfx.Heun() whenever a diffusion is present and no solver is given);
`diffeqsolve_settings={"solver": "euler"}` selects Euler-Maruyama (BASELINE config 5)."""
from typing import Any, List, NamedTuple, Optional

import numpy as np

from ..types import PosteriorGSSMFiltered
from ._common import DEFAULT_FIELDS, run_filter


class EnKFHyperParams(NamedTuple):
    dt_final: float = 1e-10
    N_particles: float = 2000
    perturb_measurements: bool = True
    key: Any = 0
    diffeqsolve_settings: dict = {}


def key_to_seed(key) -> int:
    if isinstance(key, (int, np.integer)):
        return int(key) & 0xFFFFFFFFFFFFFFFF
    k = np.asarray(key).astype(np.uint64).ravel()
    return int((int(k[0]) << 32 | int(k[-1])) & 0xFFFFFFFFFFFFFFFF) if k.size > 1 else int(k[0])


def ensemble_kalman_filter(params, emissions, t_emissions=None, hyperparams: EnKFHyperParams = EnKFHyperParams(),
                           inputs=None, output_fields: Optional[List[str]] = DEFAULT_FIELDS,
                           rng_offset: int = 0) -> PosteriorGSSMFiltered:
    settings = dict(hyperparams.diffeqsolve_settings or {})  # no solver given -> parse_settings(sde=True) -> heun
    fields = dict(dt_final=float(hyperparams.dt_final), E=int(hyperparams.N_particles),
                  perturb_measurements=int(bool(hyperparams.perturb_measurements)),
                  rng_seed=key_to_seed(hyperparams.key), rng_offset=int(rng_offset))
    post, _, _ = run_filter("cdk_enkf_filter", params, emissions, t_emissions, inputs, output_fields, fields,
                            settings_sde=True, diffeqsolve_settings=settings)
    return post
