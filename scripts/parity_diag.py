#!/usr/bin/env python
"""Diagnose the element-wise parity gate on the BASELINE-shape cases: where is the worst entry, how large is it relative to
its own moment's norm, and what do the coarser gates say.  python scripts/parity_diag.py"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import test_gpu_baseline_shapes as B  # noqa: E402
from tests.helpers import FIELD_CORE, gate_err, scaled_err  # noqa: E402
from tests.test_gpu_parity import FIELDS, api, c3_problem, linear_params_api, nonlinear_params_api  # noqa: E402
from oracle import cd_oracle as o  # noqa: E402


def report(tag, post, ref):
    for fld in FIELDS:
        a = np.asarray(getattr(post, fld), np.float64)
        b = np.asarray(ref[fld], np.float64)
        core = FIELD_CORE[fld]
        ax = tuple(range(a.ndim - core, a.ndim))
        s = np.max(np.abs(b), axis=ax, keepdims=True)
        err = np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * s)
        idx = np.unravel_index(np.argmax(err), err.shape)
        out = dict(case=tag, field=fld, elem_1e6=float(err.max()), survey_gate=gate_err(a, b),
                   per_moment=float(np.max(np.abs(a - b) / s)),
                   whole=scaled_err(a, b), worst_index=[int(i) for i in idx], worst_value=float(b[idx]),
                   worst_abs_err=float(abs(a[idx] - b[idx])), own_norm=float(np.broadcast_to(s, b.shape)[idx]))
        print(json.dumps(out), flush=True)


cd = api()
# C3, K = 1000
t, y = c3_problem(2048, 1000)
hp = cd.EKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.0025})
report("c3_k1000", cd.cdnlgssm_filter(nonlinear_params_api(B.L63), y, t[..., None], hp), B._c_oracle_ekf(y, t))
# C4 UKF n = 40
g, po, t, y = B._l96_case(N=6, K=40, seed=40)
hp = cd.UKFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.005})
report("c4_ukf_n40", cd.cdnlgssm_filter(nonlinear_params_api(g), y, t[..., None], hp),
       o.unscented_kalman_filter(po, y, t, settings=o.SolverSettings("rk4", 0.005)))
# C2 KF n = 16, K = 500
from oracle import cpu_baseline as cb  # noqa: E402
g, po, t, y = B._c2_case(64, 500)
hp = cd.KFHyperParams(diffeqsolve_settings={"solver": "rk4", "dt0": 0.01})
r = cb.filter_c("kf", y, t, g["m0"], g["P0"], g["F"], g["L"], g["Qc"], g["H"], g["d"], g["R"], bias=g["b"], solver="rk4", dt0=0.01)
report("c2_kf_k500", cd.cdlgssm_filter(linear_params_api(g), y, t[..., None], hp), r)
