"""Host-side marshalling between the reference-shaped Python API and the C ABI (include/cdk.h).

PyTorch is used only for device memory, streams and host<->device copies.  All arithmetic happens in libcdk.so.
"""
import ctypes
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import _lib as L

_TORCH_DT = {"f64": torch.float64, "f32": torch.float32}
_NP_DT = {"f64": np.float64, "f32": np.float32}

# core (per-trajectory) rank of every input slot
_CORE_NDIM = {L.IN_Y: 2, L.IN_T: 1, L.IN_U: 2, L.IN_M0: 1, L.IN_P0: 2, L.IN_F: None, L.IN_B: 1, L.IN_BU: 2,
              L.IN_L: 2, L.IN_QC: 2, L.IN_H: 2, L.IN_D: 1, L.IN_DU: 2, L.IN_R: 2, L.IN_FM: 2, L.IN_FP: 3}


def device():
    if not torch.cuda.is_available():
        raise RuntimeError("cd_dynamax_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def kind_of(x):
    """How the caller holds its data: 'cuda' (torch tensor on the GPU), 'torch' (torch CPU tensor) or 'numpy'."""
    if isinstance(x, torch.Tensor):
        return "cuda" if x.is_cuda else "torch"
    return "numpy"


def pick_dtype(x) -> str:
    """float32 data runs the fp32 kernels, everything else the fp64 kernels (the reference computes in the dtype of
    its inputs: fp32 unless jax_enable_x64)."""
    dt = x.dtype if hasattr(x, "dtype") else np.asarray(x).dtype
    return "f32" if dt in (torch.float32, np.float32, np.dtype("float32")) else "f64"


def to_dev(x, dt: str, dev) -> torch.Tensor:
    tdt = _TORCH_DT[dt]
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=tdt, non_blocking=True).contiguous()
    if hasattr(x, "__cuda_array_interface__"):
        return torch.as_tensor(x, device=dev).to(tdt).contiguous()
    arr = np.ascontiguousarray(np.asarray(x), dtype=_NP_DT[dt])
    return torch.from_numpy(arr).to(dev, non_blocking=True)


def from_dev(t: Optional[torch.Tensor], kind: str):
    if t is None:
        return None
    if kind == "cuda":
        return t
    if kind == "torch":
        return t.cpu()
    return t.cpu().numpy()


def run(entry: str, dt: str, N: int, K: int, n: int, m: int, inputs: Dict[int, object], want: Sequence[int],
        desc_fields: dict, theta_core_ndim: int = 1, scross_rows: Optional[int] = None,
        status: Optional[torch.Tensor] = None) -> Dict[int, torch.Tensor]:
    """Call one C-ABI entry point. `inputs` maps slot -> array-like (None = absent); a leading N marks it batched.
    Returns slot -> device tensor for every slot in `want` (+ OUT_STATUS)."""
    lib = L.lib()
    dev = device()
    d = L.new_desc()
    d.N, d.K, d.n, d.m = N, K, n, m
    for k, v in desc_fields.items():
        setattr(d, k, v)
    keep = []
    in_ptrs = (ctypes.c_void_p * L.NUM_IN)()
    mask = 0
    for slot in range(L.NUM_IN):
        x = inputs.get(slot)
        if x is None:
            in_ptrs[slot] = None
            continue
        t = to_dev(x, dt, dev)
        core = _CORE_NDIM[slot] if _CORE_NDIM[slot] is not None else theta_core_ndim
        if t.dim() == core + 1:
            if t.shape[0] != N:
                raise ValueError(f"input slot {slot}: leading axis {t.shape[0]} != N={N}")
            mask |= 1 << slot
        elif t.dim() != core:
            raise ValueError(f"input slot {slot}: expected rank {core} or {core + 1}, got shape {tuple(t.shape)}")
        keep.append(t)
        in_ptrs[slot] = t.data_ptr() if t.numel() > 0 else None
    d.batched_mask = mask
    tdt = _TORCH_DT[dt]
    shapes = {
        L.OUT_LL: (N,), L.OUT_FM: (N, K, n), L.OUT_FP: (N, K, n, n), L.OUT_PM: (N, K, n), L.OUT_PP: (N, K, n, n),
        L.OUT_LLCUM: (N, K), L.OUT_SM: (N, K, n), L.OUT_SP: (N, K, n, n),
        L.OUT_SCROSS: (N, max(K - 1, 0) if scross_rows is None else scross_rows, n, n),
    }
    out = {}
    out_ptrs = (ctypes.c_void_p * L.NUM_OUT)()
    for slot in range(L.NUM_OUT):
        out_ptrs[slot] = None
    for slot in want:
        t = torch.empty(shapes[slot], dtype=tdt, device=dev)
        out[slot] = t
        out_ptrs[slot] = t.data_ptr() if t.numel() > 0 else None
    if status is None:
        status = torch.zeros((N,), dtype=torch.int32, device=dev)
    out[L.OUT_STATUS] = status
    out_ptrs[L.OUT_STATUS] = status.data_ptr() if N > 0 else None
    nscratch = lib.cdk_scratch_bytes(ctypes.byref(d), entry.encode())
    if nscratch:
        scratch = torch.empty((nscratch,), dtype=torch.uint8, device=dev)
        keep.append(scratch)
        out_ptrs[L.OUT_SCRATCH] = scratch.data_ptr()
    stream = torch.cuda.current_stream(dev).cuda_stream
    fn = getattr(lib, f"{entry}_{dt}")
    rc = fn(ctypes.byref(d), in_ptrs, out_ptrs, ctypes.c_void_p(stream))
    L.check(rc, f"{entry}_{dt}")
    # staged inputs / scratch were allocated on this stream: the caching allocator keeps them valid until the kernel
    # has consumed them (stream-ordered reuse), so dropping `keep` here is safe.
    del keep
    return out


def ll_sum(ll: torch.Tensor) -> torch.Tensor:
    """Deterministic on-device sum of per-trajectory log-likelihoods -> float64 scalar tensor
    (`vmap(marginal_log_prob)(...).sum()`, src/ssm_temissions.py:555-567)."""
    lib = L.lib()
    out = torch.empty((1,), dtype=torch.float64, device=ll.device)
    fn = lib.cdk_ll_sum_f64 if ll.dtype == torch.float64 else lib.cdk_ll_sum_f32
    stream = torch.cuda.current_stream(ll.device).cuda_stream
    L.check(fn(ctypes.c_void_p(ll.data_ptr()), ll.numel(), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(stream)),
            "cdk_ll_sum")
    return out


# ---- diffeqsolve_settings -> {solver, dt0, max_steps} -----------------------------------------------------------------
def parse_settings(settings: Optional[dict], sde: bool = False) -> dict:
    """Map the reference's `diffeqsolve_settings` dict (forwarded verbatim to diffrax, src/utils/diffrax_utils.py:40-57)
    onto the fixed-step solver registry.  Unsupported settings are rejected loudly."""
    settings = dict(settings or {})
    solver = settings.pop("solver", None)
    if solver is None:
        name = "heun" if sde else "dopri5"  # diffrax_utils.py:121-127
    elif isinstance(solver, str):
        name = solver.lower()
    else:
        name = type(solver).__name__.lower()  # works for diffrax solver instances and cd_dynamax_b200.solvers
    if name not in L.SOLVERS:
        raise NotImplementedError(f"solver {name!r} is not in the fixed-step registry {sorted(L.SOLVERS)}")
    ctrl = settings.pop("stepsize_controller", None)
    if ctrl is not None and type(ctrl).__name__ != "ConstantStepSize":
        raise NotImplementedError("only diffrax.ConstantStepSize is supported (adaptive PID controllers are not)")
    settings.pop("adjoint", None)  # gradient checkpointing strategy: irrelevant without autodiff
    settings.pop("tol_vbt", None)
    settings.pop("debug", None)
    dt0 = float(settings.pop("dt0", 0.01))
    max_steps = int(settings.pop("max_steps", 1e5))
    if settings:
        raise NotImplementedError(f"unsupported diffeqsolve_settings: {sorted(settings)}")
    return {"solver": L.SOLVERS[name], "dt0": dt0, "max_steps": max_steps}
