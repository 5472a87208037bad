"""Trajectory sharding across the GPUs of one box.

Trajectories are independent (the reference's only cross-trajectory operation is `vmap(marginal_log_prob)(...).sum()`,
src/ssm_temissions.py:555-567), so the batch axis N is cut into contiguous slices, one per rank (one process per GPU),
every rank filters its slice with no data-path communication, and the summed log-likelihood is combined with ONE
all-reduce of a single float64 (NCCL over NVLink on GPUs; `gloo` in the CPU tests).  Hosts that do not use
torch.distributed call `cdk_ll_allreduce(ncclComm_t, ...)` from include/cdk.h instead.
"""
from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_traj: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[start, stop) of the contiguous slice of trajectories owned by `rank`; the first n_traj % world_size ranks get one
    extra trajectory, empty slices are allowed (n_traj < world_size)."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of size {world_size}")
    base, extra = divmod(int(n_traj), int(world_size))
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(x, rank: int, world_size: int):
    """Slice the leading (trajectory) axis of an array-like for this rank; None passes through."""
    if x is None:
        return None
    s, e = shard_bounds(x.shape[0], rank, world_size)
    return x[s:e]


def allreduce_loglik(local_sum: torch.Tensor, group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Sum a float64 scalar tensor over the ranks, in place (the analogue of `.sum()` at ssm_temissions.py:567)."""
    if local_sum.dtype != torch.float64 or local_sum.numel() != 1:
        raise ValueError("allreduce_loglik expects one float64 value")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(local_sum, op=dist.ReduceOp.SUM, group=group)
    return local_sum


def sharded_marginal_log_prob(filter_fn: Callable, emissions, t_emissions, inputs=None,
                              group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Total marginal log-likelihood of a batch that is replicated on (or addressable by) every rank.

    `filter_fn(emissions, t_emissions, inputs, rng_offset)` is called once on this rank's contiguous slice and must return the
    per-trajectory log-likelihoods of that slice (e.g. `lambda y, t, u, off: cdnlgssm_filter(params, y, t, hp, u,
    output_fields=[]).marginal_loglik`; `rng_offset` = index of the slice's first trajectory, which the EnKF needs so that
    its random stream does not depend on the sharding).  Returns a float64 scalar tensor holding the sum over ALL ranks.
    """
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    s, e = shard_bounds(emissions.shape[0], rank, world)
    ll = filter_fn(emissions[s:e], None if t_emissions is None else t_emissions[s:e],
                   None if inputs is None else inputs[s:e], s)
    ll = ll if isinstance(ll, torch.Tensor) else torch.as_tensor(ll)
    if ll.is_cuda:
        from . import _engine as E
        total = E.ll_sum(ll.contiguous()).reshape(1) if ll.numel() > 0 else torch.zeros(1, dtype=torch.float64, device=ll.device)
    else:
        total = ll.to(torch.float64).sum().reshape(1)
    return allreduce_loglik(total, group)[0]


# ---- cdk_ll_allreduce with a raw ncclComm_t (hosts that do not use torch.distributed: the JAX process of INTEGRATION.md) --
class RawNcclComm:
    """An `ncclComm_t` created directly on libnccl (the copy torch bundles is already in the process), i.e. the kind of
    communicator a non-torch host hands to `cdk_ll_allreduce` (include/cdk.h).  The 128-byte ncclUniqueId of rank 0 has to
    reach the other ranks out of band: pass `unique_id` (bytes), or leave it None to have it broadcast through an
    initialised torch.distributed group.  One communicator per process / GPU."""

    def __init__(self, rank: int, world_size: int, device=None, unique_id: Optional[bytes] = None, group=None):
        import ctypes

        class _UniqueId(ctypes.Structure):
            _fields_ = [("internal", ctypes.c_char * 128)]

        self._nccl = ctypes.CDLL("libnccl.so.2")
        if device is not None:
            torch.cuda.set_device(device)
        uid = _UniqueId()
        if unique_id is None:
            if rank == 0:
                rc = self._nccl.ncclGetUniqueId(ctypes.byref(uid))
                if rc != 0:
                    raise RuntimeError(f"ncclGetUniqueId failed ({rc})")
            if world_size > 1:
                if not (dist.is_available() and dist.is_initialized()):
                    raise ValueError("world_size > 1 needs unique_id bytes or an initialised torch.distributed group")
                box = [bytes(uid) if rank == 0 else None]
                dist.broadcast_object_list(box, src=0, group=group)
                ctypes.memmove(ctypes.byref(uid), box[0], 128)
        else:
            ctypes.memmove(ctypes.byref(uid), unique_id, 128)
        self.unique_id = bytes(uid)
        self._comm = ctypes.c_void_p()
        self._nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _UniqueId, ctypes.c_int]
        rc = self._nccl.ncclCommInitRank(ctypes.byref(self._comm), int(world_size), uid, int(rank))
        if rc != 0:
            raise RuntimeError(f"ncclCommInitRank failed ({rc})")
        self.rank, self.world_size = rank, world_size

    @property
    def handle(self) -> int:
        if self._comm is None:
            raise RuntimeError("communicator already destroyed")
        return self._comm.value

    def destroy(self):
        if self._comm is not None and self._comm.value:
            self._nccl.ncclCommDestroy(self._comm)
        self._comm = None


def allreduce_loglik_nccl(local_sum: torch.Tensor, comm: RawNcclComm) -> torch.Tensor:
    """In-place sum of one float64 device value over the ranks of a raw NCCL communicator through the C ABI's
    `cdk_ll_allreduce` (enqueued on the current stream; no host synchronisation)."""
    import ctypes

    from . import _lib as L
    if local_sum.dtype != torch.float64 or local_sum.numel() != 1 or not local_sum.is_cuda:
        raise ValueError("allreduce_loglik_nccl expects one float64 value on the GPU")
    stream = torch.cuda.current_stream(local_sum.device).cuda_stream
    L.check(L.lib().cdk_ll_allreduce(ctypes.c_void_p(comm.handle), ctypes.c_void_p(local_sum.data_ptr()),
                                     ctypes.c_void_p(stream)), "cdk_ll_allreduce")
    return local_sum
