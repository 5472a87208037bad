"""Batches whose RESULTS do not fit the GPU (or the host): filter them in N-chunks and consume every chunk on the device.

BASELINE config 2 (CD-KF, n = 16, N = 262,144, K = 500) returns 285 GB of filtered / predicted moments -- more than the
180 GB of HBM3e of one B200 (SURVEY section 7 "Output volume").  Trajectories are independent, so the batch axis is cut
into chunks: chunk i + 1 crosses PCIe on a copy stream into the second of two device staging buffers while chunk i is
filtered, and `consume(posterior, lo, hi)` reduces / samples / stores each chunk's device-resident result before its
memory is reused.  The kernels and the C ABI are the ordinary ones (a chunk is the same entry point with a smaller N).
"""
from typing import Callable, Optional

import torch


def filter_in_chunks(filter_fn: Callable, emissions, t_emissions, chunk: int, consume: Callable, inputs=None,
                     device: Optional[torch.device] = None) -> None:
    """Run `filter_fn(emissions_chunk, t_emissions_chunk[, inputs_chunk])` (e.g. `lambda y, t: cdlgssm_filter(params, y,
    t, hp)`) over the leading axis of HOST tensors in chunks of `chunk` trajectories with double-buffered host->device
    copies, and hand every device-resident result to `consume(result, lo, hi)` on the compute stream.

    `emissions` [N, K, m] and `t_emissions` [N, K, 1] are CPU torch tensors (pinned memory makes the copies asynchronous);
    nothing of size O(N) is ever allocated on the device: two staging buffers per input and one chunk of outputs."""
    dev = device or torch.device("cuda", torch.cuda.current_device())
    N = emissions.shape[0]
    if N == 0:
        return
    chunk = max(1, min(int(chunk), N))
    cur = torch.cuda.current_stream(dev)
    copy = torch.cuda.Stream(dev)
    hosts = [emissions, t_emissions] + ([inputs] if inputs is not None else [])
    stage = [[torch.empty((chunk,) + tuple(h.shape[1:]), dtype=h.dtype, device=dev) for h in hosts] for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]  # H2D of slot s has landed
    freed = [torch.cuda.Event(), torch.cuda.Event()]  # the kernels that read slot s have been enqueued and finished
    bounds = list(range(0, N, chunk)) + [N]

    def issue(ci):
        s = ci & 1
        lo, hi = bounds[ci], bounds[ci + 1]
        with torch.cuda.stream(copy):
            if ci >= 2:
                copy.wait_event(freed[s])
            for h, d in zip(hosts, stage[s]):
                d[: hi - lo].copy_(h[lo:hi], non_blocking=True)
            ready[s].record(copy)

    copy.wait_stream(cur)
    issue(0)
    for ci in range(len(bounds) - 1):
        s = ci & 1
        lo, hi = bounds[ci], bounds[ci + 1]
        if ci + 1 < len(bounds) - 1:
            issue(ci + 1)
        cur.wait_event(ready[s])
        args = [d[: hi - lo] for d in stage[s]]
        result = filter_fn(*args)
        consume(result, lo, hi)
        freed[s].record(cur)
        del result
    cur.synchronize()
