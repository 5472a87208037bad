#!/usr/bin/env python
"""Measure the FP64 FMA-pipe and FP64 tensor-core (DMMA) peaks on cuda:0 with the library's probe kernels."""
import ctypes, json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cd_dynamax_b200 import _lib as L
lib = L.lib()
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream(dev)
out = {}
for name, fn, flops_per in (("fp64_fma", lib.cdk_fma_probe_f64, lambda b, i: 2.0 * 16 * i * b * 256),
                            ("fp64_dmma_m8n8k4", lib.cdk_dmma_probe_f64, lambda b, i: 2.0 * 256 * 8 * i * b * 8)):
    blocks, iters = 148 * 8, 20000
    sink = torch.empty(blocks * 256, dtype=torch.float64, device=dev)
    best = 0.0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        e0.record()
        L.check(fn(blocks, iters, ctypes.c_void_p(sink.data_ptr()), ctypes.c_void_p(stream.cuda_stream)), name)
        e1.record()
        torch.cuda.synchronize()
        best = max(best, flops_per(blocks, iters) / (e0.elapsed_time(e1) * 1e-3) / 1e12)
    out[name + "_tflops"] = best
print(json.dumps(out))
