#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/c3_eks_probe.py
for w in 7 8 12; do CDK_LW_WARPS=$w timeout 300 python scripts/c3_eks_probe.py; done
