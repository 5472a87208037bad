#!/usr/bin/env python
"""Grep-able SASS evidence per kernel of libcdk.so: instruction counts of the mnemonics that prove which hardware path a
kernel uses (DFMA / DADD / DMUL = FP64 FMA pipe; DMMA = FP64 tensor cores; UTMASTG = TMA tensor stores; LDGSTS = cp.async;
MUFU = special-function unit), plus registers / spills from cuobjdump --dump-resource-usage.
    python scripts/sass_summary.py > profiles/rNN_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cd_dynamax_b200", "lib", "libcdk.so")
KEYS = ["DFMA", "DADD", "DMUL", "DMMA", "UTMASTG", "LDGSTS", "MUFU", "FFMA", "IMAD", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "LDL", "STL"]
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
res = subprocess.run(["cuobjdump", "--dump-resource-usage", LIB], capture_output=True, text=True).stdout
usage = {}
for mfn, mreg in re.findall(r"Function (\S+):\s*\n\s*REG:(\d+)[^\n]*", res):
    usage[mfn] = int(mreg)
counts, order, cur = {}, [], None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur:
        counts[cur][m.group(1).split(".")[0]] += 1
demangle = subprocess.run(["c++filt"] + order, capture_output=True, text=True).stdout.splitlines()
print("# SASS instruction counts per kernel of cd_dynamax_b200/lib/libcdk.so (cuobjdump -sass, sm_100a)")
print("# kernel | registers | total instructions | " + " ".join(KEYS))
want = sys.argv[1:] or ["ekf_small_lw<double", "eks_small_lw<double", "kf_warp_filter", "kf_warp_smooth", "generic_filter_kernel<double",
                        "generic_smooth_kernel<double", "enkf_kernel<double", "sample_path_kernel<double", "emission_moments_kernel<double",
                        "ekf_l63_grad", "ll_sum_kernel<double"]
for fn, name in zip(order, demangle):
    short = re.sub(r"\(anonymous namespace\)::", "", name)
    short = re.sub(r"cdk::", "", short)
    if not any(w in short for w in want):
        continue
    if "ekf_small_lw<double" in short and not ("(cdk_solver)5" in name or ", 5," in short or "5, 14" in short):
        pass
    c = counts[fn]
    print(f"{short[:150]} | {usage.get(fn, '?')} | {sum(c.values())} | " + " ".join(f"{k}={c.get(k, 0)}" for k in KEYS))
