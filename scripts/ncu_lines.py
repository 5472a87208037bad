#!/usr/bin/env python
"""Samples per CUDA source line of an .ncu-rep (needs -lineinfo + --import-source on): python scripts/ncu_lines.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; hdr = None; res = []
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ix = {h: i for i, h in enumerate(hdr)}; continue
    if hdr and len(r) == len(hdr) and r[2] == "-":  # a CUDA source line (aggregated over its SASS)
        s = int(r[ix["# Samples"]] or 0)
        if s:
            st = {h: int(r[i]) for i, h in enumerate(hdr) if h.startswith("stall_") and r[i] not in ("", "0") and "Not" not in h}
            res.append((s, cur, int(r[0]), int(r[ix["Instructions Executed"]] or 0), r[1].strip()[:90], sorted(st.items(), key=lambda kv: -kv[1])[:2]))
tot = sum(x[0] for x in res)
print("total samples", tot)
for x in sorted(res, key=lambda x: -x[0])[:top]:
    print(f"{100 * x[0] / tot:5.1f}% {x[1]}:{x[2]:<5d} exec={x[3]:>9d} {x[4]:90s} {x[5]}")
