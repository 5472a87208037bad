#!/usr/bin/env python
"""Top stalled SASS instructions of an .ncu-rep (source page): python scripts/ncu_source_top.py rep [N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
tot = sum(int(r[ix["# Samples"]]) for r in body)
print("total samples", tot, "instructions", len(body))
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
agg = {s: sum(int(r[ix[s]]) for r in body) for s in stalls}
print({k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1]) if v})
for i, r in enumerate(body):
    r.append(i)
for r in sorted(body, key=lambda r: -int(r[ix["# Samples"]]))[:top]:
    s = {k: int(r[ix[k]]) for k in stalls if int(r[ix[k]])}
    main = sorted(s.items(), key=lambda kv: -kv[1])[:3]
    print(f'{r[-1]:5d} {int(r[ix["# Samples"]]):6d} exec={r[ix["Instructions Executed"]]:>9} thr={r[ix["Avg. Threads Executed"]]:>5} {r[ix["Source"]].strip()[:70]:70s} {main}')
