#!/usr/bin/env python
"""Group SASS instructions of an .ncu-rep by execution count (= loop nest) with stall samples: python scripts/ncu_groups.py rep [N]"""
import csv, io, subprocess, collections, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 10
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
body = [r for r in rows[2:] if len(r) == len(hdr)]
groups = collections.defaultdict(collections.Counter); tot = collections.Counter(); samp = collections.Counter()
st = collections.defaultdict(collections.Counter)
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in body:
    e = int(r[ix["Instructions Executed"]]); op = r[ix["Source"]].split()
    op = [o for o in op if not o.startswith('@')][0].split('.')[0]
    groups[e][op] += 1; tot[e] += 1; samp[e] += int(r[ix["# Samples"]])
    for s_ in stalls: st[e][s_] += int(r[ix[s_]])
print("total samples", sum(samp.values()), "total warp-instr %.3f G" % (sum(e * tot[e] for e in tot) / 1e9))
for e in sorted(tot, key=lambda e: -samp[e])[:top]:
    print(e, tot[e], 'instr -> %.3f G;' % (e * tot[e] / 1e9), 'samples', samp[e], dict(groups[e].most_common(8)), dict(st[e].most_common(5)))
