#!/usr/bin/env python3
"""Per-source-line share of executed warp instructions and stall samples from an ncu report
(ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > file; python tools/ncu_lines.py file [top])."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 60
cur = None; per = collections.Counter(); samp = collections.Counter(); ie = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r and r[0] == "Line No": ie = r.index("Instructions Executed"); ss = r.index("# Samples"); continue
    if ie is None or len(r) <= ie: continue
    if r[0] != "" and r[2] == "-":
        k = (cur, int(r[0]), r[1].strip()[:80])
        per[k] += int(r[ie]); samp[k] += int(r[ss])
tot = sum(per.values()); st = sum(samp.values())
print("total warp instr", tot, "samples", st)
byfile = collections.Counter()
for k, v in per.items(): byfile[k[0]] += v
print({k: "%.1f%%" % (100 * v / tot) for k, v in byfile.items()})
for k, v in per.most_common(top):
    print("%5.1f%% instr %5.1f%% samples  %s:%d  %s" % (100 * v / tot, 100 * samp[k] / st, k[0], k[1], k[2]))
