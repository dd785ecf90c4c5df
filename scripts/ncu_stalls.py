#!/usr/bin/env python
"""Summarise the source page of an ncu report (ncu -i X.ncu-rep --page source --csv --launch-skip N --launch-count 1):
stall samples by reason over the whole kernel, executed instructions by opcode, and the instructions with the most samples."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
col = {c: j for j, c in enumerate(H)}
stalls = [c for c in H if c.startswith("stall_") and "Not Issued" not in c]
tot = collections.Counter()
ops = collections.Counter()
samples = []
inst_total = 0
for r in rows[hdr + 1:]:
    if len(r) < len(H):
        continue
    try:
        n = int(r[col["# Samples"]] or 0)
        ex = int(r[col["Instructions Executed"]] or 0)
    except ValueError:
        continue
    src = r[col["Source"]]
    op = re.sub(r"^@!?U?P\d+\s+", "", src).split()[0].split(".")[0] if src else "?"
    ops[op] += ex
    inst_total += ex
    for s in stalls:
        v = r[col[s]]
        if v:
            tot[s] += int(v)
    samples.append((n, ex, r[col["Address"]], src, {s: int(r[col[s]]) for s in stalls if r[col[s]] and int(r[col[s]])}))
allsamp = sum(tot.values())
print("warp instructions executed: %.1f M; stall samples: %d" % (inst_total / 1e6, allsamp))
for s, v in tot.most_common():
    print("  %-24s %6.2f %%" % (s, 100.0 * v / allsamp))
print("executed by opcode:")
for op, v in ops.most_common(22):
    print("  %-10s %8.2f M %5.1f %%" % (op, v / 1e6, 100.0 * v / inst_total))
print("top instructions by samples:")
for n, ex, a, src, d in sorted(samples, reverse=True)[:top]:
    print("  %6d %5.2f%% %s  %-50s %s" % (n, 100.0 * n / allsamp, a[-5:], src[:50], dict(sorted(d.items(), key=lambda kv: -kv[1])[:3])))
