#!/usr/bin/env python
"""Static instruction count of the hot loop of a kernel: dump the SASS of one function of the built library
(cuobjdump), find the innermost loop that contains the Philox multiplies (IMAD.WIDE.U32 / IMAD.HI.U32) and print its
length and opcode histogram.  Development aid for the issue-bound photon kernels (no GPU needed).

    python scripts/sass_loop.py [mangled-name-substring] [--lib path]
"""
import collections
import re
import subprocess
import sys

lib = "physicl_b200/libphysicl_b200.so"
pat = "pcl_k_photon_multiILb0ELb0ELb0ELb0ELb1E"
args = [a for a in sys.argv[1:] if a != "--dump"]
if "--lib" in args:
    i = args.index("--lib")
    lib = args[i + 1]
    del args[i:i + 2]
if args:
    pat = args[0]
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
funcs = re.split(r"\n\s*Function : ", out)
body = [f for f in funcs if f.split("\n", 1)[0].find(pat) >= 0]
if not body:
    sys.exit("no function matches " + pat)
body = body[0]
name = body.split("\n", 1)[0]
ins = []
for line in body.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
addr_index = {a: k for k, (a, _) in enumerate(ins)}
loops = []
for k, (a, txt) in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", txt)
    if m:
        tgt = int(m.group(1), 16)
        if tgt <= a and tgt in addr_index:
            loops.append((addr_index[tgt], k))
print(name)
print("instructions in function:", len(ins), " loops (start,end,len):", [(s, e, e - s + 1) for s, e in loops])


def is_philox(t):
    return "IMAD.WIDE.U32" in t or "IMAD.HI.U32" in t


inner = [(s, e) for s, e in loops if not any((s2 > s or e2 < e) and s2 >= s and e2 <= e for s2, e2 in loops)]
cands = [(sum(is_philox(t) for _, t in ins[s:e + 1]), s, e) for s, e in inner]
if not cands or max(cands)[0] == 0:
    sys.exit("no loop with Philox multiplies")
_, s, e = max(cands)
hist = collections.Counter()
for _, t in ins[s:e + 1]:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    op = t.split()[0]
    hist[op.split(".")[0]] += 1
n = e - s + 1
print("hot loop: %d instructions per thread-iteration (4 photons) = %.1f per photon-step" % (n, n / 4.0))
for op, c in hist.most_common():
    print("  %-10s %4d  %5.1f%%" % (op, c, 100.0 * c / n))
if "--dump" in sys.argv:
    for a, t in ins[s:e + 1]:
        print("%05x  %s" % (a, t))
