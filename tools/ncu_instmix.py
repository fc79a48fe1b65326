"""Instruction mix + hottest stall lines from `ncu --page source --csv` of an .ncu-rep."""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# first line is the kernel name; the header follows
h = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[h]
si, ei, st = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
tot, stall = collections.Counter(), []
for r in rows[h + 1:]:
    if len(r) <= ei or r[0] == "Address":
        break
    try:
        n = int(r[ei])
    except ValueError:
        continue
    toks = r[si].split()
    op = toks[1] if toks and toks[0].startswith("@") else (toks[0] if toks else "")
    tot[".".join(op.split(".")[:2]) if op.startswith(("LD", "ST", "RED", "ATOM")) else op.split(".")[0]] += n
    stall.append((int(r[st] or 0), r[si].strip()[:90]))
s = sum(tot.values())
print("total warp instructions", s)
for op, n in tot.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 28):
    print("%-12s %12d %5.1f%%" % (op, n, 100.0 * n / s))
print("-- hottest stall samples")
for n, src in sorted(stall, reverse=True)[:14]:
    print("%7d  %s" % (n, src))
