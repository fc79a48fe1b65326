"""Per-instruction view of one kernel of an .ncu-rep (source page, SASS): stall-sample totals by
reason and the instructions that collect the most samples.  Used on k_own with every rating on
ONE item (tools/own_study.py ... ni=1): a single owner warp is then active, so the samples are the
stall profile of the chain link itself.
usage: python tools/ncu_source_top.py gpurun_out/x.ncu-rep [marker-instruction] [min-executions] [all]
("all": every instruction of the hot loop in program order, not only the 40 with the most samples)"""
import csv
import subprocess
import sys

rep = sys.argv[1]
marker = sys.argv[2] if len(sys.argv) > 2 else "TRYWAIT"
min_ex = int(sys.argv[3]) if len(sys.argv) > 3 else 100000
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                     text=True).stdout
rows = list(csv.reader(raw.splitlines()))
print(rows[0][1])
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
start = next((n for n, r in enumerate(data) if marker in r[ix["Source"]]), 0)
part = [(n, r) for n, r in enumerate(data) if n >= start - 40]
tot = sum(int(r[ix["# Samples"]]) for _, r in part)
agg = {}
for _, r in part:
    for c in stall_cols:
        agg[c] = agg.get(c, 0) + int(r[ix[c]])
print("samples from the first %s on: %d" % (marker, tot))
print("by reason:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]))
hot = [(n, r) for n, r in part if int(r[ix["Instructions Executed"]]) >= min_ex]
per = sum(int(r[ix["Instructions Executed"]]) for _, r in hot) / max(1, max(int(r[ix["Instructions Executed"]]) for _, r in hot))
print("instructions executed >= %d times: %d (%.1f per trip of the hottest)" % (min_ex, len(hot), per))
show = hot if "all" in sys.argv[4:] else sorted(hot, key=lambda x: -int(x[1][ix["# Samples"]]))[:40]
for n, r in sorted(show):
    st = {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]]) > 0}
    top = sorted(st.items(), key=lambda x: -x[1])[:2]
    print("%5d  %-60s %6s %9s  %s" % (n, r[ix["Source"]].strip()[:60], r[ix["# Samples"]], r[ix["Instructions Executed"]], top))
