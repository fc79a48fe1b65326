"""Summarise svdfeature_b200/build_ptxas.log: registers / spills / stack per kernel."""
import re
import subprocess
import sys

log = open(sys.argv[1] if len(sys.argv) > 1 else "svdfeature_b200/build_ptxas.log").read()
pat = re.compile(r"Compiling entry function '(\S+)' for 'sm_100a'\n(.*?)(?=ptxas info    : Compiling entry|\Z)", re.S)
rows = []
for name, body in pat.findall(log):
    dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    dem = re.sub(r"\(.*", "", dem).replace("void svdk::", "")
    regs = re.search(r"Used (\d+) registers", body)
    spill = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", body)
    smem = re.search(r"(\d+) bytes smem", body)
    rows.append((dem, int(regs.group(1)) if regs else -1, int(spill.group(2)) if spill else 0,
                 int(spill.group(1)) if spill else 0, int(smem.group(1)) if smem else 0))
flt = sys.argv[2] if len(sys.argv) > 2 else ""
for r in sorted(rows):
    if flt in r[0]:
        print("%-46s regs=%3d spill_st=%4dB stack=%4dB smem=%6dB" % r)
