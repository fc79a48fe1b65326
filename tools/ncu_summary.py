"""Print the metrics we care about from an .ncu-rep (run here, no GPU needed)."""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "lts__t_sectors_srcunit_tex_op_red.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__t_sectors_srcunit_tex_op_write.sum", "lts__t_sectors_srcunit_tex.sum",
]
STALL = "smsp__average_warps_issue_stalled_"

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, data = rows[0], rows[1], rows[2:]
kn = hdr.index("Kernel Name")
for r in data:
    print("kernel:", r[kn][:110])
    for w in WANT:
        if w in hdr:
            i = hdr.index(w)
            print("  %-62s %s %s" % (w, r[i], units[i]))
    st = [(float(r[i]), hdr[i][len(STALL):].replace("_per_issue_active.ratio", "")) for i in range(len(hdr))
          if hdr[i].startswith(STALL) and hdr[i].endswith("_per_issue_active.ratio") and r[i]]
    for v, n in sorted(st, reverse=True)[:8]:
        print("  stall %-28s %.3f" % (n, v))
    break
