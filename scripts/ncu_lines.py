"""Per-source-line instruction counts of one kernel from an .ncu-rep captured with --import-source on.
usage: ncu_lines.py report.ncu-rep [warps] [min_inst_per_warp]"""
import csv, subprocess, sys
rep = sys.argv[1]
W = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 4.0
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(raw.splitlines()))
d = dict(zip(r[0], r[2]))
for k in ("gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
          "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
          "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum"):
    print(k, d.get(k))
for k in r[0]:
    if "issue_stalled" in k and k.endswith("per_issue_active.ratio") and float(d[k] or 0) > 0.15:
        print("  stall", k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), d[k])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
lines = []
for row in csv.reader(src.splitlines()):
    if len(row) > 1 and row[0] == "File Path":
        cur = row[1].split("/")[-1]
    if len(row) > 8 and row[0].isdigit() and row[2] == "-":
        try:
            lines.append((cur, int(row[0]), int(row[7]), int(row[4]), row[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(l[2] for l in lines)
print("total inst/warp %.1f" % (tot / W))
for f, ln, inst, samp, text in sorted(lines):
    if inst / W > thr:
        print("%-18s %4d %8.1f  %s" % (f[:18], ln, inst / W, text))
