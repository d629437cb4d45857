"""Filter `ncu --page raw --csv` output down to the metrics named by a regex (one line per
metric per kernel launch)."""
import csv
import re
import sys

pat = re.compile(sys.argv[1])
rows = list(csv.reader(sys.stdin))
if len(rows) < 3:
    sys.exit(0)
header, units = rows[0], rows[1]
for r in rows[2:]:
    rec = dict(zip(header, r))
    print("== kernel:", rec.get("Kernel Name", "?")[:110], "| grid", rec.get("Grid Size"), "| block", rec.get("Block Size"))
    for h, u, v in zip(header, units, r):
        if pat.search(h):
            print("   %-75s %s %s" % (h, v, u))
