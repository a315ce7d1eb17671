"""print one steady-state step of an ncu launch list (csv from --metrics gpu__time_duration.sum): kernel, microseconds"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
seq = [(r[ki], float(r[vi].replace(",", ""))) for r in rows[hdr + 1:] if len(r) > vi]
anchor = sys.argv[2] if len(sys.argv) > 2 else "cand_count_kernel"
idx = [i for i, (k, v) in enumerate(seq) if anchor in k]
which = int(sys.argv[3]) if len(sys.argv) > 3 else len(idx) - 2
a, b = idx[which], idx[which + 1]
tot = 0.0
for k, v in seq[a:b]:
    name = re.sub(r"\(.*", "", k).replace("void ", "")
    print("%-72s %8.1f us" % (name[:72], v / 1000))
    tot += v
print("total %.1f us, %d launches" % (tot / 1000, b - a))
