"""Print the SASS instructions with the most warp-stall samples from `ncu --page source --csv`."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si = hdr.index("Warp Stall Sampling (All Samples)")
ie = hdr.index("Instructions Executed")
data = []
for r in rows[2:]:
    if r and r[0] == "Kernel Name":   # next launch of the capture
        break
    data.append(r)
tot = sum(int(r[si] or 0) for r in data)
print("total samples", tot, "instructions", len(data))
top = sorted(range(len(data)), key=lambda i: -int(data[i][si] or 0))[: int(sys.argv[2]) if len(sys.argv) > 2 else 25]
for i in sorted(top):
    r = data[i]
    print("%5d %6.2f%% exec=%-8s %s" % (i, 100.0 * int(r[si]) / tot, r[ie], r[1].strip()[:110]))
