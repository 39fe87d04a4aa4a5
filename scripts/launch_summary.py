"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel count, total and mean ms."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    name = r[ki].split("(")[0][-48:]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += float(r[vi].replace(",", "")) / 1e6
tot = sum(a[1] for a in agg.values())
for name, (n, ms) in agg.items():
    print("%-50s n=%3d total %9.3f ms  mean %8.3f ms  %5.1f%%" % (name, n, ms, ms / n, 100 * ms / tot))
