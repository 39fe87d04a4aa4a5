"""Per-source-line stall samples from an ncu report captured with --import-source on (-lineinfo build)."""
import csv, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = next(r for r in rows if len(r) > 10 and r[0] == "Line No")
si = hdr.index("# Samples"); ii = hdr.index("Instructions Executed")
names = {h: i for i, h in enumerate(hdr)}
agg = []
for r in rows:
    if len(r) > 10 and r[0].isdigit() and r[2] == "-":
        try: agg.append((int(r[si]), int(r[0]), r[1].strip(), int(r[ii]), r))
        except ValueError: pass
tot = sum(a[0] for a in agg)
print("total samples", tot)
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for n, ln, src, inst, r in sorted(agg, reverse=True)[:top]:
    st = sorted(((int(r[names[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
    print("%5.1f%% L%-4d inst=%-10d %-22s %s" % (100.0 * n / tot, ln, inst, ",".join("%s:%d" % (c[6:], v) for v, c in st), src[:95]))
