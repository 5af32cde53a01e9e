"""Pivot an ncu --csv metrics log into one line per launch."""
import csv, sys
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
h = rows[0]
iid, ik, im, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
out = {}
for r in rows[1:]:
    out.setdefault((int(r[iid]), r[ik][:48]), {})[r[im]] = r[iv]
for (i, k), m in sorted(out.items()):
    print(i, k, " ".join(f"{a.split('.')[0].split('__')[-1][:28]}={b}" for a, b in m.items()))
