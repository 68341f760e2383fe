"""Average gpu__time_duration per kernel name from an `ncu --csv --metrics gpu__time_duration.sum` log."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
acc = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split("(")[0].replace("effdet::", "").replace("void ", "")
    a = acc.setdefault(k, [0.0, 0]); a[0] += float(r[vi].replace(",", "")); a[1] += 1
for k, (t, n) in acc.items():
    print("%-60s n=%4d avg %9.2f us  total %10.1f us" % (k[:60], n, t / n / 1e3, t / 1e3))
