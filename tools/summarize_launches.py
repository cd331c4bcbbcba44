"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel launches / total / share / avg / max (us)."""
import csv
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1], errors="ignore")) if len(r) > 10]
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = OrderedDict()
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    name = r[ik].split("(")[0]
    a = agg.setdefault(name, [])
    a.append(v / 1e3 if "nsecond" in r[hdr.index("Metric Unit")] or r[hdr.index("Metric Unit")] == "ns" else v)
tot = sum(sum(a) for a in agg.values())
print("sum of kernel durations: %.1f us" % tot)
print("| kernel | launches | total us | share | avg us | max us |\n|---|---|---|---|---|---|")
for k, a in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
    print("| %s | %d | %.1f | %.1f%% | %.1f | %.1f |" % (k, len(a), sum(a), 100 * sum(a) / tot, sum(a) / len(a), max(a)))
if len(sys.argv) > 2:
    for k, a in agg.items():
        if sys.argv[2] in k:
            print(k, " ".join("%.1f" % x for x in a))
