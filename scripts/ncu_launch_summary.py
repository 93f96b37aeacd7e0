"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file X ...).
usage: python scripts/ncu_launch_summary.py launches.csv [skip_first_n_launches]"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = OrderedDict()
for r in rows[skip:]:
    name = re.sub(r"\(.*", "", r[ki].replace("mdb::", "").replace("void ", ""))
    if "<" in r[ki]:
        name = re.sub(r"\((int|bool)\)", "", r[ki].replace("mdb::", "").replace("void ", "").split(">(")[0] + ">")
    ms = float(r[vi].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "nsecond": 1e-6, "msecond": 1.0}.get(r[ui], 1e-6)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += ms
tot = sum(a[1] for a in agg.values())
print("%-64s %6s %12s %12s %7s" % ("kernel", "n", "total ms", "avg ms", "share"))
for name, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %6d %12.3f %12.4f %6.1f%%" % (name[:64], n, ms, ms / n, 100.0 * ms / tot))
print("%-64s %6d %12.3f" % ("total", sum(a[0] for a in agg.values()), tot))
