"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none ... --csv`).

    python tools/launch_summary.py gpurun_out/r02_bench_launches.csv > table.md

Times under ncu are cold-cache and serialised: the SHARE of a kernel in the step is what carries over to the bench run."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr = rows[hi]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    n = re.sub(r"\(.*$", "", r[kn]).replace("<unnamed>::", "").replace("void ", "").strip()
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total ms | avg us | share |")
print("|---|---:|---:|---:|---:|")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("| %s | %d | %.2f | %.2f | %.1f%% |" % (n[:60], c, t / 1e6, t / 1e3 / c, 100.0 * t / tot))
print("| **all** | %d | %.2f | | |" % (sum(v[0] for v in agg.values()), tot / 1e6))
