"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py: per-kernel share of the LAST
complete Gibbs sweep (launch-list times are cold-cache and serialised: compare shares, not absolutes)."""
import collections
import csv
import sys

path, out = sys.argv[1], sys.argv[2]
rows = [r for r in csv.reader(open(path)) if len(r) > 10]
hdr, rows = rows[0], rows[1:]
ki, vi, gi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Grid Size")
rows = [r for r in rows if "k_peak_" not in r[ki]]          # the FP64 peak microbenchmark runs after the timed region
# a sweep starts with k_fill_normal (draw_f) and ends before the next one
starts = [i for i, r in enumerate(rows) if "k_fill_normal" in r[ki]]
a, b = starts[-2], starts[-1]                               # last COMPLETE sweep [a, b)  (b.. is the final one)
last = rows[b:]
if len(last) < (b - a):                                     # final sweep incomplete in the capture -> use the previous
    last = rows[a:b]
tot = sum(float(r[vi].replace(",", "")) for r in last)
agg = collections.OrderedDict()
for r in last:
    name = r[ki].split("(")[0].replace("void ", "").replace("gpirt::", "").strip()
    agg.setdefault(name, []).append(float(r[vi].replace(",", "")))
with open(out, "w") as fh:
    fh.write("# ncu launch list, one Gibbs sweep (%d launches, %.2f ms serialised cold-cache)\n\n" % (len(last), tot / 1e6))
    fh.write("source: `%s` (command: see profiles/README.md)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n" % path.split("/")[-1])
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        fh.write("| `%s` | %d | %.1f | %.1f%% |\n" % (k, len(v), sum(v) / 1e3, 100 * sum(v) / tot))
print(open(out).read())
