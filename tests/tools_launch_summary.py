"""Diagnostics (not a test): per-kernel totals of the LAST training step in an ncu launch list of `bench.py --workload train`
(a step ends with the Adam launch).  usage: python tests/tools_launch_summary.py launches.csv"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [r for r in rows[hdr + 2:] if len(r) > mv]
adam = [i for i, r in enumerate(data) if "adam_kernel" in r[kn]]
ends = [adam[i] for i in range(len(adam)) if i == len(adam) - 1 or adam[i + 1] != adam[i] + 1]
lo, hi = ends[-2] + 1, ends[-1] + 1
agg, cnt = collections.Counter(), collections.Counter()
for r in data[lo:hi]:
    k = r[kn][:100]
    agg[k] += float(r[mv].replace(",", ""))
    cnt[k] += 1
tot = sum(agg.values())
print("ncu launch list %s: launches %d..%d = the last train step (times under ncu are serialised and cold-cache: shares, not absolutes)" % (sys.argv[1], lo, hi))
for k, v in agg.most_common(40):
    print("%8.3f ms %5.1f%% x%4d %s" % (v / 1e6, 100 * v / tot, cnt[k], k))
print("total %.3f ms over %d launches" % (tot / 1e6, sum(cnt.values())))
