"""Diagnostics (not a test): join the TTS_GEMM_LOG call log with an ncu launch list of the same run (by order) and print the
bf16 GEMM launches of the LAST train step grouped by shape / epilogue.  usage: python tests/tools_gemm_join.py log.txt launches.csv"""
import collections
import csv
import sys

log = [l.strip() for l in open(sys.argv[1]) if l.startswith("M=")]
rows = list(csv.reader(open(sys.argv[2])))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv = h.index("Kernel Name"), h.index("Metric Value")
data = [r for r in rows[hdr + 2:] if len(r) > mv]
gem = [(i, float(r[mv].replace(",", "")) / 1e3) for i, r in enumerate(data) if "gemm_bf16" in r[kn]]
assert len(gem) == len(log), (len(gem), len(log))
adam = [i for i, r in enumerate(data) if "adam_kernel" in r[kn]]
ends = [adam[i] for i in range(len(adam)) if i == len(adam) - 1 or adam[i + 1] != adam[i] + 1]
lo, hi = ends[-2] + 1, ends[-1] + 1
agg, cnt = collections.Counter(), collections.Counter()
for (i, us), l in zip(gem, log):
    if lo <= i < hi:
        agg[l] += us
        cnt[l] += 1
tot = sum(agg.values())
for l, us in agg.most_common():
    f = dict(kv.split("=") for kv in l.split())
    fl = 2.0 * int(f["M"]) * int(f["N"]) * int(f["K"]) * int(f["taps"])
    print("%8.1f us %5.1f%% x%3d  %6.1f us each %6.0f TFLOP/s  %s" % (us, 100 * us / tot, cnt[l], us / cnt[l], fl * cnt[l] / us / 1e6, l))
print("total %.2f ms over %d launches" % (tot / 1e3, sum(cnt.values())))
