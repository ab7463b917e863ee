"""Per-kernel launch count / total time / share from an `ncu --metrics gpu__time_duration.sum --csv` log.

    python tools/launch_list.py gpurun_out/<tag>_launches.csv
"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    tot[r[ik]] += float(r[iv].replace(",", "")) * scale.get(r[iu], 1e-6)
    cnt[r[ik]] += 1
s = sum(tot.values())
print(f"{'kernel':62s} {'launches':>8s} {'total ms':>10s} {'share':>7s}")
for k, v in tot.most_common():
    print(f"{k[:60]:62s} {cnt[k]:8d} {v:10.3f} {100 * v / s:6.1f}%")
