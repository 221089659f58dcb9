"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.
    python tools/ncu_launch_summary.py gpurun_out/x_launches.csv [title] > profiles/x_launches_summary.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) < 15:
        continue
    name = r[4].split("(")[0].replace("void ", "")[-70:]
    agg[name][0] += 1
    agg[name][1] += float(r[-1].replace(",", ""))
tot = sum(v[1] for v in agg.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
print(f"launches {sum(v[0] for v in agg.values())}  total {tot / 1e6:.2f} ms (durations are cold-cache and serialised)")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1] / 1e3:12.1f} us {v[0]:6d}  {k}")
