"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: totals by kernel and the last render in launch order.
usage: python profiles/launch_summary.py launches.csv [first-kernel-of-a-render]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = rows[0]
ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
launches = []
for r in rows[1:]:
    try:
        v = float(r[iv].replace(",", ""))
    except ValueError:
        continue
    v = v / 1000.0 if r[iu] in ("ns", "nsecond") else v
    name = re.sub(r"^void ", "", r[ik])
    name = re.sub(r"\(.*", "", name).replace("ars::", "")
    launches.append((name, v))
tot = collections.OrderedDict()
for n, v in launches:
    t = tot.setdefault(n, [0.0, 0])
    t[0] += v
    t[1] += 1
total = sum(t[0] for t in tot.values())
for n, (v, c) in sorted(tot.items(), key=lambda kv: -kv[1][0]):
    print(f"{v:10.1f} us total  n={c:3d}  avg {v / c:8.1f} us  {100 * v / total:5.1f}%  {n[:90]}")
first = sys.argv[2] if len(sys.argv) > 2 else "ir_scatter_kernel"
starts = [i for i, (n, _) in enumerate(launches) if n.startswith(first)]
if len(starts) >= 2:
    seg = launches[starts[-2]:starts[-1]]
    print(f"\none render, in launch order (from {first}):")
    for n, v in seg:
        print(f"  {v:8.1f} us  {n[:100]}")
    print(f"total {sum(v for _, v in seg):.1f} us")
