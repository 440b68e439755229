"""Turn an `ncu --metrics gpu__time_duration.sum --clock-control none --csv` launch list into the markdown table kept under profiles/.

    python tools/launch_summary.py profiles/rN_launches.csv "title" > profiles/rN_launches.md"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = collections.OrderedDict()
for r in rows:
    name = re.sub(r"\(.*", "", r[4]).replace("mpp::", "").strip()
    name = re.sub(r"^void ", "", name)
    t = float(r[14]) * (1e-6 if r[13] == "ns" else (1e-3 if r[13] in ("us", "usecond") else 1.0))
    a = tot.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += t; a[2] = min(a[2], t); a[3] = max(a[3], t)
allms = sum(a[1] for a in tot.values())
print("# %s\n" % sys.argv[2])
print("Source: `%s` (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`; %d launches). Per-launch times under ncu are cold-cache and"
      % (sys.argv[1], len(rows)))
print("serialised: a kernel's SHARE is what must agree with the CUDA-event timings of `bench.py`, not the absolute.\n")
print("| kernel | launches | total ms | share | min – max ms |")
print("|---|---|---|---|---|")
for name, a in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("| `%s` | %d | %.3f | %.1f %% | %.3f – %.3f |" % (name, a[0], a[1], 100.0 * a[1] / allms, a[2], a[3]))
