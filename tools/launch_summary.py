"""Summary of an ncu launch list (gpu__time_duration.sum CSV): launches, time and share per kernel.
usage: python tools/launch_summary.py <launches.csv> <out.json> [substeps]"""
import collections
import csv
import json
import re
import sys

src, out = sys.argv[1], sys.argv[2]
nsub = int(sys.argv[3]) if len(sys.argv) > 3 else None
lines = [l for l in open(src) if not l.startswith("==")]
agg = collections.OrderedDict()
n = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", row["Kernel Name"])[:100]
    v = float(row["Metric Value"].replace(",", "")) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[row["Metric Unit"]]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
    n += 1
tot = sum(a[1] for a in agg.values())
rec = {"source": src.split("/")[-1] + ": tools/gpu_launches_r2.sh (ncu --metrics gpu__time_duration.sum --clock-control none, -c 400, "
       "torch's element-wise kernels of the field synthesis and the one-off Poisson set-up kernels excluded by base name; "
       "python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-parity)",
       "note": "per-launch times under ncu are cold-cache and serialised: the SHARES are what should agree with bench.py's "
               "breakdown_ms, not the absolutes",
       "launches": n, "total_ms": round(tot, 3),
       "kernels": [{"kernel": k, "launches": a[0], "ms": round(a[1], 3), "share": round(a[1] / tot, 4)}
                   for k, a in sorted(agg.items(), key=lambda x: -x[1][1])]}
if nsub:
    rec["substeps_covered"] = nsub
    rec["ms_per_substep"] = round(tot / nsub, 3)
    rec["launches_per_substep"] = round(n / nsub, 2)
json.dump(rec, open(out, "w"), indent=1)
print(json.dumps(rec)[:600])
