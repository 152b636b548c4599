#!/usr/bin/env python
"""Summarise an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum` launch list of
one RK substep of bench.py into DRAM bytes per point and launch for each line-kernel class (profiles/ncu_dram_bench_r01.json).

  ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv \
      -k regex:lines2_|poisson_team --launch-skip 60 --launch-count 20 --log-file gpurun_out/dram_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu
  python tools/ncu_dram_summary.py gpurun_out/dram_bench.csv 1024 512 1024
"""
import collections
import csv
import json
import os
import sys


def classify(name):
    # lines2_contig / lines2_strided <MODE, periodic, need_1der>; in the C3 bench y is the only non-periodic direction
    import re
    if "poisson_team_kernel" in name or "poisson_modes_kernel" in name:
        return "poisson_y"
    m = re.search(r"lines2_(contig|strided|march)\w*<(?:\(int\))?(\d+), (?:\(bool\))?(\d), (?:\(bool\))?(\d)", name)
    if not m:
        return None
    kind, mode, per = m.group(1), int(m.group(2)), int(m.group(3))
    d = "x" if kind == "contig" else ("z" if per else "y")      # (march: the second template argument is PER as well)
    return {1: "partial_", 2: "partial_", 3: "partial_", 4: "burgers_", 5: "neumann_"}[mode] + d


def main():
    path, nx, ny, nz = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    n = nx * ny * nz
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    idx = {h: i for i, h in enumerate(rows[hi])}
    per_launch = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(idx):
            continue
        key = (r[idx["ID"]], r[idx["Kernel Name"]])
        v = float(r[idx["Metric Value"]].replace(",", ""))
        unit = r[idx["Metric Unit"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-6, "us": 1e-3, "ms": 1.0,
                 "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0}.get(unit, 1.0)
        per_launch.setdefault(key, {})[r[idx["Metric Name"]]] = v * scale
    agg = collections.OrderedDict()
    for (_, name), m in per_launch.items():
        c = classify(name)
        if c is None:
            continue
        a = agg.setdefault(c, {"launches": 0, "bytes": 0.0, "ms": 0.0})
        a["launches"] += 1
        a["bytes"] += m.get("dram__bytes_read.sum", 0.0) + m.get("dram__bytes_write.sum", 0.0)
        a["ms"] += m.get("gpu__time_duration.sum", 0.0)
    out = {"source": os.path.basename(path), "grid": [nx, ny, nz],
           "bytes_per_point_per_launch": {k: v["bytes"] / v["launches"] / n for k, v in agg.items()},
           "launches": {k: v["launches"] for k, v in agg.items()},
           "ms_per_launch_under_ncu": {k: v["ms"] / v["launches"] for k, v in agg.items()}}
    print(json.dumps(out, indent=1))
    name = sys.argv[5] if len(sys.argv) > 5 else "ncu_dram_bench_r02.json"
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "profiles", name), "w"), indent=1)


if __name__ == "__main__":
    main()
