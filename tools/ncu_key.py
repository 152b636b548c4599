#!/usr/bin/env python
"""Key metrics of an .ncu-rep (last kernel in the report): python tools/ncu_key.py file.ncu-rep [substring ...]"""
import csv, subprocess, sys
rep = sys.argv[1]
extra = sys.argv[2:]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_lg',
        'lts__throughput.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct', 'smsp__issue_active.avg.pct', 'smsp__inst_executed.sum',
        'smsp__average_warps_issue_stalled', 'l1tex__data_bank_conflicts', 'smsp__inst_executed_op_shared', 'l1tex__t_sectors_pipe_lsu_mem_local',
        'sm__inst_executed_pipe_lsu', 'launch__shared_mem', 'launch__grid_size', 'launch__block_size', 'sm__ctas_launched', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'l1tex__t_sector_hit_rate', 'lts__t_sector_hit_rate'] + extra
for v in rows[2:]:
    print("==", v[h.index("Kernel Name")][:80] if "Kernel Name" in h else "")
    for i, n in enumerate(h):
        if any(w in n for w in want) and "pcsamp" not in n and "not_issued" not in n:
            try:
                x = float(v[i])
                if "stalled" in n and x < 0.3:
                    continue
            except ValueError:
                pass
            print("  %-90s %s" % (n, v[i]))
