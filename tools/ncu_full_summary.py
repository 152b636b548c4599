"""Key metrics per kernel launch from `ncu -i <rep> --page raw --csv` (the reports themselves exceed what comes back from the box).
usage: python tools/ncu_full_summary.py <raw.csv> <out.json> "<capture command>" "<reading>" """
import csv
import json
import sys

KEYS = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'lts__t_sector_hit_rate.pct',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
kern = []
for r in data:
    k = {"Kernel Name": r[ix["Kernel Name"]].replace("tlab::<", "")}
    for key in KEYS:
        if key in ix:
            k[key] = "%s %s" % (r[ix[key]], units[ix[key]])
    kern.append(k)
json.dump({"capture": sys.argv[3], "reading": sys.argv[4], "kernels": kern}, open(sys.argv[2], "w"), indent=0)
for k in kern:
    print(k["Kernel Name"][:60], k.get('gpu__time_duration.sum'), k.get('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'),
          'bar', k.get('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio'),
          'lsb', k.get('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio'),
          'spill', k.get('l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum'), 'inst', k.get('smsp__inst_executed.sum'))
