"""Prints the key counters of every kernel in an ncu raw CSV (ncu -i X.ncu-rep --page raw --csv > raw.csv)."""
import csv
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__grid_size', 'launch__block_size',
        'launch__registers_per_thread', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio']

rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
only_first = '--all' not in sys.argv
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')].split('(')[0]
    if only_first and name in seen:
        continue
    seen.add(name)
    print('==', name)
    for w in WANT:
        if w in hdr:
            print(f"   {w:85s} {r[hdr.index(w)]:>16s} {units[hdr.index(w)]}")
if '--list' in sys.argv:
    print([h for h in hdr if 'stall' in h][:80])
