"""Print selected metrics per kernel from `ncu -i rep --page raw --csv`.  usage: ncu -i x.ncu-rep --page raw --csv | python tools/ncu_raw.py"""
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = rows[0]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.sum.pct',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared',
        'gpu__dram_throughput.avg.pct', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__average_warps_issue_stalled_long_scoreboard_per', 'smsp__average_warps_issue_stalled_short_scoreboard_per',
        'smsp__average_warps_issue_stalled_mio_throttle_per', 'smsp__average_warps_issue_stalled_barrier_per',
        'smsp__average_warps_issue_stalled_math_pipe', 'smsp__average_warps_issue_stalled_not_selected_per',
        'smsp__average_warps_issue_stalled_wait_per', 'smsp__average_warps_issue_stalled_dispatch',
        'smsp__average_warps_issue_stalled_lg_throttle_per', 'smsp__average_warps_issue_stalled_no_instruction_per']
ki = hdr.index('Kernel Name')
for r in rows[2:]:
    print('=====', r[ki][:90])
    for h, v in zip(hdr, r):
        if any(h.startswith(w) for w in want) and 'not_issued' not in h and 'per_second' not in h and '.pct_of_peak_sustained_elapsed' not in h.replace('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','x').replace('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed','x'):
            print('  ', h.replace('smsp__average_warps_issue_stalled_', 'stall_').replace('_per_issue_active.ratio', ''), v[:14])
