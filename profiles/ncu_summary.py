"""Key metrics of every kernel in an `ncu --page raw --csv` dump.  usage: python profiles/ncu_summary.py raw.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
keys = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__block_size', 'launch__registers_per_thread', 'launch__shared_mem_per_block_dynamic',
        'smsp__inst_executed.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio']
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print('==', d.get('Kernel Name', '?')[:80])
    for k in keys:
        if k in d and d[k] != '':
            print('  %-90s %s' % (k, d[k]))
