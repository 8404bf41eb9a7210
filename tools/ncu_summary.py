#!/usr/bin/env python3
"""Summarise the key metrics of every kernel in an .ncu-rep (raw page) -- used for profiles/*.md."""
import csv
import subprocess
import sys

out = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = [('gpu__time_duration.sum', 'time'), ('dram__bytes_read.sum', 'dram rd'), ('dram__bytes_write.sum', 'dram wr'),
        ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram %'), ('lts__t_bytes.sum', 'L2 bytes'),
        ('sm__throughput.avg.pct_of_peak_sustained_elapsed', 'SM %'), ('sm__warps_active.avg.pct_of_peak_sustained_active', 'occupancy %'),
        ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue %'), ('launch__registers_per_thread', 'regs'),
        ('launch__grid_size', 'grid'), ('launch__block_size', 'block'),
        ('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 'stall long_sb'),
        ('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'stall barrier'),
        ('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 'stall short_sb'),
        ('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'stall mio'),
        ('smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio', 'stall lg'),
        ('smsp__average_warps_issue_stalled_wait_per_issue_active.ratio', 'stall wait'),
        ('smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio', 'stall not_sel'),
        ('smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio', 'stall math'),
        ('smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio', 'stall no_inst'),
        ('smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio', 'stall branch'),
        ('smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio', 'stall dispatch'),
        ('l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smem conflicts'), ('smsp__inst_executed.sum', 'warp insts')]
ik = hdr.index('Kernel Name')
units = rows[1]
for r in rows[2:]:
    print('###', r[ik].split('(')[0])
    for key, name in want:
        if key in hdr:
            i = hdr.index(key)
            print('  %-16s %s %s' % (name, r[i], units[i]))
