#!/bin/bash
# key metrics of an ncu report: tools/ncu_key.sh report.ncu-rep
ncu -i "$1" --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin))
hd=rows[0]; un=rows[1]; vals=rows[2]
want=['gpu__time_duration.sum','smsp__inst_executed.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_issued.avg.pct_of_peak_sustained_active','dram__bytes_read.sum','dram__bytes_write.sum','dram__cycles_active.avg.pct_of_peak_sustained_elapsed','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.pct_of_peak_sustained_active','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct','smsp__average_warp_latency_per_inst_issued.ratio','smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio','smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio','smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio']
for w in want:
    for h,u,v in zip(hd,un,vals):
        if h==w: print(f'{w:85s} {u:>10s} {v}')
"
