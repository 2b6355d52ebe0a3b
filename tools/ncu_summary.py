"""Summary CSV of one `ncu --set full` report for profiles/: python tools/ncu_summary.py REPORT.ncu-rep "comment line" > profiles/NAME.csv"""
import csv
import io
import subprocess
import sys

WANT = """gpu__time_duration.sum dram__bytes_read.sum dram__bytes_write.sum gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
dram__cycles_active.avg.pct_of_peak_sustained_elapsed l1tex__t_sector_hit_rate.pct l1tex__throughput.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed lts__t_sector_hit_rate.pct lts__throughput.avg.pct_of_peak_sustained_elapsed
launch__block_size launch__grid_size launch__registers_per_thread launch__shared_mem_per_block_static launch__occupancy_limit_registers
sm__throughput.avg.pct_of_peak_sustained_elapsed sm__warps_active.avg.pct_of_peak_sustained_active sm__inst_issued.avg.pct_of_peak_sustained_active
smsp__issue_active.avg.pct_of_peak_sustained_active smsp__inst_executed.sum smsp__thread_inst_executed_per_inst_executed.ratio
smsp__average_warp_latency_per_inst_issued.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_wait_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio
smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio
sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active
smsp__inst_executed_op_global_red.sum l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
sm__cycles_elapsed.avg""".split()


def main():
    report = sys.argv[1]
    out = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hd, un, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hd)}
    for c in sys.argv[2:]:
        print("# " + c)
    print(f"Kernel Name,,{vals[col['Kernel Name']]}")
    for w in WANT:
        if w in col:
            print(f"{w},{un[col[w]]},{vals[col[w]]}")


if __name__ == "__main__":
    main()
