"""Condense an .ncu-rep (one kernel launch, --set full) into the text summaries kept in this directory.
usage: python profiles/ncu_summary.py report.ncu-rep "header line" > profiles/rNN_xxx_ncu.txt"""
import csv, io, subprocess, sys

KEYS = """Kernel Name
dram__bytes_read.sum
dram__bytes_write.sum
gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed
gpu__time_duration.sum
l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum
l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum
smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum
l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum
l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum
l1tex__throughput.avg.pct_of_peak_sustained_elapsed
launch__block_size
launch__grid_size
launch__occupancy_limit_registers
launch__occupancy_limit_shared_mem
launch__registers_per_thread
launch__shared_mem_per_block_dynamic
lts__throughput.avg.pct_of_peak_sustained_elapsed
sm__cycles_elapsed.avg.per_second
sm__inst_executed.sum
sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active
sm__inst_issued.avg.pct_of_peak_sustained_active
sm__throughput.avg.pct_of_peak_sustained_elapsed
sm__warps_active.avg.pct_of_peak_sustained_active
smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio
smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio
smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio
smsp__average_warps_issue_stalled_wait_per_issue_active.ratio""".split("\n")

rep, header = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
col = {h: i for i, h in enumerate(hdr)}
print(header)
for k in KEYS:
    if k in col:
        print("%-96s%-17s%s" % (k, units[col[k]], vals[col[k]]))
