"""Print a compact summary of an .ncu-rep (one line block per captured launch)."""
import csv, subprocess, sys, io
KEYS = ['gpu__time_duration.sum','sm__cycles_elapsed.avg.per_second','launch__grid_size','launch__block_size','launch__registers_per_thread',
 'launch__waves_per_multiprocessor','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps',
 'sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
 'dram__bytes_read.sum','dram__bytes_write.sum','dram__bytes_read.sum.per_second','dram__bytes_write.sum.per_second',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_bytes.sum','l1tex__t_bytes.sum',
 'memory_l1_wavefronts_shared','memory_l1_wavefronts_shared_ideal','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum','l1tex__t_requests_pipe_lsu_mem_global_op_st.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
 'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_membar_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio']
rep = sys.argv[1]
out = subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, units = rows[0], rows[1]
for row in rows[2:]:
    d = dict(zip(hdr,row)); u = dict(zip(hdr,units))
    print('== kernel', d.get('Kernel Name'), 'id', d.get('ID'))
    for k in KEYS:
        if k in d: print(f"  {k:88s} {d[k]:>18s} {u[k]}")
