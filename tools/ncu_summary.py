"""Print the handful of ncu metrics we care about from a .ncu-rep (read here, no GPU)."""
import csv, subprocess, sys
WANT = ['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__registers_per_thread',
 'sm__warps_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed',
 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sectors.sum','lts__t_sector_hit_rate.pct',
 'smsp__inst_executed.sum','smsp__thread_inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active',
 'launch__grid_size','launch__block_size','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem',
 'launch__occupancy_limit_warps','sm__maximum_warps_per_active_cycle_pct','smsp__thread_inst_executed_per_inst_executed.ratio',
 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum',
 'smsp__average_warp_latency_issue_stalled_barrier.ratio','smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio','smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_wait_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_membar_per_issue_active.ratio','smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio','smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_drain_per_issue_active.ratio','smsp__average_warps_issue_stalled_imc_miss_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio','smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio',
 'smsp__average_warps_issue_stalled_tex_throttle_per_issue_active.ratio','smsp__average_warps_issue_stalled_selected_per_issue_active.ratio',
 'sm__cycles_active.avg','sm__cycles_elapsed.max','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
 'l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum','lts__t_sectors_op_red.sum','lts__t_sectors_op_read.sum','lts__t_sectors_op_write.sum',
 'sm__ctas_launched.sum','launch__waves_per_multiprocessor']
out = subprocess.run(['ncu','-i',sys.argv[1],'--page','raw','--csv'],capture_output=True,text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
for vals in rows[2:]:
    print('== kernel:', vals[hdr.index('Kernel Name')][:90])
    for h,u,v in zip(hdr,units,vals):
        if h in WANT: print(f'  {h:95s} {v} {u}')
