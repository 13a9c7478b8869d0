import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','dram__bytes_read.sum','dram__bytes_write.sum','sm__warps_active.avg.pct_of_peak_sustained_active','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__occupancy_limit_registers','launch__occupancy_limit_shared','sm__throughput.avg.pct_of_peak_sustained_elapsed','smsp__pcsamp_warps_issue_stalled','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','l1tex__t_sector_hit_rate','sm__inst_executed_pipe_fp64','smsp__issue_active.avg.pct','lts__t_sector_hit_rate','sm__inst_executed_pipe_xu','smsp__inst_executed_pipe_lsu','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__warp_issue_stalled','launch__block_size','launch__waves']
idx=[i for i,h in enumerate(hdr) if any(w in h for w in want) and 'not_issued' not in h and '.max' not in h and '.min' not in h]
for r in rows[2:]:
    for i in idx:
        if r[i] not in ('0',''): print(hdr[i],'=',r[i][:60], rows[1][i])
    print('----')
