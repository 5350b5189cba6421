import csv, sys, subprocess, collections, re
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','smsp__inst_executed.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','l1tex__t_sector_hit_rate.pct','sm__cycles_elapsed.max','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__inst_executed_pipe_fma.sum','smsp__thread_inst_executed_per_inst_executed.ratio','dram__bytes_read.sum','dram__bytes_write.sum','sm__inst_executed_pipe_lsu.sum','smsp__inst_executed_pipe_fmaheavy.sum','smsp__inst_executed_pipe_fmalite.sum', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_tensor_op_hmma.sum','smsp__inst_executed_pipe_tensor.sum','launch__grid_size','launch__block_size','launch__shared_mem_per_block_dynamic']
for r in rows[2:]:
    for w in want:
        if w in hdr: print(w,'=',r[hdr.index(w)])
    for i,h in enumerate(hdr):
        if 'issue_stalled' in h and h.endswith('per_issue_active.ratio'):
            try:
                if float(r[i])>0.05: print('  ',h.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''), r[i])
            except: pass
