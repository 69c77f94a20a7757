import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]; units=rows[1]; vals=rows[2]
want=['Kernel Name','gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_executed_pipe_alu.sum.pct','sm__inst_executed_pipe_fma.sum.pct','sm__pipe_alu_cycles_active.avg.pct','sm__pipe_fma_cycles_active.avg.pct','sm__pipe_fmaheavy','sm__inst_executed_pipe_lsu.sum.pct','sm__inst_executed_pipe_uniform','smsp__issue_active.avg.pct','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','sm__throughput.avg.pct','smsp__thread_inst_executed_per_inst_executed.ratio','smsp__average_warp','smsp__warps_eligible','smsp__inst_executed_pipe_alu.sum ','smsp__inst_executed_pipe_fma.sum ','lts__t_bytes.sum ','smsp__cycles_active.avg ']
for i,h in enumerate(hdr):
    if any(h.startswith(w.strip()) for w in want) and 'per_second' not in h: print(h, '|', units[i], '|', vals[i])
print('--- stalls')
st=[(float(vals[i]),h) for i,h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio') or (h.startswith('smsp__average_warp_latency_issue_stalled') )]
for v,h in sorted(st,reverse=True)[:12]: print('%8.3f'%v,h)
