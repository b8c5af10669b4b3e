import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','launch__shared_mem_per_block_dynamic','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','sm__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','smsp__inst_executed.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','smsp__cycles_active.avg','sm__inst_executed_pipe_lsu.sum','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__average_warp_latency_issue_stalled_long_scoreboard.pct' ]
idx=[(w,hdr.index(w)) for w in want if w in hdr]
units=rows[1]
for r in rows[2:]:
    print('---')
    for w,i in idx: print(f'  {w} = {r[i]} {units[i]}')
