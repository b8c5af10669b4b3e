import csv,sys,subprocess
rep=sys.argv[1]
out=subprocess.run(['ncu','-i',rep,'--page','raw','--csv'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[0]
cols={'dur':'gpu__time_duration.sum','rd':'dram__bytes_read.sum','wr':'dram__bytes_write.sum','warps':'sm__warps_active.avg.pct_of_peak_sustained_active','grid':'launch__grid_size','smem':'launch__shared_mem_per_block_dynamic','smthr':'sm__throughput.avg.pct_of_peak_sustained_elapsed','inst':'smsp__inst_executed.sum','regs':'launch__registers_per_thread','l2':'lts__throughput.avg.pct_of_peak_sustained_elapsed','l1':'l1tex__throughput.avg.pct_of_peak_sustained_elapsed','occ_smem':'launch__occupancy_limit_shared_mem','occ_reg':'launch__occupancy_limit_registers','ach_occ':'sm__warps_active.avg.per_cycle_active','waves':'launch__waves_per_multiprocessor'}
idx={k:hdr.index(v) for k,v in cols.items() if v in hdr}
for r in rows[2:]:
    print(' '.join(f"{k}={r[i][:9]}" for k,i in idx.items()))
