timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc4_gemm_kernel --launch-skip 60 -c 12 -f -o /tmp/full_tails python bench.py --profile-step --no-cpu-baseline --no-ref-gpu --no-fwd > gpurun_out/full_tails.log 2>&1; tail -1 gpurun_out/full_tails.log
python tools/ncu_rep_summary.py /tmp/full_tails.ncu-rep gpurun_out/ncu_tails.json "tc4_gemm launches 60..71 of one eager training step (backward)" > /dev/null 2>&1
ncu -i /tmp/full_tails.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]
want=['Kernel Name','gpu__time_duration.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','l1tex__throughput.avg.pct_of_peak_sustained_elapsed','lts__throughput.avg.pct_of_peak_sustained_elapsed','sm__inst_executed_pipe_lsu.sum','smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct','smsp__warp_issue_stalled_barrier_per_warp_active.pct','smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum','launch__grid_size']
idx=[h.index(w) for w in want if w in h]
for r in rows[2:]: print([r[i][:60] for i in idx])
" > gpurun_out/ncu_tails_raw.txt
