# usage: ncu_src.sh TAG SKIP COUNT — ncu --set full + source counters of COUNT tc4_gemm_kernel launches (after SKIP) of one eager
# training step; brings back the per-instruction source page (CSV) and the raw page; the report itself stays on the box
TAG=$1; SKIP=$2; CNT=$3
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:tc4_gemm_kernel --launch-skip $SKIP -c $CNT -f -o /tmp/src_$TAG python bench.py --profile-step --no-cpu-baseline --no-ref-gpu --no-fwd > gpurun_out/src_$TAG.log 2>&1; tail -1 gpurun_out/src_$TAG.log
ncu -i /tmp/src_$TAG.ncu-rep --page source --csv > gpurun_out/src_$TAG.csv 2>/dev/null
ncu -i /tmp/src_$TAG.ncu-rep --page raw --csv > gpurun_out/raw_$TAG.csv 2>/dev/null
ls -la gpurun_out/src_$TAG.csv gpurun_out/raw_$TAG.csv
