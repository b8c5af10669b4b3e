# usage: gpu_full.sh TAG   — full GPU check: -m gpu tests, smoke, default bench, reference arm, ncu launch list of one eager step
TAG=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/gpu_tests_$TAG.log 2>&1; tail -3 gpurun_out/gpu_tests_$TAG.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke_$TAG.log 2>&1; tail -2 gpurun_out/smoke_$TAG.log
timeout 900 python bench.py 2>gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.json; tail -c 600 gpurun_out/bench_$TAG.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.json; tail -c 400 gpurun_out/bench_ref_$TAG.json
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --profile-step --no-cpu-baseline --no-ref-gpu --no-fwd > gpurun_out/launch_run_$TAG.log 2>&1; tail -2 gpurun_out/launch_run_$TAG.log
# ncu --set full of the kernels named in $2.. (regex on the kernel name), ten launches each, inside one profiled eager step
shift
for K in "$@"; do
  timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:$K -c 10 -f -o /tmp/full_${K}_$TAG python bench.py --profile-step --no-cpu-baseline --no-ref-gpu --no-fwd > gpurun_out/full_${K}_$TAG.log 2>&1; tail -1 gpurun_out/full_${K}_$TAG.log
  # the report stays on the box (gpurun brings back at most 64 MiB): only its JSON summary travels
  python tools/ncu_rep_summary.py /tmp/full_${K}_$TAG.ncu-rep gpurun_out/r02_ncu_full_${K}.json "ncu --set full --clock-control none, the ten launches of one eager training step (128 clips, bf16), final round-2 build" > /dev/null 2>&1; ls -la gpurun_out/r02_ncu_full_${K}.json
done
