# usage: gpu_ab2.sh TAG "ENV=.." ["ENV=.."...] — full -m gpu tests once, then the train bench (no CPU baseline) under each environment
TAG=$1; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu --tb=short > gpurun_out/gpu_tests_$TAG.log 2>&1; tail -3 gpurun_out/gpu_tests_$TAG.log
i=0
for E in "$@"; do
  i=$((i+1))
  env $E timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-fwd 2>gpurun_out/bench_${TAG}_$i.err | tee gpurun_out/bench_${TAG}_$i.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$E', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['per_kernel_ms_per_step'])"
done
