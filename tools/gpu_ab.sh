# usage: gpu_ab.sh TAG
TAG=$1
timeout 600 python -m pytest tests -q -m gpu --tb=short > gpurun_out/gpu_tests_$TAG.log 2>&1; tail -4 gpurun_out/gpu_tests_$TAG.log
DSG_SIDE_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>gpurun_out/bench_$TAG.err | tee gpurun_out/bench_$TAG.json | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['roofline']['per_kernel_ms_per_step'])"
timeout 600 python bench.py --steps 10 --warmup 3 --batch 128 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('side-stream on:', d['value'], d['ms_per_step'])"
