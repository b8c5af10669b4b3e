# usage: gpu_knobs.sh — micro-benchmark of the tc4 GEMM engine under the experiment knobs (DSG_TC4_*); DBG results are timing-only
mkdir -p gpurun_out
CASES=${CASES:-fwd64,pre64,act64,ext64,bwd64,dx64,fwd128,bwd128,fwd256,act256,bwd256,cx64,cx256}
run() { echo "== $*"; env "$@" python tools/bench_gemm.py --cases $CASES --reps 20 2>&1 | grep -v Warning; }
{
run DSG_TC4_EPIALT=0
run DSG_TC4_EPIALT=1
run DSG_TC4_EPIALT=1 DSG_TC4_OB=4
run DSG_TC4_EPIALT=1 DSG_TC4_DBG=1
} > gpurun_out/knobs.log 2>&1
timeout 900 python -m pytest tests/test_kernels.py tests/test_units.py -q -m gpu -x --tb=short > gpurun_out/knobs_tests.log 2>&1; tail -5 gpurun_out/knobs_tests.log
