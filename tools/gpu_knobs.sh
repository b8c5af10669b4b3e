# usage: gpu_knobs.sh — micro-benchmark (tools/bench_gemm.py) of the tc4 GEMM engine under its experiment knobs, one table per setting
# in gpurun_out/knobs.log.  Knobs (environment, read once per process): DSG_TC4_S / DSG_TC4_OB (ring depths), DSG_TC4_DEFER (deferred
# drain retirement), DSG_TC4_XFMAP (prologue mapping), DSG_TC4_TAILTMA (staged tails), DSG_TC4_GRAMONES ([tile | ones] statistics);
# DSG_TC4_DBG (timing-only ablation) needs a library built with -DDSG_TC4_ABLATION.
mkdir -p gpurun_out
CASES=${CASES:-fwd64,pre64,act64,ext64,bwd64,dx64,fwd128,bwd128,fwd256,act256,bwd256,cx64,cx256}
run() { echo "== $*"; env "$@" python tools/bench_gemm.py --cases $CASES --reps 20 2>&1 | grep -v Warning; }
{
run DSG_TC4_GRAMONES=0 DSG_TC4_TAILTMA=0 DSG_TC4_XFMAP=0
run DSG_TC4_GRAMONES=1 DSG_TC4_TAILTMA=0 DSG_TC4_XFMAP=0
run DSG_TC4_GRAMONES=1 DSG_TC4_TAILTMA=1 DSG_TC4_XFMAP=0
run DSG_TC4_GRAMONES=1 DSG_TC4_TAILTMA=1 DSG_TC4_XFMAP=1
run DSG_TC4_OB=4
run DSG_TC4_S=3
} > gpurun_out/knobs.log 2>&1
