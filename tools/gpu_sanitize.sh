# usage: gpu_sanitize.sh TAG — compute-sanitizer memcheck over one case of each TMA / tcgen05 kernel family (summary -> gpurun_out/)
TAG=$1
mkdir -p gpurun_out
T=tests/test_kernels.py
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest -q -m gpu -x --tb=short -p no:cacheprovider \
  "$T::test_conv_gemm_fast_engines[cuda-fwd-64-64]" "$T::test_conv_gemm_fast_engines[cuda-bwd-64-64]" "$T::test_conv_gemm_fast_engines[cuda-bwd_same-96-256]" \
  "$T::test_conv_gemm_fast_engines[cuda-dx-24-64]" "$T::test_conv_gemm_fast_engines[cuda-ext_in-64-64]" "$T::test_conv_gemm_fast_engines[cuda-contract-128-128]" \
  "$T::test_conv_wgrad_tma_engine[cuda-act-64-64]" "$T::test_conv_wgrad_tma_engine[cuda-ext-128-352]" \
  "$T::test_ms_conv_tap_shifted_tma[1-widths0]" "$T::test_ms_conv_tap_shifted_tma[2-widths1]" "$T::test_ms_conv_wgrad_tma[1-widths0-26]" \
  "$T::test_conv_gemm_fused_adjacency_contraction[True-25-24-64-100]" > gpurun_out/sanitize_$TAG.log 2>&1
echo "sanitizer rc=$?" >> gpurun_out/sanitize_$TAG.log
grep -E "ERROR SUMMARY|passed|failed|sanitizer rc|error" gpurun_out/sanitize_$TAG.log | tail -8
