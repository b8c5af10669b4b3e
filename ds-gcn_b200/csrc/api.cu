// extern "C" entry points of libdsgcn_b200.so (see include/dsgcn_b200.h).
#include "dsg_common.h"
#include "conv_gemm.cuh"
#include "conv_gemm_tc.cuh"
#include "conv_gemm_tc2.cuh"
#include "conv_gemm_tc3.cuh"
#include "tc4_gemm.cuh"
#include "tc4_wgrad.cuh"
#include "tc4_tconv.cuh"
#include "tc4_twgrad.cuh"
#include "graph_agg.cuh"
#include "ms_temporal_tc.cuh"
#include "topology.cuh"
#include "ctr_topology.cuh"
#include "misc.cuh"
#include "ms_mix.cuh"
#include "head.cuh"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

static thread_local char g_err[512] = "";
static long long g_counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};

static int fail(const char* where, const char* msg) {
    snprintf(g_err, sizeof(g_err), "%s: %s", where, msg);
    return 1;
}
#define DSG_RET(where, expr) do { const char* e__ = (expr); return e__ ? fail(where, e__) : 0; } while (0)

static bool dtype_ok(int d) { return d == DSG_F32 || d == DSG_BF16; }

// DSG_DISABLE_TC=1 routes bf16 GEMMs to the CUDA-core engine (debugging aid; both engines are sm_100a device code)
static bool tc3_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSG_DISABLE_TC3"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static bool tc2_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSG_DISABLE_TC2"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}
static bool tc_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("DSG_DISABLE_TC"); v = (e && e[0] == '1') ? 0 : 1; }
    return v == 1;
}

extern "C" {

const char* dsg_last_error(void) { return g_err; }
long long dsg_debug_counter(int id) { return (id >= 0 && id < 8) ? g_counters[id] : -1; }
int dsg_abi_version(void) { return DSG_ABI_VERSION; }
int dsg_is_device_build(void) {
#ifdef DSG_EMU
    return 0;
#else
    return 1;
#endif
}

int dsg_conv_gemm(const dsg_conv_gemm_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype)) return fail("dsg_conv_gemm", "bad arguments");
    if (a->taps < 1 || a->t_div < 1 || a->Vin < 1 || a->K < 1) return fail("dsg_conv_gemm", "bad shape");
    if ((a->stat_sum == nullptr) != (a->stat_sq == nullptr)) return fail("dsg_conv_gemm", "stat_sum and stat_sq go together");
    if (!a->out_f32 && !a->adyn && dsg::conv_gemm_skinny_ok(*a)) {               // N <= 8: a stream over the input rows, no GEMM tile
        if (a->dtype == DSG_BF16) DSG_RET("dsg_conv_gemm", dsg::launch_conv_gemm_skinny<bf16>(*a, (dsg_stream_t)stream));
        DSG_RET("dsg_conv_gemm", dsg::launch_conv_gemm_skinny<float>(*a, (dsg_stream_t)stream));
    }
#ifndef DSG_EMU
    if (a->dtype == DSG_BF16 && tc_enabled()) {      // tcgen05 engine (sm_100a); shapes it does not take fall through
        bool handled = false;
        const char* e = dsg::tc4::launch_conv_gemm_tc4(*a, (dsg_stream_t)stream, &handled);                            // TMA-fed warp-specialised engine
        if (e) return fail("dsg_conv_gemm", e);
        if (handled) { ++g_counters[a->adyn ? 2 : 0]; return 0; }
        if (a->adyn) return fail("dsg_conv_gemm", "fused adjacency contraction: shape not taken (bf16, K <= 64, N <= 128, taps == 1, one source tensor)");
        if (a->out_f32) return fail("dsg_conv_gemm", "out_f32: shape not taken by the TMA-fed engine (needs K, N % 8 == 0, taps == 1, no addends / mask / statistics)");
        e = tc3_enabled() ? dsg::tc::launch_conv_gemm_tc3(*a, (dsg_stream_t)stream, &handled) : nullptr;               // persistent pipelined engine
        if (e) return fail("dsg_conv_gemm", e);
        if (handled) return 0;
        e = tc2_enabled() ? dsg::tc::launch_conv_gemm_tc2(*a, (dsg_stream_t)stream, &handled) : nullptr;               // row-per-thread engine
        if (e) return fail("dsg_conv_gemm", e);
        if (handled) return 0;
        e = dsg::tc::launch_conv_gemm_tc(*a, (dsg_stream_t)stream, &handled);
        if (e) return fail("dsg_conv_gemm", e);
        if (handled) return 0;
    }
#endif
    if (a->out_f32) return fail("dsg_conv_gemm", "out_f32 needs bf16 sources and the tcgen05 engine");
    if (a->adyn) return fail("dsg_conv_gemm", "the fused adjacency contraction needs bf16 sources and the tcgen05 engine");
    if (a->dtype == DSG_BF16) DSG_RET("dsg_conv_gemm", dsg::launch_conv_gemm<bf16>(*a, (dsg_stream_t)stream));
    DSG_RET("dsg_conv_gemm", dsg::launch_conv_gemm<float>(*a, (dsg_stream_t)stream));
}

long long dsg_conv_gemm_wpack_bytes(int K, int N) {
#ifdef DSG_EMU
    (void)K; (void)N;
    return 0;
#else
    if (K <= 0 || N <= 0) return 0;
    const long long a_ = dsg::tc::conv_wpack_bytes(K, N), b_ = dsg::tc4::tc4_wpack_bytes(K, N);
    return a_ > b_ ? a_ : b_;
#endif
}

int dsg_conv_wgrad(const dsg_conv_wgrad_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype)) return fail("dsg_conv_wgrad", "bad arguments");
    if (a->taps < 1 || a->t_div < 1 || a->Vin < 1 || a->K < 1 || a->N < 1) return fail("dsg_conv_wgrad", "bad shape");
#ifndef DSG_EMU
    if (a->dtype == DSG_BF16 && tc_enabled()) {
        bool handled = false;
        const char* e = dsg::tc4::launch_conv_wgrad_tc4(*a, (dsg_stream_t)stream, &handled);                           // TMA-fed engine (1x1 convolutions)
        if (e) return fail("dsg_conv_wgrad", e);
        if (handled) { ++g_counters[1]; return 0; }
        e = dsg::tc::launch_conv_wgrad_tc(*a, (dsg_stream_t)stream, &handled);
        if (e) return fail("dsg_conv_wgrad", e);
        if (handled) return 0;
    }
#endif
    if (a->dtype == DSG_BF16) DSG_RET("dsg_conv_wgrad", dsg::launch_conv_wgrad<bf16>(*a, (dsg_stream_t)stream));
    DSG_RET("dsg_conv_wgrad", dsg::launch_conv_wgrad<float>(*a, (dsg_stream_t)stream));
}

int dsg_bn_finalize(const dsg_bn_job* jobs, int njobs, void* stream) {
    if (njobs < 0 || (njobs > 0 && !jobs)) return fail("dsg_bn_finalize", "bad arguments");
    for (int j0 = 0; j0 < njobs; j0 += dsg::BNF_MAX_JOBS) {
        dsg::BnJobs pack;
        memset(&pack, 0, sizeof(pack));
        int nj = njobs - j0 < dsg::BNF_MAX_JOBS ? njobs - j0 : dsg::BNF_MAX_JOBS;
        int cmax = 1;
        for (int j = 0; j < nj; ++j) {
            pack.j[j] = jobs[j0 + j];
            if (pack.j[j].mode < 0 || pack.j[j].mode > 4) return fail("dsg_bn_finalize", "bad mode");
            if (pack.j[j].C > cmax) cmax = pack.j[j].C;
        }
        dsg_launch(dsg::bn_finalize_kernel, dim3((cmax + 127) / 128, nj), dim3(128), 0, (dsg_stream_t)stream, pack);
        const char* e = dsg_launch_error();
        if (e) return fail("dsg_bn_finalize", e);
    }
    return 0;
}

int dsg_tmean2(const void* x, int dtype, long long ld, int n_samples, int T, int V, int C, float* xm, void* xm_bf16, void* stream) {
    if (!dtype_ok(dtype) || T < 1) return fail("dsg_tmean", "bad arguments");
    if (n_samples <= 0) return 0;
    bf16* xb = reinterpret_cast<bf16*>(xm_bf16);
    dim3 grid((V * C + 255) / 256, n_samples);
    if (dtype == DSG_BF16 && C % 8 == 0 && ld % 8 == 0 && (uintptr_t)x % 16 == 0 && (uintptr_t)xm_bf16 % 16 == 0) {
        dsg_launch(dsg::tmean_vec_kernel, dim3((V * (C / 8) + 127) / 128, n_samples), dim3(128), 0, (dsg_stream_t)stream, (const bf16*)x, ld, T, V, C, xm, xb);
        DSG_RET("dsg_tmean", dsg_launch_error());
    }
    if (dtype == DSG_BF16) dsg_launch(dsg::tmean_kernel<bf16>, grid, dim3(256), 0, (dsg_stream_t)stream, (const bf16*)x, ld, T, V, C, xm, xb);
    else dsg_launch(dsg::tmean_kernel<float>, grid, dim3(256), 0, (dsg_stream_t)stream, (const float*)x, ld, T, V, C, xm, xb);
    DSG_RET("dsg_tmean", dsg_launch_error());
}
int dsg_tmean(const void* x, int dtype, long long ld, int n_samples, int T, int V, int C, float* xm, void* stream) {
    return dsg_tmean2(x, dtype, ld, n_samples, T, V, C, xm, nullptr, stream);
}

int dsg_topology_fwd(const dsg_topology_args* a, void* stream) {
    if (!a || !dtype_ok(a->adyn_dtype)) return fail("dsg_topology_fwd", "bad arguments");
    DSG_RET("dsg_topology_fwd", dsg::launch_topology(*a, false, (dsg_stream_t)stream));
}
int dsg_topology_bwd(const dsg_topology_args* a, void* stream) {
    if (!a) return fail("dsg_topology_bwd", "bad arguments");
    DSG_RET("dsg_topology_bwd", dsg::launch_topology(*a, true, (dsg_stream_t)stream));
}

int dsg_head_ce_fwd(const float* pooled, const float* W, const float* b, const long long* label, int N, int C, int K,
                    float* logits, float* stats, void* stream) {
    DSG_RET("dsg_head_ce_fwd", dsg::launch_head_ce_fwd(pooled, W, b, label, N, C, K, logits, stats, (dsg_stream_t)stream));
}
int dsg_head_ce_bwd(const float* logits, const long long* label, const float* pooled, const float* W, const float* gscale,
                    int N, int C, int K, float* dlogits, float* dpooled, float* dW, float* db, void* stream) {
    DSG_RET("dsg_head_ce_bwd", dsg::launch_head_ce_bwd(logits, label, pooled, W, gscale, N, C, K, dlogits, dpooled, dW, db, (dsg_stream_t)stream));
}

int dsg_ctr_topology_fwd(const dsg_ctr_topology_args* a, void* stream) {
    if (!a || !dtype_ok(a->adyn_dtype)) return fail("dsg_ctr_topology_fwd", "bad arguments");
    DSG_RET("dsg_ctr_topology_fwd", dsg::launch_ctr_topology(*a, false, (dsg_stream_t)stream));
}
int dsg_ctr_topology_bwd(const dsg_ctr_topology_args* a, void* stream) {
    if (!a || !a->dadyn || !a->dH) return fail("dsg_ctr_topology_bwd", "bad arguments");
    DSG_RET("dsg_ctr_topology_bwd", dsg::launch_ctr_topology(*a, true, (dsg_stream_t)stream));
}

int dsg_graph_agg(const dsg_graph_agg_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype) || a->mode < 0 || a->mode > 3) return fail("dsg_graph_agg", "bad arguments");
    if ((a->stat_sum == nullptr) != (a->stat_sq == nullptr)) return fail("dsg_graph_agg", "stat_sum and stat_sq go together");
    if (a->dtype == DSG_BF16) DSG_RET("dsg_graph_agg", dsg::launch_agg<bf16>(*a, (dsg_stream_t)stream));
    DSG_RET("dsg_graph_agg", dsg::launch_agg<float>(*a, (dsg_stream_t)stream));
}

int dsg_graph_agg_dadj(const dsg_graph_agg_dadj_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype)) return fail("dsg_graph_agg_dadj", "bad arguments");
    if (a->dtype == DSG_BF16) DSG_RET("dsg_graph_agg_dadj", dsg::launch_dadj<bf16>(*a, (dsg_stream_t)stream));
    DSG_RET("dsg_graph_agg_dadj", dsg::launch_dadj<float>(*a, (dsg_stream_t)stream));
}

int dsg_ms_combine_fwd(const dsg_ms_combine_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype) || a->V > 32) return fail("dsg_ms_combine_fwd", "bad arguments");
    long long n_frames = (long long)a->n_samples * a->T_out;
    if (n_frames <= 0) return 0;
    {
        bool handled = false;                                          // 16-byte vector kernel (bf16, aligned)
        const char* e = dsg::launch_ms_mix_fwd(*a, (dsg_stream_t)stream, &handled);
        if (e) return fail("dsg_ms_combine_fwd", e);
        if (handled) return 0;
    }
    dim3 grid((unsigned)((n_frames + 7) / 8), (a->C + dsg::PW_CT - 1) / dsg::PW_CT);
    if (a->dtype == DSG_BF16) dsg_launch(dsg::ms_combine_fwd_kernel<bf16>, grid, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
    else dsg_launch(dsg::ms_combine_fwd_kernel<float>, grid, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
    DSG_RET("dsg_ms_combine_fwd", dsg_launch_error());
}

static int ms_combine_bwd_impl(const dsg_ms_combine_args* a, int parts, void* stream) {
    if (!a || !dtype_ok(a->dtype) || a->V > 32) return fail("dsg_ms_combine_bwd", "bad arguments");
    dsg_stream_t st = (dsg_stream_t)stream;
    long long n_out = (long long)a->n_samples * a->T_out, n_in = (long long)a->n_samples * a->T_in;
    if (n_out <= 0) return 0;
    {
        bool handled = false;                                          // 16-byte vector kernels (bf16, aligned, full-width d_o)
        const char* e = dsg::launch_ms_mix_bwd(*a, parts, st, &handled);
        if (e) return fail("dsg_ms_combine_bwd", e);
        if (handled) return 0;
    }
    if (parts & 1) {
        dim3 g1((unsigned)((n_out + 7) / 8), (a->C + dsg::PW_CT - 1) / dsg::PW_CT);
        if (a->dtype == DSG_BF16) dsg_launch(dsg::ms_combine_bwd_o_kernel<bf16>, g1, dim3(dsg::PW_THREADS), 0, st, *a);
        else dsg_launch(dsg::ms_combine_bwd_o_kernel<float>, g1, dim3(dsg::PW_THREADS), 0, st, *a);
        const char* e = dsg_launch_error();
        if (e) return fail("dsg_ms_combine_bwd", e);
    }
    if (parts & 2) {
        const int lo[2] = {a->max_lo, a->pass_lo}, hi[2] = {a->max_hi, a->pass_hi};
        for (int i = 0; i < 2; ++i) {
            if (hi[i] <= lo[i]) continue;
            dim3 g2((unsigned)((n_in + 7) / 8), (hi[i] - lo[i] + dsg::PW_CT - 1) / dsg::PW_CT);
            if (a->dtype == DSG_BF16) dsg_launch(dsg::ms_combine_bwd_e_kernel<bf16>, g2, dim3(dsg::PW_THREADS), 0, st, *a, lo[i], hi[i]);
            else dsg_launch(dsg::ms_combine_bwd_e_kernel<float>, g2, dim3(dsg::PW_THREADS), 0, st, *a, lo[i], hi[i]);
            const char* e = dsg_launch_error();
            if (e) return fail("dsg_ms_combine_bwd", e);
        }
    }
    return 0;
}

int dsg_ms_combine_bwd(const dsg_ms_combine_args* a, void* stream) { return ms_combine_bwd_impl(a, 3, stream); }
int dsg_ms_combine_bwd_part(const dsg_ms_combine_args* a, int parts, void* stream) {
    if (parts < 1 || parts > 3) return fail("dsg_ms_combine_bwd_part", "parts must be 1 (branch-output gradient), 2 (max / pass ranges) or 3");
    return ms_combine_bwd_impl(a, parts, stream);
}

int dsg_ms_temporal_supported(const dsg_ms_temporal_args* a) {
#ifdef DSG_EMU
    (void)a;
    return 0;
#else
    if (!a || !tc_enabled()) return 0;
    return (dsg::tc::ms_host_geom(*a, a->V + a->has_ext).ok && dsg::tc::ms_args_ok(*a)) ? 1 : 0;
#endif
}

long long dsg_ms_temporal_wpack_bytes(const dsg_ms_temporal_args* a) {
#ifdef DSG_EMU
    (void)a;
    return 0;
#else
    return a ? dsg::tc::ms_wpack_bytes(*a) : 0;
#endif
}

int dsg_ms_temporal_fwd(const dsg_ms_temporal_args* a, void* stream) {
#ifdef DSG_EMU
    (void)a; (void)stream;
    return fail("dsg_ms_temporal_fwd", "tcgen05 kernel: not available in the host-side simulator build");
#else
    if (!a) return fail("dsg_ms_temporal_fwd", "bad arguments");
    DSG_RET("dsg_ms_temporal_fwd", dsg::tc::launch_ms_temporal_fwd(*a, (dsg_stream_t)stream));
#endif
}

int dsg_ms_temporal_bwd_data(const dsg_ms_temporal_args* a, void* stream) {
#ifdef DSG_EMU
    (void)a; (void)stream;
    return fail("dsg_ms_temporal_bwd_data", "tcgen05 kernel: not available in the host-side simulator build");
#else
    if (!a) return fail("dsg_ms_temporal_bwd_data", "bad arguments");
    DSG_RET("dsg_ms_temporal_bwd_data", dsg::tc::launch_ms_temporal_bwd_data(*a, (dsg_stream_t)stream));
#endif
}

int dsg_ms_temporal_bwd_weight(const dsg_ms_temporal_args* a, void* stream) {
#ifdef DSG_EMU
    (void)a; (void)stream;
    return fail("dsg_ms_temporal_bwd_weight", "tcgen05 kernel: not available in the host-side simulator build");
#else
    if (!a) return fail("dsg_ms_temporal_bwd_weight", "bad arguments");
    DSG_RET("dsg_ms_temporal_bwd_weight", dsg::tc::launch_ms_temporal_bwd_weight(*a, (dsg_stream_t)stream));
#endif
}

long long dsg_ms_conv_wpack_bytes(const dsg_ms_conv_args* a) {
#ifdef DSG_EMU
    (void)a;
    return 0;
#else
    return a ? 2 * dsg::tc4::tc4_tconv_wpack_bytes(*a) : 0;          // two parity planes for a strided data gradient
#endif
}

int dsg_ms_conv(const dsg_ms_conv_args* a, int* handled, void* stream) {
    if (!a || !handled) return fail("dsg_ms_conv", "bad arguments");
    *handled = 0;
#ifdef DSG_EMU
    (void)stream;
    return 0;                                                         // tcgen05 + TMA kernel: the simulator build always declines
#else
    if (!tc_enabled()) return 0;
    bool h = false;
    const char* e = dsg::tc4::launch_ms_conv_tc4(*a, (dsg_stream_t)stream, &h);
    if (e) return fail("dsg_ms_conv", e);
    if (h) { *handled = 1; ++g_counters[3]; }
    return 0;
#endif
}

int dsg_ms_conv_wgrad(const dsg_ms_conv_args* a, int* handled, void* stream) {
    if (!a || !handled) return fail("dsg_ms_conv_wgrad", "bad arguments");
    *handled = 0;
#ifdef DSG_EMU
    (void)stream;
    return 0;                                                         // tcgen05 + TMA kernel: the simulator build always declines
#else
    if (!tc_enabled()) return 0;
    bool h = false;
    const char* e = dsg::tc4::launch_ms_conv_wgrad_tc4(*a, (dsg_stream_t)stream, &h);
    if (e) return fail("dsg_ms_conv_wgrad", e);
    if (h) { *handled = 1; ++g_counters[4]; }
    return 0;
#endif
}

int dsg_pointwise(const dsg_pointwise_args* a, void* stream) {
    if (!a || !dtype_ok(a->dtype) || !dtype_ok(a->out_dtype)) return fail("dsg_pointwise", "bad arguments");
    if ((a->stat_sum == nullptr) != (a->stat_sq == nullptr)) return fail("dsg_pointwise", "stat_sum and stat_sq go together");
    if (a->rows <= 0 || a->C <= 0) return 0;
    if (dsg::pointwise_vec_ok(*a)) {
        dim3 gv((unsigned)((a->rows + dsg::PV_ROWS - 1) / dsg::PV_ROWS), (a->C + dsg::PW_CT - 1) / dsg::PW_CT);
        const bool simple = a->src.x2 == nullptr && !(a->has_mask && a->mask.x2 != nullptr);
        if (simple) dsg_launch(dsg::pointwise_vec_kernel<true>, gv, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
        else dsg_launch(dsg::pointwise_vec_kernel<false>, gv, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
        DSG_RET("dsg_pointwise", dsg_launch_error());
    }
    dim3 grid((unsigned)((a->rows + dsg::PW_ROWS - 1) / dsg::PW_ROWS), (a->C + dsg::PW_CT - 1) / dsg::PW_CT);
    if (a->dtype == DSG_BF16) dsg_launch(dsg::pointwise_kernel<bf16>, grid, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
    else dsg_launch(dsg::pointwise_kernel<float>, grid, dim3(dsg::PW_THREADS), 0, (dsg_stream_t)stream, *a);
    DSG_RET("dsg_pointwise", dsg_launch_error());
}

int dsg_sgd_step(float* p, const float* grad, float* buf, long long n, float lr, float momentum, float wd,
                 int nesterov, float grad_scale, void* stream) {
    if (n <= 0) return 0;
    dsg_launch(dsg::sgd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (dsg_stream_t)stream, p, grad, buf, n, lr, (const float*)nullptr,
               momentum, wd, nesterov, grad_scale);
    DSG_RET("dsg_sgd_step", dsg_launch_error());
}

int dsg_sgd_step_dev(float* p, const float* grad, float* buf, long long n, const float* lr_dev, float momentum, float wd,
                     int nesterov, float grad_scale, void* stream) {
    if (n <= 0) return 0;
    if (!lr_dev) return fail("dsg_sgd_step_dev", "lr_dev is null");
    dsg_launch(dsg::sgd_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, (dsg_stream_t)stream, p, grad, buf, n, 0.f, lr_dev, momentum, wd,
               nesterov, grad_scale);
    DSG_RET("dsg_sgd_step_dev", dsg_launch_error());
}

}  // extern "C"
