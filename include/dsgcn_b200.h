/* dsgcn_b200.h — C ABI of the B200-native DS-GCN backbone kernels (libdsgcn_b200.so).
 *
 * The reference (davelailai/DS-GCN, a PYSKL fork) has no FFI layer: its hot path is torch.nn
 * modules.  This ABI is the new layer *below* those modules.  Each entry point is
 * one fused device operation; the comment above it names the reference code it replaces
 * (paths relative to the reference root).  The Python host side (ds-gcn_b200/) mirrors the
 * reference's module API (unit_gcn, dgphgcn1, unit_tcn, mstcn, dgmstcn, DGBlock, DGSTGCN) and
 * reaches these entry points through ctypes; INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - plain C types only: device pointers, sizes, a dtype enum, a CUDA stream passed as void*.
 *  - the caller owns every buffer (inputs, outputs, statistics, workspaces); nothing is
 *    allocated, freed or cached by the library; no global mutable state; calls are
 *    stream-ordered and never synchronise.
 *  - every function returns 0 on success, non-zero on error; dsg_last_error() returns a
 *    thread-local message for the last failure on the calling thread.
 *  - activations are channels-last: logical [n, C, t, v] (the reference's NCHW) is stored as
 *    rows r = (n*T + t)*V + v of C contiguous channels, row pitch `ld` elements (so a channel
 *    slice of a wider buffer is addressable).  dtype is DSG_F32 or DSG_BF16 for activations;
 *    parameters, statistics and their gradients are always fp32 (statistics sums are fp64).
 */
#ifndef DSGCN_B200_H
#define DSGCN_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSG_ABI_VERSION 1

typedef enum { DSG_F32 = 0, DSG_BF16 = 1 } dsg_dtype;

/* A logical activation defined on the fly (fused prologue):
 *   value(r,c) = f( a1[c]*x1[r*ld1+c] + b1[c] + a2[c]*x2[r*ld2+c] + b2[c] ),  f = relu | identity.
 * NULL a => 1, NULL b => 0, NULL x2 => term dropped.  Covers BatchNorm-apply(+ReLU) on a raw
 * convolution output, relu(bn(z) + residual) (gcn.py:2365, dgstgcn.py:64-65) and the BatchNorm
 * backward combination dy = ca*e + cb*y + cc. */
typedef struct dsg_act_src {
    const void* x1;
    const void* x2;
    const float* a1;
    const float* b1;
    const float* a2;
    const float* b2;
    long long ld1;
    long long ld2;
    int relu;
    int pad_;
} dsg_act_src;

/* ---- dsg_conv_gemm --------------------------------------------------------------------------
 * out[f', j, n] = bias[n] + sum_{tap,k} src[frame(f',tap), j, k] * W[n*ws_n + k*ws_k + tap*ws_tap]
 * A (taps x 1) temporal convolution over channels-last frames, i.e. nn.Conv2d(k x 1) with stride,
 * dilation and zero padding in T; taps == 1 is the 1x1 convolution.  Replaces the F.conv2d calls of
 * gcn.py:62,2165-2169,2187-2195,2203,2211 and tcn.py:21-27,383-402 and, with the transposed frame
 * map / weight strides, their data gradients.
 *   frame map : out frame f' = n*T_out + t'.  num = t'*t_mul + tap*tap_step + tap_off; the tap
 *               contributes iff num % t_div == 0 and 0 <= num/t_div < T_in; source frame n*T_in + num/t_div.
 *   rows      : Vin rows per source frame; ext_in appends the joint-mean row (tcn.py:409) so the
 *               output has Vin+1 rows per frame; contract_ext folds the last row back
 *               (out[j] = acc[j] + acc[Vin-1]/(Vin-1), rows Vin-1 per frame: gradient of that mean).
 *   epilogue  : + bias; + add + add2 (same shape as out, dtype); + bcast[n, j, c]*bcast_scale (fp32, per
 *               sample: gradient of the temporal mean, gcn.py:2246); * [mask > 0]; statistics
 *               stat_sum[c] += v, stat_sq[c] += v * partner(r,c) (partner NULL => v: BatchNorm batch
 *               statistics; partner = saved raw output: BatchNorm backward sums). */
typedef struct dsg_conv_gemm_args {
    dsg_act_src src;
    int dtype;            /* dsg_dtype of src, out, add, mask, partner */
    int K, N;
    const float* W;
    long long ws_n, ws_k, ws_tap;
    const float* bias;
    int taps, tap_step, tap_off, t_mul, t_div;
    int n_samples, T_in, T_out, Vin;
    int ext_in, contract_ext;
    void* out;
    long long ld_out;
    const void* add;
    long long ld_add;
    const void* add2;     /* second addend, same conventions as add */
    long long ld_add2;
    const float* bcast;   /* [n_samples, Vout, N] fp32 */
    float bcast_scale;
    int has_mask;
    dsg_act_src mask;     /* evaluated at output rows */
    double* stat_sum;
    double* stat_sq;
    const void* partner;
    long long ld_partner;
    void* wpack;          /* optional caller-owned workspace of dsg_conv_gemm_wpack_bytes(K, N) bytes (256-byte aligned): the
                             call first rewrites it with bf16 UMMA-ready weight tiles, which every CTA of the tcgen05 engine
                             then pulls with one bulk copy instead of converting fp32 weights itself; NULL = convert in-kernel */
    int out_f32;          /* 1 (bf16 sources only, no addends / mask / statistics): `out` is fp32 — the accumulator is stored
                             unrounded (the topology-feature convolutions conv1/conv2/conv1_se, gcn.py:2248-2259, feed tanh / softmax) */
    int pad2_;
    /* Fused spatial graph convolution (north-star kernel (a), gcn.py:2350-2363): when `adyn` is set the source rows are first
     * contracted frame by frame with the per-sample, per-channel adjacency, y[n,t,w,k] = sum_u src[n,t,u,k] * adyn[n,u,w,k], INSIDE
     * the kernel (operand producer of the tensor core), and out = bias + y W^T.  bf16, K <= 64, N <= 128, taps == 1.  y_out
     * (optional, [rows, ld_y]) also stores y — the training forward keeps it for the weight gradient; inference never writes it. */
    const void* adyn;     /* [n_samples, Vin, Vin, K] bf16 */
    void* y_out;
    long long ld_y;
} dsg_conv_gemm_args;
int dsg_conv_gemm(const dsg_conv_gemm_args* a, void* stream);
long long dsg_conv_gemm_wpack_bytes(int K, int N);

/* ---- dsg_conv_wgrad -------------------------------------------------------------------------
 * dW[n*ws_n + k*ws_k + tap*ws_tap] += sum_{f',j} A[frame(f',tap), j, k] * B[f', j, n];  db[n] += sum B.
 * Weight/bias gradient of dsg_conv_gemm (autograd of the conv calls listed above).  A uses the same
 * frame map and ext_in as the forward; B is defined on output rows.  dW/db are accumulated with
 * atomics and must be zero-initialised by the caller. */
typedef struct dsg_conv_wgrad_args {
    dsg_act_src A;
    dsg_act_src B;
    int dtype;
    int K, N;
    float* dW;
    long long ws_n, ws_k, ws_tap;
    float* db;
    int taps, tap_step, tap_off, t_mul, t_div;
    int n_samples, T_in, T_out, Vin;
    int ext_in;
} dsg_conv_wgrad_args;
int dsg_conv_wgrad(const dsg_conv_wgrad_args* a, void* stream);

/* ---- dsg_bn_finalize ------------------------------------------------------------------------
 * Turns accumulated statistics into the per-channel coefficients the fused prologues consume.
 * Replaces the statistics half of nn.BatchNorm2d (F.batch_norm) for every BN on the path.
 *  forward  (mode 0): mean = sum/count, var = sq/count - mean^2 (biased), a = gamma*rsqrt(var+eps),
 *                     b = beta - mean*a; running stats updated with momentum and unbiased variance;
 *                     save_mean/save_invstd written.  eval (mode 1): a,b from the running stats.
 *  backward (mode 2): s1 = sum e, s2r = sum e*y  ->  dgamma = invstd*(s2r - mean*s1), dbeta = s1,
 *                     ca = gamma*invstd, cb = -gamma*invstd^2*dgamma/count,
 *                     cc = -ca*s1/count - cb*mean   (dy = ca*e + cb*y + cc).
 *  backward-eval (mode 3): BN used running stats: ca = gamma*invstd, cb = cc = 0.
 *  identity (mode 4): a = 1, b = 0 (channels that have no BN, e.g. the '1x1' branch tcn.py:383). */
typedef struct dsg_bn_job {
    int mode, C;
    const double* sum;
    const double* sq;
    double count;
    const float* gamma;
    const float* beta;
    float* running_mean;
    float* running_var;
    float* save_mean;
    float* save_invstd;
    float* a;       /* forward: a ; backward: ca */
    float* b;       /* forward: b ; backward: cb */
    float* c;       /* backward: cc */
    float* dgamma;
    float* dbeta;
    float momentum, eps;
} dsg_bn_job;
int dsg_bn_finalize(const dsg_bn_job* jobs, int njobs, void* stream);

/* ---- dsg_tmean ------------------------------------------------------------------------------
 * xm[n, v, c] = mean_t x[n, t, v, c]  (fp32 out).  gcn.py:2246 `tmp_x.mean(dim=-2)`. */
int dsg_tmean(const void* x, int dtype, long long ld, int n_samples, int T, int V, int C, float* xm, void* stream);
/* the same with a second, bf16 copy of the result (operand of the tensor-core topology-feature GEMM); xm_bf16 may be NULL */
int dsg_tmean2(const void* x, int dtype, long long ld, int n_samples, int T, int V, int C, float* xm, void* xm_bf16, void* stream);

/* ---- dsg_topology_fwd / dsg_topology_bwd ----------------------------------------------------
 * The per-sample dynamic semantic adjacency of dgphgcn1 (gcn.py:2239-2337, north-star flags).
 * H[n, v, 9R] = xm @ [conv1; conv2; conv1_se]^T + bias (computed with dsg_conv_gemm) holds
 * x1n (cols 0..2R), x2n (2R..4R) and the 5 node-type variants of the semantic feature
 * (col 4R + c*5 + type).  Output adyn[n, u, w, k*R + c] = A[k,u,w] + alpha_k*tanh(.) + beta_k*softmax_u(.)
 * (dtype `adyn_dtype`), plus the column-softmax S[n, k, u, w] (fp32) kept for the backward pass.
 * Backward consumes dAdyn[n, u, w, 3R] (fp32) and produces dH (same layout as H, untouched entries
 * zero), and atomically accumulates dA[3,V,V], dalpha[3], dbeta[3], dWe[15R,R], dbe[15R]. */
typedef struct dsg_topology_args {
    const float* H;
    long long ld_h;
    int n_samples, V, R;
    const int* node_type;     /* [V] device */
    const int* edge_type;     /* [V*V] device */
    const float* A;           /* [3,V,V] */
    const float* alpha;       /* [3] */
    const float* beta;        /* [3] */
    const float* We;          /* [15R, R] edge_linears.weight */
    const float* be;          /* [15R] */
    void* adyn;
    int adyn_dtype;
    float* S;                 /* [n,3,V,V] */
    /* backward only */
    const float* dadyn;       /* [n, V, V, 3R] */
    float* dH;
    float* dA;
    float* dalpha;
    float* dbeta;
    float* dWe;
    float* dbe;
    void* dH_bf16;            /* optional bf16 copy of dH (same pitch): operand of the tensor-core GEMMs that consume it */
    /* flag variants of the same unit (gcn.py:1445-1584 dggcn; dghgcn :1586 / dgphgcn :1808 / dgphgcn1 with the attention flags off):
     * variant 1 = plain DG-GCN topology: H[n, v, 6R] = [conv1 (3R) | conv2 (3R)], every subset uses tanh(x1[u]-x2[w]); node_type,
     * edge_type, We, be, dWe, dbe are unused (may be NULL).  subset_wise = 0: alpha[0] / beta[0] scale every subset (and receive
     * the whole gradient), gcn.py:1543-1546. */
    int variant;
    int subset_wise;
} dsg_topology_args;
int dsg_topology_fwd(const dsg_topology_args* a, void* stream);
int dsg_topology_bwd(const dsg_topology_args* a, void* stream);

/* ---- dsg_ctr_topology_fwd / dsg_ctr_topology_bwd ---------------------------------------------
 * Channel-wise topology refinement of CTR-GCN: the three CTRGC modules of one unit_ctrgcn (gcn.py:650-659, :914-918).
 * H[n, v, 6R] = xm @ [conv1_0; conv1_1; conv1_2; conv2_0; conv2_1; conv2_2]^T + bias (dsg_conv_gemm on the temporal mean).
 * adyn[n, u, w, k*C + c] = alpha * (sum_r W4_k[c,r] * tanh(x1_k[r,u] - x2_k[r,w]) + b4_k[c]) + A[k,u,w]   (dtype adyn_dtype),
 * the operand of dsg_graph_agg mode 0 over conv3's output (einsum 'ncuv,nctu->nctv').  Backward consumes dAdyn[n,u,w,3C]
 * (fp32), writes dH (layout of H) and atomically accumulates dA[3,V,V], dalpha[1], dW4[3,C,R], db4[3,C]. */
typedef struct dsg_ctr_topology_args {
    const float* H;
    long long ld_h;
    int n_samples, V, R, C;
    const float* A;           /* [3,V,V] */
    const float* alpha;       /* [1] */
    const float* W4;          /* [3,C,R]  conv4 weights of convs[0..2] */
    const float* b4;          /* [3,C] */
    void* adyn;
    int adyn_dtype;
    /* backward only */
    const float* dadyn;       /* [n, V, V, 3C] */
    float* dH;
    void* dH_bf16;            /* optional bf16 copy of dH */
    float* dA;
    float* dalpha;
    float* dW4;
    float* db4;
} dsg_ctr_topology_args;
int dsg_ctr_topology_fwd(const dsg_ctr_topology_args* a, void* stream);
int dsg_ctr_topology_bwd(const dsg_ctr_topology_args* a, void* stream);

/* ---- dsg_graph_agg --------------------------------------------------------------------------
 * y[n,t,w,kc] = sum_u p[n,t,u,kc] * adj(n,kc,u,w)    — the adjacency contraction.
 *  mode 0 (dynamic): adj = adyn[n,u,w,kc]           gcn.py:2352  einsum('nkctv,nkcvw->nkctw')
 *  mode 1 (dynamic, transposed): adj = adyn[n,w,u,kc]  -> gradient w.r.t. p
 *  mode 2 (static, subset-summed): y[n,t,w,c] = sum_k sum_u p[n,t,u,k*C+c]*A[k,u,w]   gcn.py:88
 *  mode 3 (static transposed, subset-expanding): y[n,t,u,k*C+c] = sum_w p[n,t,w,c]*A[k,u,w]
 * `p` is an activation source (fused BN+ReLU); the epilogue optionally masks and accumulates
 * BatchNorm-backward statistics exactly like dsg_conv_gemm. */
typedef struct dsg_graph_agg_args {
    dsg_act_src src;
    int dtype;
    int mode;
    int n_samples, T, V, KC;      /* KC = channels of the *dynamic* tensor / of y for mode 2 */
    int Ksub;                     /* static modes: number of subsets K */
    const void* adyn;             /* dynamic: [n,V,V,KC] (dtype) */
    const float* A;               /* static: [K,V,V] fp32 */
    void* out;
    long long ld_out;
    int has_mask;
    dsg_act_src mask;
    double* stat_sum;
    double* stat_sq;
    const void* partner;
    long long ld_partner;
} dsg_graph_agg_args;
int dsg_graph_agg(const dsg_graph_agg_args* a, void* stream);

/* ---- dsg_graph_agg_dadj ---------------------------------------------------------------------
 * dynamic: dadyn[n,u,w,kc] = sum_t p[n,t,u,kc] * dy[n,t,w,kc]      (fp32 out)           SURVEY §7.1
 * static : dA[k,u,w] += sum_{n,t,c} p[n,t,u,k*C+c] * dy[n,t,w,c]   (atomic, fp32, zero-initialised) */
typedef struct dsg_graph_agg_dadj_args {
    dsg_act_src p;
    dsg_act_src dy;
    int dtype;
    int is_static;
    int n_samples, T, V, KC, Ksub;
    float* dadj;
} dsg_graph_agg_dadj_args;
int dsg_graph_agg_dadj(const dsg_graph_agg_dadj_args* a, void* stream);

/* ---- dsg_ms_combine_fwd / dsg_ms_combine_bwd ------------------------------------------------
 * The tail of the multi-scale temporal unit's branches (tcn.py:390-396, 412-420): for the channel
 * ranges of the 'max' branch (3x1 max-pool, stride s, pad 1, on relu(bn(B))) and the '1x1' branch
 * (strided pass-through of B), and the already-convolved ranges in O, builds
 *   feat[n,t',v,c] = o[n,t',v,c] + o[n,t',V,c]*add_coeff[v]      (has_ext: dgmstcn; else feat = o: mstcn)
 * and accumulates BatchNorm statistics of feat.  oglob[n,t',c] = o[n,t',V,c] is kept for backward.
 * Backward: given dfeat (activation source), writes do (conv ranges, Vp rows per frame), the masked
 * gradients of the max / pass ranges into E (input frames, with BN-backward statistics for the max
 * range), and accumulates dadd_coeff[v]. */
typedef struct dsg_ms_combine_args {
    int dtype;
    int n_samples, T_in, T_out, stride, V, has_ext, C;
    int conv_lo, conv_hi;     /* channels produced by temporal convs (read from o) */
    int max_lo, max_hi;       /* 'max' branch channels (from b) */
    int pass_lo, pass_hi;     /* '1x1' branch channels (from b) */
    dsg_act_src b;            /* branch pre-activations with per-channel BN(+identity) coefficients, relu flag ignored */
    const void* o;            /* conv outputs [n,T_out,Vp,C] (dtype) */
    long long ld_o;
    const float* add_coeff;   /* [>=V] */
    void* feat;               /* [n,T_out,V,C] */
    long long ld_feat;
    float* oglob;             /* [n,T_out,C] fp32 */
    double* stat_sum;
    double* stat_sq;
    /* backward */
    dsg_act_src dfeat;
    void* d_o;                /* [n,T_out,Vp,C] */
    long long ld_do;
    void* e;                  /* [n,T_in,Vp,C] */
    long long ld_e;
    const void* b_raw;        /* partner for the BN-backward statistics of the max range */
    long long ld_b;
    double* e_sum;
    double* e_sq;
    float* dadd_coeff;
    int d_o_full;             /* 1: d_o holds all C channels (ld_do >= C): the gradient w.r.t. every branch output (joint-mean row
                                 included) is written by the per-output-frame pass and read back by the per-input-frame pass */
    int pad_;
} dsg_ms_combine_args;
int dsg_ms_combine_fwd(const dsg_ms_combine_args* a, void* stream);
int dsg_ms_combine_bwd(const dsg_ms_combine_args* a, void* stream);
/* parts: 1 = the per-output-frame pass only (d_o of the conv range incl. the joint-mean row, dadd_coeff); 2 = the per-input-frame
 * pass only (e of the max / pass ranges + their sums); 3 = both.  The split lets the conv branches' data gradient (dsg_ms_conv,
 * which pads its last 16-byte channel chunk with zeros) run between the two. */
int dsg_ms_combine_bwd_part(const dsg_ms_combine_args* a, int parts, void* stream);

/* ---- dsg_ms_temporal_fwd / _bwd_data / _bwd_weight -------------------------------------------------------------
 * The whole branch stage of the multi-scale temporal unit in one kernel per direction (tcn.py:383-396, 407-420):
 * dilated (3 x 1) convolutions of the conv branches as an implicit GEMM on tcgen05 (the post-BN-ReLU tile is staged
 * once in shared memory, each tap is a shifted UMMA descriptor into it), the 3x1 max-pool branch, the strided
 * pass-through branch and `local + global (x) add_coeff`, writing complete `feat` rows (and BatchNorm statistics for
 * transform.0).  bf16 only; shapes it does not take (fp32, V+1 > 32, kernel size != 3, > 8 branches) are run by the
 * per-branch path (dsg_conv_gemm + dsg_ms_combine_*).
 *   fwd        : b (raw branch pre-activations + BN coefficients) -> feat [n,T_out,V,C], oglob [n,T_out,C]
 *   bwd_data   : dfeat -> e [n,T_in,Vp,C] (gradient w.r.t. the pre-ReLU branch activations, masked; pass range: plain),
 *                BN-backward sums e_sum/e_sq (partner = b raw), dadd_coeff
 *   bwd_weight : dW/db of the temporal convolutions (atomic accumulation, zero-initialised by the caller) */
typedef struct dsg_ms_branch {
    int kind;             /* 0 = temporal conv, 1 = max-pool(3), 2 = pass-through ('1x1') */
    int lo, hi;           /* channel range */
    int dilation;
    const float* W;       /* [w, w, 3, 1] */
    const float* bias;    /* [w] */
    float* dW;
    float* db;
} dsg_ms_branch;

typedef struct dsg_ms_temporal_args {
    int n_samples, T_in, T_out, stride, V, has_ext, C, n_branches;
    dsg_ms_branch br[8];
    dsg_act_src b;        /* [n,T_in,Vp,C] bf16; a1/b1 = BN coefficients (identity on the pass range) */
    const float* add_coeff;
    void* feat;
    long long ld_feat;
    float* oglob;
    double* stat_sum;
    double* stat_sq;
    dsg_act_src dfeat;    /* [n,T_out,V,C] */
    void* e;
    long long ld_e;
    double* e_sum;
    double* e_sq;
    float* dadd_coeff;
    void* wpack;          /* caller-owned workspace of dsg_ms_temporal_wpack_bytes(): bf16 UMMA-ready weight tiles,
                             (re)written by _fwd and _bwd_data before use */
} dsg_ms_temporal_args;
int dsg_ms_temporal_supported(const dsg_ms_temporal_args* a);
long long dsg_ms_temporal_wpack_bytes(const dsg_ms_temporal_args* a);
int dsg_ms_temporal_fwd(const dsg_ms_temporal_args* a, void* stream);
int dsg_ms_temporal_bwd_data(const dsg_ms_temporal_args* a, void* stream);
int dsg_ms_temporal_bwd_weight(const dsg_ms_temporal_args* a, void* stream);

/* ---- dsg_ms_conv ----------------------------------------------------------------------------
 * All dilated (3 x 1) temporal convolutions of a multi-scale unit (the `unit_tcn(norm=None)` tails of the conv branches,
 * tcn.py:383-391) in ONE launch on the TMA-fed tcgen05 engine, or their data gradient:
 *   transposed = 0:  out[n,t',r,co] = bias[co] + sum_tap sum_ci W[co,ci,tap] * src[n, s*t' + (tap-1)*d, r, ci]   (zero padding in t)
 *                    src [n,T_in,Vr,.], out [n,T_out,Vr,.]
 *   transposed = 1:  out[n,t,r,ci]  = sum_tap sum_co W[co,ci,tap] * src[n, (t - (tap-1)*d)/s, r, co] where s divides,
 *                    src [n,T_out,Vr,.], out [n,T_in,Vr,.]; epilogue as dsg_conv_gemm: ReLU mask, BN-backward sums with partner
 * Branch j maps channels [lo,hi) of src to the same channels of out (block-diagonal); every tap is a tap-shifted TMA load of
 * the branch's channel window (4-D tensor map: frames outside a sample are zero-filled = the zero padding), so src must be a
 * plain bf16 tensor (the caller materialises relu(bn(.)) / the branch-output gradient once).  Channels of out outside every
 * branch range but inside [br[0].lo, br[last].hi) are written as zeros (+0 bias).  Returns 0 and sets *handled = 0 when the
 * shape is not taken (caller falls back to per-branch dsg_conv_gemm calls). */
typedef struct dsg_ms_conv_args {
    int n_samples, T_in, T_out, stride, Vr, transposed, n_branches;
    dsg_ms_branch br[8];  /* conv branches only (kind 0), ascending adjacent channel ranges, br[0].lo % 8 == 0 */
    const void* src;      /* bf16, channel 0 = absolute channel 0 of the branch layout */
    long long ld_src;
    void* out;
    long long ld_out;
    int has_mask;
    dsg_act_src mask;     /* absolute channels, like src */
    const void* partner;
    long long ld_partner;
    double* stat_sum;     /* absolute channels */
    double* stat_sq;
    void* wpack;          /* caller-owned workspace of dsg_ms_conv_wpack_bytes() */
} dsg_ms_conv_args;
long long dsg_ms_conv_wpack_bytes(const dsg_ms_conv_args* a);
int dsg_ms_conv(const dsg_ms_conv_args* a, int* handled, void* stream);
/* Weight / bias gradients of the same convolutions (tcn.py:383-391, backward) on the same engine:
 *   br[j].dW[co,ci,tap] += sum_{n,t',r} out[n,t',r,lo+co] * src[n, s*t' + (tap-1)*d, r, lo+ci],   br[j].db[co] += sum out[n,t',r,lo+co]
 * with src = the forward input of the convs (relu(bn(B)), [n,T_in,Vr,.]) and `out` = the gradient w.r.t. the branch outputs
 * ([n,T_out,Vr,.], read only), both plain bf16 with absolute channel indexing; transposed / mask / partner / stat_* / wpack are
 * ignored.  Accumulates with atomics (callers pre-zero dW / db).  *handled = 0: shape not taken (older engine: dsg_ms_temporal_bwd_weight). */
int dsg_ms_conv_wgrad(const dsg_ms_conv_args* a, int* handled, void* stream);

/* ---- dsg_pointwise --------------------------------------------------------------------------
 * out(r,c) = src(r,c) (any activation source: BN-apply, +residual, ReLU), optional mask
 * [mask(r,c) > 0], optional statistics (as in dsg_conv_gemm; partner may have its own dtype);
 * out == NULL runs the statistics only.  Replaces the stand-alone BatchNorm / add / ReLU modules at
 * unit and block ends (gcn.py:94,2365; tcn.py:424-428; dgstgcn.py:64-65), their backward, and
 * DGSTGCN.data_bn (dgstgcn.py:158-164): x[N,M,T,V,C] viewed as rows (n*M+m, t) of V*C channels is
 * already the channels-last activation, so BatchNorm1d over channel v*C+c is two point-wise passes
 * (statistics, then a*x+b with a dtype conversion) and the reference's two permutes disappear. */
typedef struct dsg_pointwise_args {
    dsg_act_src src;
    int dtype;            /* of src and mask */
    int C;
    long long rows;
    void* out;
    long long ld_out;
    int out_dtype;
    int has_mask;
    dsg_act_src mask;
    double* stat_sum;
    double* stat_sq;
    const void* partner;
    long long ld_partner;
    int partner_dtype;
    int pad_;
} dsg_pointwise_args;
int dsg_pointwise(const dsg_pointwise_args* a, void* stream);

/* ---- dsg_head_ce_fwd / dsg_head_ce_bwd ------------------------------------------------------
 * The classification head after the pooled backbone feature (heads/simple_head.py:93-96 fc_cls; losses/cross_entropy_loss.py:77-80
 * F.cross_entropy with class-index labels; core/evaluation top_k_accuracy), fp32:
 *   fwd: logits[n,k] = pooled[n,:] . W[k,:] + b[k];  stats[n*3 + {0,1,2}] = {cross-entropy of sample n, top-1 hit, top-5 hit}
 *        (stats / label may be NULL: scores only).  The batch means are one dsg_tmean over stats ([1, N, 1, 3]).
 *   bwd: dlogits[n,k] = (softmax(logits[n,:])[k] - [k == label[n]]) * gscale[0]  (gscale on the device: upstream gradient * loss_weight / N),
 *        dpooled[n,:] = dlogits[n,:] @ W (may be NULL), dW[k,:] += sum_n dlogits[n,k] pooled[n,:], db[k] += sum_n dlogits[n,k]
 *        (dW / db may be NULL; they ACCUMULATE: the caller pre-zeroes them or passes its gradient sink). */
int dsg_head_ce_fwd(const float* pooled, const float* W, const float* b, const long long* label, int N, int C, int K,
                    float* logits, float* stats, void* stream);
int dsg_head_ce_bwd(const float* logits, const long long* label, const float* pooled, const float* W, const float* gscale,
                    int N, int C, int K, float* dlogits, float* dpooled, float* dW, float* db, void* stream);

/* ---- dsg_sgd_step ---------------------------------------------------------------------------
 * Fused SGD (momentum, weight decay, Nesterov) over a flat fp32 parameter buffer:
 * configs/_init_/lr_schedual.py:11.  g = grad*grad_scale + wd*p; buf = mom*buf + g;
 * p -= lr*(nesterov ? g + mom*buf : buf). */
int dsg_sgd_step(float* p, const float* grad, float* buf, long long n, float lr, float momentum, float wd,
                 int nesterov, float grad_scale, void* stream);
/* The same update with the learning rate read from device memory (`lr_dev[0]`), so a CUDA-graph-captured step follows
 * the cosine schedule (configs/_init_/lr_schedual.py:12: CosineAnnealing, by_epoch=False) without re-capture. */
int dsg_sgd_step_dev(float* p, const float* grad, float* buf, long long n, const float* lr_dev, float momentum, float wd,
                     int nesterov, float grad_scale, void* stream);

/* Launch counters of the engines behind the entry points (diagnostics for the tests and bench.py: which engine ran).
 * id 0: TMA-fed tcgen05 GEMM (tc4)   1: TMA-fed tcgen05 weight gradient (tc4w)   2: fused adjacency-contraction + post GEMM
 *    3: tap-shifted temporal convolutions (dsg_ms_conv)   4: their TMA-fed weight gradient (dsg_ms_conv_wgrad).
 * Monotonic, process-wide, never read by the kernels. */
long long dsg_debug_counter(int id);

const char* dsg_last_error(void);
int dsg_abi_version(void);
/* 1 when built for the GPU (sm_100a), 0 for the host-side simulator used by the CPU test-suite. */
int dsg_is_device_build(void);

#ifdef __cplusplus
}
#endif
#endif /* DSGCN_B200_H */
