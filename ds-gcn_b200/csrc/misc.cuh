// BatchNorm coefficient finalisation, temporal mean, fused point-wise pass, multi-scale branch
// combine (max-pool / pass-through / joint-mean column) and the fused SGD update.
#pragma once
#include "dsg_common.h"

namespace dsg {

DSG_D float ld_any(const void* p, int dtype, long long i) {
    return dtype == DSG_BF16 ? __bfloat162float(reinterpret_cast<const bf16*>(p)[i]) : reinterpret_cast<const float*>(p)[i];
}
DSG_D void st_any(void* p, int dtype, long long i, float v) {
    if (dtype == DSG_BF16) reinterpret_cast<bf16*>(p)[i] = __float2bfloat16(v);
    else reinterpret_cast<float*>(p)[i] = v;
}

// ---------------------------------------------------------------------------------------------
constexpr int BNF_MAX_JOBS = 8;
struct BnJobs { dsg_bn_job j[BNF_MAX_JOBS]; };

__global__ void bn_finalize_kernel(BnJobs jobs) {
    const dsg_bn_job& J = jobs.j[blockIdx.y];
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= J.C) return;
    const float gamma = J.gamma ? J.gamma[c] : 1.f;
    const float beta = J.beta ? J.beta[c] : 0.f;
    if (J.mode == 0) {
        double mean = J.sum[c] / J.count;
        double var = J.sq[c] / J.count - mean * mean;
        if (var < 0) var = 0;
        float invstd = (float)(1.0 / sqrt(var + (double)J.eps));
        float av = gamma * invstd;
        J.a[c] = av;
        J.b[c] = beta - (float)mean * av;
        if (J.save_mean) J.save_mean[c] = (float)mean;
        if (J.save_invstd) J.save_invstd[c] = invstd;
        if (J.running_mean) {
            double unb = J.count > 1 ? var * J.count / (J.count - 1) : var;
            J.running_mean[c] = (1.f - J.momentum) * J.running_mean[c] + J.momentum * (float)mean;
            J.running_var[c] = (1.f - J.momentum) * J.running_var[c] + J.momentum * (float)unb;
        }
    } else if (J.mode == 1) {
        float invstd = 1.f / sqrtf(J.running_var[c] + J.eps);
        float av = gamma * invstd;
        J.a[c] = av;
        J.b[c] = beta - J.running_mean[c] * av;
        if (J.save_mean) J.save_mean[c] = J.running_mean[c];
        if (J.save_invstd) J.save_invstd[c] = invstd;
    } else if (J.mode == 2 || J.mode == 3) {
        double s1 = J.sum[c], s2r = J.sq[c];
        double mean = J.save_mean[c], invstd = J.save_invstd[c];
        double dgamma = invstd * (s2r - mean * s1);
        if (J.dgamma) J.dgamma[c] = (float)dgamma;
        if (J.dbeta) J.dbeta[c] = (float)s1;
        double ca = gamma * invstd, cb = 0, cc = 0;
        if (J.mode == 2) {
            cb = -gamma * invstd * invstd * dgamma / J.count;
            cc = -ca * s1 / J.count - cb * mean;
        }
        J.a[c] = (float)ca;
        J.b[c] = (float)cb;
        J.c[c] = (float)cc;
    } else {
        J.a[c] = 1.f;
        J.b[c] = 0.f;
        if (J.c) J.c[c] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------
template <class T>
__global__ void tmean_kernel(const T* x, long long ld, int T_, int V, int C, float* xm, bf16* xm_bf) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // v*C + c
    if (idx >= V * C) return;
    const int c = idx % C, v = idx / C;
    float s = 0.f;
    for (int t = 0; t < T_; ++t) s += ldf<T>(x + (((long long)n * T_ + t) * V + v) * ld + c);
    xm[(long long)n * V * C + idx] = s / (float)T_;
    if (xm_bf) xm_bf[(long long)n * V * C + idx] = __float2bfloat16(s / (float)T_);
}

// bf16 fast path: a thread owns (joint, 8 channels) and walks the frames with 4 independent 16-byte loads in flight
__global__ void tmean_vec_kernel(const bf16* x, long long ld, int T_, int V, int C, float* xm, bf16* xm_bf) {
    const int n = blockIdx.y;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;   // v*(C/8) + chunk
    const int nch = C >> 3;
    if (idx >= V * nch) return;
    const int ch = idx % nch, v = idx / nch;
    const bf16* p = x + ((long long)n * T_ * V + v) * ld + ch * 8;
    const long long fstep = (long long)V * ld;
    float s[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) s[e] = 0.f;
    int t = 0;
    for (; t + 4 <= T_; t += 4) {
        uint4 r[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) r[b] = *reinterpret_cast<const uint4*>(p + (t + b) * fstep);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            float f[8];
            unpack8(r[b], f);
#pragma unroll
            for (int e = 0; e < 8; ++e) s[e] += f[e];
        }
    }
    for (; t < T_; ++t) {
        float f[8];
        unpack8(*reinterpret_cast<const uint4*>(p + t * fstep), f);
#pragma unroll
        for (int e = 0; e < 8; ++e) s[e] += f[e];
    }
    float* dst = xm + (long long)n * V * C + v * C + ch * 8;
    const float inv = 1.f / (float)T_;
#pragma unroll
    for (int e = 0; e < 8; ++e) { s[e] *= inv; dst[e] = s[e]; }
    if (xm_bf) *reinterpret_cast<uint4*>(xm_bf + (long long)n * V * C + v * C + ch * 8) = pack8(s);
}

// ---------------------------------------------------------------------------------------------
constexpr int PW_THREADS = 256;
constexpr int PW_CT = 64;        // channels per CTA
constexpr int PW_ROWS = 128;     // rows per CTA

template <class T>
__global__ void __launch_bounds__(PW_THREADS) pointwise_kernel(dsg_pointwise_args a) {
    DSG_SHARED float s_red[2][PW_THREADS / PW_CT][PW_CT];
    const int tid = threadIdx.x, cl = tid % PW_CT, rg = tid / PW_CT;
    const int c = blockIdx.y * PW_CT + cl;
    const long long r0 = (long long)blockIdx.x * PW_ROWS;
    float s1 = 0.f, s2 = 0.f;
    if (c < a.C) {
        for (int i = rg; i < PW_ROWS; i += PW_THREADS / PW_CT) {
            long long r = r0 + i;
            if (r >= a.rows) break;
            float v = act_value<T>(a.src, r, c);
            if (a.has_mask && !(act_value<T>(a.mask, r, c) > 0.f)) v = 0.f;
            if (a.stat_sum) {
                float p = a.partner ? ld_any(a.partner, a.partner_dtype, r * a.ld_partner + c) : v;
                s1 += v;
                s2 += v * p;
            }
            if (a.out) st_any(a.out, a.out_dtype, r * a.ld_out + c, v);
        }
    }
    if (a.stat_sum) {
        s_red[0][rg][cl] = s1;
        s_red[1][rg][cl] = s2;
        __syncthreads();
        if (tid < PW_CT && blockIdx.y * PW_CT + tid < a.C) {
            float t1 = 0.f, t2 = 0.f;
            for (int g = 0; g < PW_THREADS / PW_CT; ++g) { t1 += s_red[0][g][tid]; t2 += s_red[1][g][tid]; }
            atomicAdd(a.stat_sum + blockIdx.y * PW_CT + tid, (double)t1);
            atomicAdd(a.stat_sq + blockIdx.y * PW_CT + tid, (double)t2);
        }
    }
}

// Vectorised fast path (bf16 in / bf16 out, everything 16-byte aligned, C % 8 == 0): a thread owns one 8-channel
// chunk and walks rows; 8 consecutive threads cover 64 channels = 128 B per row.  A pure streaming kernel, so what
// matters is bytes in flight: per-channel coefficients live in shared memory (registers are for data), the loads of U
// rows are issued before any of them is used, and the register budget keeps 3 (SIMPLE: one tensor per source) or 2
// CTAs per SM resident (2 for both variants).
constexpr int PV_ROWS = 256;     // rows per CTA
DSG_D void pv_eval(const uint4& x1, const uint4& x2, bool has_x2, int relu, const float* cf /* [3][PW_CT] */, int c8, float* v) {
    float t[8], c1[8], cb[8];
    unpack8(x1, t);
    load8f(cf + c8, c1, 1.f);
    load8f(cf + PW_CT + c8, cb, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(t[j], c1[j], cb[j]);
    if (has_x2) {
        unpack8(x2, t);
        load8f(cf + 2 * PW_CT + c8, c1, 1.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(t[j], c1[j], v[j]);
    }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
}
DSG_D void pv_stage_coefs(const ActSrc& s, bool on, int c0, int C, float* cf /* [3][PW_CT] */) {
    for (int i = threadIdx.x; i < PW_CT; i += PW_THREADS) {
        const int ch = c0 + i;
        const bool in = on && ch < C;
        cf[i] = (in && s.a1) ? s.a1[ch] : 1.f;
        cf[PW_CT + i] = ((in && s.b1) ? s.b1[ch] : 0.f) + ((in && s.b2) ? s.b2[ch] : 0.f);
        cf[2 * PW_CT + i] = (in && s.a2) ? s.a2[ch] : 1.f;
    }
}

template <bool SIMPLE>
__global__ void __launch_bounds__(PW_THREADS, 2) pointwise_vec_kernel(dsg_pointwise_args a) {
    constexpr int U = SIMPLE ? 4 : 2;
    DSG_SHARED float s_red[2][PW_THREADS / 8][PW_CT];
    DSG_SHARED __align__(16) float cf_s[3 * PW_CT];
    DSG_SHARED __align__(16) float cf_m[3 * PW_CT];
    const int tid = threadIdx.x, cc = tid & 7, rl = tid >> 3;         // 8 chunks x 32 row lanes
    const int c0 = blockIdx.y * PW_CT, c = c0 + cc * 8;
    const long long r0 = (long long)blockIdx.x * PV_ROWS;
    pv_stage_coefs(a.src, true, c0, a.C, cf_s);
    pv_stage_coefs(a.mask, a.has_mask != 0, c0, a.C, cf_m);
    __syncthreads();
    float s1[8], s2[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (c < a.C) {
        const bf16* x1p = reinterpret_cast<const bf16*>(a.src.x1) + c;
        const bf16* x2p = (!SIMPLE && a.src.x2) ? reinterpret_cast<const bf16*>(a.src.x2) + c : nullptr;
        const bf16* m1p = a.has_mask ? reinterpret_cast<const bf16*>(a.mask.x1) + c : nullptr;
        const bf16* m2p = (!SIMPLE && a.has_mask && a.mask.x2) ? reinterpret_cast<const bf16*>(a.mask.x2) + c : nullptr;
        const bf16* partner = a.partner ? reinterpret_cast<const bf16*>(a.partner) + c : nullptr;
        bf16* out = a.out ? reinterpret_cast<bf16*>(a.out) + c : nullptr;
        uint4 z4;
        z4.x = z4.y = z4.z = z4.w = 0u;
        for (int i0 = 0; i0 < PV_ROWS / (PW_THREADS / 8); i0 += U) {
            uint4 x1[U], x2[U], m1[U], m2[U], pp[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const long long r = r0 + rl + (long long)(i0 + u) * (PW_THREADS / 8);
                ok[u] = r < a.rows;
                x1[u] = x2[u] = m1[u] = m2[u] = pp[u] = z4;
                if (ok[u]) {
                    x1[u] = *reinterpret_cast<const uint4*>(x1p + r * a.src.ld1);
                    if (!SIMPLE && x2p) x2[u] = *reinterpret_cast<const uint4*>(x2p + r * a.src.ld2);
                    if (m1p) m1[u] = *reinterpret_cast<const uint4*>(m1p + r * a.mask.ld1);
                    if (!SIMPLE && m2p) m2[u] = *reinterpret_cast<const uint4*>(m2p + r * a.mask.ld2);
                    if (partner) pp[u] = *reinterpret_cast<const uint4*>(partner + r * a.ld_partner);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (!ok[u]) continue;
                const long long r = r0 + rl + (long long)(i0 + u) * (PW_THREADS / 8);
                float v[8];
                pv_eval(x1[u], x2[u], !SIMPLE && x2p != nullptr, a.src.relu, cf_s, cc * 8, v);
                if (m1p) {
                    float m[8];
                    pv_eval(m1[u], m2[u], !SIMPLE && m2p != nullptr, a.mask.relu, cf_m, cc * 8, m);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
                }
                if (a.stat_sum) {
                    float p[8];
                    if (partner) unpack8(pp[u], p);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] += v[j] * (partner ? p[j] : v[j]); }
                }
                if (out) *reinterpret_cast<uint4*>(out + r * a.ld_out) = pack8(v);
            }
        }
    }
    if (a.stat_sum) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { s_red[0][rl][cc * 8 + j] = s1[j]; s_red[1][rl][cc * 8 + j] = s2[j]; }
        __syncthreads();
        if (tid < PW_CT && blockIdx.y * PW_CT + tid < a.C) {
            float t1 = 0.f, t2 = 0.f;
            for (int g = 0; g < PW_THREADS / 8; ++g) { t1 += s_red[0][g][tid]; t2 += s_red[1][g][tid]; }
            atomicAdd(a.stat_sum + blockIdx.y * PW_CT + tid, (double)t1);
            atomicAdd(a.stat_sq + blockIdx.y * PW_CT + tid, (double)t2);
        }
    }
}

static inline bool pointwise_vec_ok(const dsg_pointwise_args& a) {
    if (a.dtype != DSG_BF16 || a.C % 8 != 0 || !act8_ok(a.src)) return false;
    if (a.out && (a.out_dtype != DSG_BF16 || (uintptr_t)a.out % 16 != 0 || a.ld_out % 8 != 0)) return false;
    if (a.has_mask && !act8_ok(a.mask)) return false;
    if (a.partner && (a.partner_dtype != DSG_BF16 || (uintptr_t)a.partner % 16 != 0 || a.ld_partner % 8 != 0)) return false;
    return true;
}

// ---------------------------------------------------------------------------------------------
// multi-scale combine.  Channel c belongs to the conv range (value read from o), the max range
// (3x1 max-pool over relu(a*B+b)) or the pass range (a*B+b at frame s*t').
DSG_D int ms_kind(const dsg_ms_combine_args& a, int c) {
    if (c >= a.conv_lo && c < a.conv_hi) return 0;
    if (c >= a.max_lo && c < a.max_hi) return 1;
    if (c >= a.pass_lo && c < a.pass_hi) return 2;
    return 3;
}

// branch output O at (n, t', column j in [0,Vp), c)
template <class T>
DSG_D float ms_branch_out(const dsg_ms_combine_args& a, int kind, int n, int tp, int j, int c, int Vp) {
    if (kind == 0) return ldf<T>(reinterpret_cast<const T*>(a.o) + (((long long)n * a.T_out + tp) * Vp + j) * a.ld_o + c);
    if (kind == 2) return act_value<T>(a.b, ((long long)n * a.T_in + tp * a.stride) * Vp + j, c);
    if (kind == 1) {
        float m = -3.0e38f;
        for (int dt = -1; dt <= 1; ++dt) {
            int t = tp * a.stride + dt;
            if (t < 0 || t >= a.T_in) continue;
            float h = fmaxf(act_value<T>(a.b, ((long long)n * a.T_in + t) * Vp + j, c), 0.f);
            m = fmaxf(m, h);
        }
        return m;
    }
    return 0.f;
}

template <class T>
__global__ void __launch_bounds__(PW_THREADS) ms_combine_fwd_kernel(dsg_ms_combine_args a) {
    DSG_SHARED float s_red[2][PW_THREADS / PW_CT][PW_CT];
    const int tid = threadIdx.x, cl = tid % PW_CT, rg = tid / PW_CT;
    const int c = blockIdx.y * PW_CT + cl;
    const int Vp = a.V + a.has_ext;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    float s1 = 0.f, s2 = 0.f;
    // one frame (n,t') per row group iteration
    for (long long f = (long long)blockIdx.x * 8 + rg; f < n_frames && f < (long long)(blockIdx.x + 1) * 8; f += PW_THREADS / PW_CT) {
        if (c >= a.C) break;
        const int n = (int)(f / a.T_out), tp = (int)(f - (long long)n * a.T_out);
        const int kind = ms_kind(a, c);
        float glob = 0.f;
        if (a.has_ext) {
            glob = ms_branch_out<T>(a, kind, n, tp, a.V, c, Vp);
            a.oglob[f * a.C + c] = glob;
        }
        for (int v = 0; v < a.V; ++v) {
            float val = ms_branch_out<T>(a, kind, n, tp, v, c, Vp);
            if (a.has_ext) val = fmaf(glob, a.add_coeff[v], val);
            s1 += val;
            s2 += val * val;
            stf<T>(reinterpret_cast<T*>(a.feat) + (f * a.V + v) * a.ld_feat + c, val);
        }
    }
    if (a.stat_sum) {
        s_red[0][rg][cl] = s1;
        s_red[1][rg][cl] = s2;
        __syncthreads();
        if (tid < PW_CT && blockIdx.y * PW_CT + tid < a.C) {
            float t1 = 0.f, t2 = 0.f;
            for (int g = 0; g < PW_THREADS / PW_CT; ++g) { t1 += s_red[0][g][tid]; t2 += s_red[1][g][tid]; }
            atomicAdd(a.stat_sum + blockIdx.y * PW_CT + tid, (double)t1);
            atomicAdd(a.stat_sq + blockIdx.y * PW_CT + tid, (double)t2);
        }
    }
}

// backward part 1: per output frame (n,t'): d_o for the conv range (V local rows + the joint-mean row),
// and dadd_coeff[v] += sum_c dfeat[v,c]*oglob[c]
template <class T>
__global__ void __launch_bounds__(PW_THREADS) ms_combine_bwd_o_kernel(dsg_ms_combine_args a) {
    DSG_SHARED float s_dadd[32];
    const int tid = threadIdx.x, cl = tid % PW_CT, rg = tid / PW_CT;
    const int c = blockIdx.y * PW_CT + cl;
    const int Vp = a.V + a.has_ext;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (tid < 32) s_dadd[tid] = 0.f;
    __syncthreads();
    for (int i = 0; i < 8 / (PW_THREADS / PW_CT); ++i) {
        long long f = (long long)blockIdx.x * 8 + rg + i * (PW_THREADS / PW_CT);
        const bool ok = f < n_frames && c < a.C;
        const int kind = ok ? ms_kind(a, c) : 3;
        float og = (ok && a.has_ext) ? a.oglob[f * a.C + c] : 0.f;
        float gsum = 0.f;
        for (int v = 0; v < a.V; ++v) {
            float d = ok ? act_value<T>(a.dfeat, f * a.V + v, c) : 0.f;
            if (a.has_ext) {
                gsum = fmaf(d, a.add_coeff[v], gsum);
                float t = warp_sum(d * og);                   // all lanes participate
                if ((tid & 31) == 0) atomicAdd(&s_dadd[v], t);
            }
            if (ok && kind == 0) stf<T>(reinterpret_cast<T*>(a.d_o) + (f * Vp + v) * a.ld_do + c, d);
        }
        if (ok && kind == 0 && a.has_ext) stf<T>(reinterpret_cast<T*>(a.d_o) + (f * Vp + a.V) * a.ld_do + c, gsum);
    }
    if (a.has_ext) {
        __syncthreads();
        if (tid < a.V) atomicAdd(a.dadd_coeff + tid, s_dadd[tid]);
    }
}

// gradient w.r.t. branch output O at (n,t',column j,c), from dfeat
template <class T>
DSG_D float ms_dout(const dsg_ms_combine_args& a, int n, int tp, int j, int c) {
    const long long f = (long long)n * a.T_out + tp;
    if (j < a.V) return act_value<T>(a.dfeat, f * a.V + j, c);
    float s = 0.f;
    for (int v = 0; v < a.V; ++v) s = fmaf(act_value<T>(a.dfeat, f * a.V + v, c), a.add_coeff[v], s);
    return s;
}

// backward part 2: per input frame (n,t): E for the max and pass ranges (+ BN-backward sums for max)
template <class T>
__global__ void __launch_bounds__(PW_THREADS) ms_combine_bwd_e_kernel(dsg_ms_combine_args a, int c_lo, int c_hi) {
    DSG_SHARED float s_red[2][PW_THREADS / PW_CT][PW_CT];
    const int tid = threadIdx.x, cl = tid % PW_CT, rg = tid / PW_CT;
    const int c = c_lo + blockIdx.y * PW_CT + cl;
    const int Vp = a.V + a.has_ext;
    const long long n_frames = (long long)a.n_samples * a.T_in;
    float s1 = 0.f, s2 = 0.f;
    for (long long f = (long long)blockIdx.x * 8 + rg; f < n_frames && f < (long long)(blockIdx.x + 1) * 8; f += PW_THREADS / PW_CT) {
        if (c >= c_hi) break;
        const int n = (int)(f / a.T_in), t = (int)(f - (long long)n * a.T_in);
        const int kind = ms_kind(a, c);
        for (int j = 0; j < Vp; ++j) {
            const long long r = f * Vp + j;
            float e = 0.f;
            if (kind == 2) {
                if (t % a.stride == 0 && t / a.stride < a.T_out) e = ms_dout<T>(a, n, t / a.stride, j, c);
            } else if (kind == 1) {
                const float hc = fmaxf(act_value<T>(a.b, r, c), 0.f);
                if (hc > 0.f) {
                    for (int dt = -1; dt <= 1; ++dt) {            // windows t' with s*t' + dt == t
                        int num = t - dt;
                        if (num < 0 || num % a.stride != 0) continue;
                        int tp = num / a.stride;
                        if (tp >= a.T_out) continue;
                        // arg-max of window tp (first maximum wins, as in ATen max_pool2d)
                        float m = -3.0e38f; int am = -2;
                        for (int d2 = -1; d2 <= 1; ++d2) {
                            int t2 = tp * a.stride + d2;
                            if (t2 < 0 || t2 >= a.T_in) continue;
                            float h2 = fmaxf(act_value<T>(a.b, ((long long)n * a.T_in + t2) * Vp + j, c), 0.f);
                            if (h2 > m) { m = h2; am = d2; }
                        }
                        if (am == dt) e += ms_dout<T>(a, n, tp, j, c);
                    }
                }
                s1 += e;
                s2 += e * ldf<T>(reinterpret_cast<const T*>(a.b_raw) + r * a.ld_b + c);
            }
            stf<T>(reinterpret_cast<T*>(a.e) + r * a.ld_e + c, e);
        }
    }
    if (a.e_sum) {
        s_red[0][rg][cl] = s1;
        s_red[1][rg][cl] = s2;
        __syncthreads();
        const int cc = c_lo + blockIdx.y * PW_CT + tid;
        if (tid < PW_CT && cc < c_hi && ms_kind(a, cc) == 1) {
            float t1 = 0.f, t2 = 0.f;
            for (int g = 0; g < PW_THREADS / PW_CT; ++g) { t1 += s_red[0][g][tid]; t2 += s_red[1][g][tid]; }
            atomicAdd(a.e_sum + cc, (double)t1);
            atomicAdd(a.e_sq + cc, (double)t2);
        }
    }
}

// ---------------------------------------------------------------------------------------------
__global__ void sgd_kernel(float* p, const float* g, float* buf, long long n, float lr, const float* lr_dev, float mom, float wd, int nesterov, float gscale) {
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (lr_dev) lr = lr_dev[0];
    float gr = g[i] * gscale + wd * p[i];
    float b = mom * buf[i] + gr;
    buf[i] = b;
    p[i] -= lr * (nesterov ? gr + mom * b : b);
}

}  // namespace dsg
