// Adjacency contraction kernels (dynamic per-sample/per-channel and static per-subset) and the
// adjacency gradients.  CUDA-core kernels: per channel this is a [T x V] x [V x V] product with a
// distinct V x V operand, HBM/L2-bound and far below any MMA tile (SURVEY.md "hard parts" 4).
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int AG_THREADS = 256;
constexpr int AG_TF = 4;       // frames processed together per warp (dynamic modes)

// Fused tail shared by the aggregation kernels: mask, statistics (per lane == per channel), store.
template <class T>
DSG_D void agg_store(const dsg_graph_agg_args& a, long long orow, int ch, float v, float& s1, float& s2) {
    if (a.has_mask && !(act_value<T>(a.mask, orow, ch) > 0.f)) v = 0.f;
    if (a.stat_sum) {
        float p = a.partner ? ldf<T>(reinterpret_cast<const T*>(a.partner) + orow * a.ld_partner + ch) : v;
        s1 += v;
        s2 += v * p;
    }
    stf<T>(reinterpret_cast<T*>(a.out) + orow * a.ld_out + ch, v);
}

// 8 consecutive channels of an activation source (vector path for aligned bf16, scalar otherwise) -> fp32
template <class T> DSG_D void agg_load8(const ActSrc& s, long long row, int c, int C, bool vec, float* v);
template <> DSG_D void agg_load8<bf16>(const ActSrc& s, long long row, int c, int C, bool vec, float* v) {
    if (vec && c + 8 <= C) {
        unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x1) + row * s.ld1 + c), v);
        if (s.a1) { float t[8]; load8f(s.a1 + c, t, 1.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] *= t[j]; }
        if (s.b1) { float t[8]; load8f(s.b1 + c, t, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j]; }
        if (s.x2) {
            float w[8];
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x2) + row * s.ld2 + c), w);
            if (s.a2) { float t[8]; load8f(s.a2 + c, t, 1.f);
#pragma unroll
                for (int j = 0; j < 8; ++j) w[j] *= t[j]; }
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += w[j];
        }
        if (s.b2) { float t[8]; load8f(s.b2 + c, t, 0.f);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += t[j]; }
        if (s.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = (c + j < C) ? act_value<bf16>(s, row, c + j) : 0.f;
    }
}
template <> DSG_D void agg_load8<float>(const ActSrc& s, long long row, int c, int C, bool, float* v) {
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c + j < C) ? act_value<float>(s, row, c + j) : 0.f;
}

// per-channel coefficients (a1, b1+b2, a2) of a source for the 32-channel slice starting at c0, staged in shared memory
DSG_D void agg_stage_coefs(const ActSrc& s, int c0, int C, float* cf /* [3][32] */) {
    const int l = threadIdx.x;
    if (l < 32) {
        const int ch = c0 + l;
        const bool in = ch < C;
        cf[l] = (in && s.a1) ? s.a1[ch] : 1.f;
        cf[32 + l] = ((in && s.b1) ? s.b1[ch] : 0.f) + ((in && s.b2) ? s.b2[ch] : 0.f);
        cf[64 + l] = (in && s.a2) ? s.a2[ch] : 1.f;
    }
}
// 8 channels at slice offset q*8 with staged coefficients (vector path) or the generic path
template <class T> DSG_D void agg_load8_s(const ActSrc& s, long long row, int c0, int q, int C, bool vec, const float* cf, float* v) {
    agg_load8<T>(s, row, c0 + q * 8, C, vec, v);
}
template <> DSG_D void agg_load8_s<bf16>(const ActSrc& s, long long row, int c0, int q, int C, bool vec, const float* cf, float* v) {
    const int c = c0 + q * 8;
    if (vec && c + 8 <= C) {
        float x[8];
        unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x1) + row * s.ld1 + c), x);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], cf[q * 8 + j], cf[32 + q * 8 + j]);
        if (s.x2) {
            unpack8(*reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x2) + row * s.ld2 + c), x);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], cf[64 + q * 8 + j], v[j]);
        }
        if (s.relu) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
    } else agg_load8<bf16>(s, row, c, C, false, v);
}

constexpr int AG_TCH = 8;      // frames staged per step
constexpr int AG_DYN_THREADS = 512;

// Dynamic contraction: the per-sample adjacency slice (32 channels) and AG_TCH frames of the operand live in shared
// memory (fp32, channel fastest: conflict-free); a warp owns joints w, w+8, ... and keeps the adjacency column in
// registers, so the inner loop is one shared load + FMA per (frame, source joint).
template <class T, int V, int NT, int WB>
__global__ void __launch_bounds__(NT, 2) agg_dyn_kernel(dsg_graph_agg_args a, int t_chunk, int vec) {
    DSG_DYN_SMEM(smem_raw);
    float* adj = reinterpret_cast<float*>(smem_raw);      // [V*V][32]
    float* Ps = adj + V * V * 32;                         // [AG_TCH][V][32]
    DSG_SHARED float s_red[2][NT / 32][32];
    DSG_SHARED float cf_src[96];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, kc0 = blockIdx.y * 32;
    const int tbeg = blockIdx.z * t_chunk;
    const int tend = tbeg + t_chunk < a.T ? tbeg + t_chunk : a.T;
    const int ch = kc0 + lane;
    const bool ch_ok = ch < a.KC;
    agg_stage_coefs(a.src, kc0, a.KC, cf_src);
    // this lane's mask coefficients (the mask is evaluated per output element of one channel)
    float mk_a1 = 1.f, mk_b = 0.f, mk_a2 = 1.f;
    if (a.has_mask && ch_ok) {
        if (a.mask.a1) mk_a1 = a.mask.a1[ch];
        if (a.mask.b1) mk_b += a.mask.b1[ch];
        if (a.mask.b2) mk_b += a.mask.b2[ch];
        if (a.mask.a2) mk_a2 = a.mask.a2[ch];
    }
    {   // adjacency slice (transposed on the fly for the gradient w.r.t. p)
        const T* adyn = reinterpret_cast<const T*>(a.adyn) + (long long)n * V * V * a.KC;
        ActSrc as{};
        as.x1 = adyn; as.ld1 = a.KC;
        const bool avec = vec && ((uintptr_t)a.adyn % 16 == 0) && (a.KC % 8 == 0);
        for (int idx0 = tid; idx0 < V * V * 4; idx0 += NT * 4) {      // 4 independent loads in flight
            float v[4][8];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int idx = idx0 + b * NT;
                if (idx < V * V * 4) agg_load8<T>(as, idx >> 2, kc0 + (idx & 3) * 8, a.KC, avec, v[b]);
            }
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int idx = idx0 + b * NT;
                if (idx >= V * V * 4) continue;
                const int q = idx & 3, uw = idx >> 2;
                int dst = uw;
                if (a.mode == 1) { int u = uw / V, w = uw - u * V; dst = w * V + u; }
                float4* d4 = reinterpret_cast<float4*>(adj + dst * 32 + q * 8);
                d4[0] = make_float4(v[b][0], v[b][1], v[b][2], v[b][3]);
                d4[1] = make_float4(v[b][4], v[b][5], v[b][6], v[b][7]);
            }
        }
    }
    float s1 = 0.f, s2 = 0.f;
    for (int t0 = tbeg; t0 < tend; t0 += AG_TCH) {
        __syncthreads();
        if (sizeof(T) == 2 && vec) {
            // vector path: all of this thread's 16-byte loads of the chunk are issued before the first one is used
            constexpr int NIT = (AG_TCH * V * 4 + NT - 1) / NT;
            const bf16* X1 = reinterpret_cast<const bf16*>(a.src.x1);
            const bf16* X2 = reinterpret_cast<const bf16*>(a.src.x2);
            uint4 rx1[NIT], rx2[NIT];
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int idx = tid + i * NT;
                const int q = idx & 3, rv = idx >> 2, tt = rv / V, c = kc0 + q * 8;
                rx1[i] = rx2[i] = make_uint4(0u, 0u, 0u, 0u);
                if (idx < AG_TCH * V * 4 && t0 + tt < tend && c + 8 <= a.KC) {
                    const long long row = ((long long)n * a.T + t0) * V + rv;
                    rx1[i] = *reinterpret_cast<const uint4*>(X1 + row * a.src.ld1 + c);
                    if (X2) rx2[i] = *reinterpret_cast<const uint4*>(X2 + row * a.src.ld2 + c);
                }
            }
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const int idx = tid + i * NT;
                if (idx >= AG_TCH * V * 4) continue;
                const int q = idx & 3, rv = idx >> 2, tt = rv / V, c = kc0 + q * 8;
                float v[8];
                if (c + 8 <= a.KC) {
                    float x[8];
                    unpack8(rx1[i], x);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], cf_src[q * 8 + j], cf_src[32 + q * 8 + j]);
                    if (X2) {
                        unpack8(rx2[i], x);
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], cf_src[64 + q * 8 + j], v[j]);
                    }
                    if (a.src.relu) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
                    }
                    if (t0 + tt >= tend) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) v[j] = 0.f;
                    }
                } else if (t0 + tt < tend) agg_load8_s<T>(a.src, ((long long)n * a.T + t0) * V + rv, kc0, q, a.KC, true, cf_src, v);
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
                float4* dst = reinterpret_cast<float4*>(Ps + rv * 32 + q * 8);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        } else {
            for (int idx = tid; idx < AG_TCH * V * 4; idx += NT) {
                const int q = idx & 3, rv = idx >> 2;          // rv = tt*V + u
                const int tt = rv / V;
                float v[8];
                if (t0 + tt < tend) agg_load8_s<T>(a.src, ((long long)n * a.T + t0) * V + rv, kc0, q, a.KC, vec != 0, cf_src, v);
                else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = 0.f;
                }
                float4* dst = reinterpret_cast<float4*>(Ps + rv * 32 + q * 8);
                dst[0] = make_float4(v[0], v[1], v[2], v[3]);
                dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            }
        }
        __syncthreads();
        // a warp owns WB joints at a time and keeps their adjacency columns in registers: one shared load of
        // p[tt][u] feeds WB FMAs (the loop is bound by shared-memory bandwidth, one 128-byte wavefront per load)
        for (int g = warp; g < (V + WB - 1) / WB; g += NT / 32) {
            const int w0 = g * WB;
            float av[WB][V];
#pragma unroll
            for (int bq = 0; bq < WB; ++bq) {
#pragma unroll
                for (int u = 0; u < V; ++u) av[bq][u] = (w0 + bq < V) ? adj[(u * V + w0 + bq) * 32 + lane] : 0.f;
            }
#pragma unroll 2
            for (int tt = 0; tt < AG_TCH; ++tt) {
                if (t0 + tt >= tend) break;
                float accv[WB];
#pragma unroll
                for (int bq = 0; bq < WB; ++bq) accv[bq] = 0.f;
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    const float pv = Ps[(tt * V + u) * 32 + lane];
#pragma unroll
                    for (int bq = 0; bq < WB; ++bq) accv[bq] = fmaf(pv, av[bq][u], accv[bq]);
                }
                if (ch_ok) {
#pragma unroll
                    for (int bq = 0; bq < WB; ++bq) {
                        const int w = w0 + bq;
                        if (w >= V) continue;
                        float acc = accv[bq];
                        const long long orow = ((long long)n * a.T + t0 + tt) * V + w;
                        if (a.has_mask) {
                            float mv = fmaf(ldf<T>(reinterpret_cast<const T*>(a.mask.x1) + orow * a.mask.ld1 + ch), mk_a1, mk_b);
                            if (a.mask.x2) mv = fmaf(ldf<T>(reinterpret_cast<const T*>(a.mask.x2) + orow * a.mask.ld2 + ch), mk_a2, mv);
                            if (!(mv > 0.f)) acc = 0.f;
                        }
                        if (a.stat_sum) {
                            const float p = a.partner ? ldf<T>(reinterpret_cast<const T*>(a.partner) + orow * a.ld_partner + ch) : acc;
                            s1 += acc;
                            s2 += acc * p;
                        }
                        stf<T>(reinterpret_cast<T*>(a.out) + orow * a.ld_out + ch, acc);
                    }
                }
            }
        }
    }
    if (a.stat_sum) {
        s_red[0][warp][lane] = s1;
        s_red[1][warp][lane] = s2;
        __syncthreads();
        if (tid < 32 && kc0 + tid < a.KC) {
            float t1 = 0.f, t2 = 0.f;
            for (int w = 0; w < NT / 32; ++w) { t1 += s_red[0][w][tid]; t2 += s_red[1][w][tid]; }
            atomicAdd(a.stat_sum + kc0 + tid, (double)t1);
            atomicAdd(a.stat_sq + kc0 + tid, (double)t2);
        }
    }
}

// static modes: mode 2  y[n,t,w,c]      = sum_k sum_u p[n,t,u,k*C+c] * A[k,u,w]   (C = a.KC)
//               mode 3  y[n,t,u,k*C+c]  = sum_w p[n,t,w,c] * A[k,u,w]             (C = a.KC)
template <class T, int V>
__global__ void __launch_bounds__(AG_THREADS) agg_static_kernel(dsg_graph_agg_args a, int t_chunk) {
    DSG_DYN_SMEM(smem_raw);
    float* As = reinterpret_cast<float*>(smem_raw);       // [K][V][V]
    DSG_SHARED float s_red[2][AG_THREADS / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, c0 = blockIdx.y * 32;
    const int tbeg = blockIdx.z * t_chunk;
    const int tend = tbeg + t_chunk < a.T ? tbeg + t_chunk : a.T;
    const int C = a.KC, K = a.Ksub;
    const int c = c0 + lane;
    const bool c_ok = c < C;
    for (int idx = tid; idx < K * V * V; idx += AG_THREADS) As[idx] = a.A[idx];
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int t = tbeg + warp; t < tend; t += AG_THREADS / 32) {
        const long long r0 = ((long long)n * a.T + t) * V;
        if (a.mode == 2) {
            float acc[V];
#pragma unroll
            for (int w = 0; w < V; ++w) acc[w] = 0.f;
            for (int k = 0; k < K; ++k) {
                float p[V];
#pragma unroll
                for (int u = 0; u < V; ++u) p[u] = c_ok ? act_value<T>(a.src, r0 + u, k * C + c) : 0.f;
#pragma unroll
                for (int w = 0; w < V; ++w) {
                    float s = acc[w];
#pragma unroll
                    for (int u = 0; u < V; ++u) s = fmaf(p[u], As[(k * V + u) * V + w], s);
                    acc[w] = s;
                }
            }
            if (c_ok) {
#pragma unroll
                for (int w = 0; w < V; ++w) agg_store<T>(a, r0 + w, c, acc[w], s1, s2);
            }
        } else {
            float p[V];
#pragma unroll
            for (int w = 0; w < V; ++w) p[w] = c_ok ? act_value<T>(a.src, r0 + w, c) : 0.f;
            for (int k = 0; k < K; ++k) {
#pragma unroll
                for (int u = 0; u < V; ++u) {
                    float s = 0.f;
#pragma unroll
                    for (int w = 0; w < V; ++w) s = fmaf(p[w], As[(k * V + u) * V + w], s);
                    // statistics are per output channel k*C+c: not lane-uniform across k, so unsupported here
                    float d1 = 0.f, d2 = 0.f;
                    if (c_ok) agg_store<T>(a, r0 + u, k * C + c, s, d1, d2);
                }
            }
        }
    }
    if (a.stat_sum && a.mode == 2) {
        s_red[0][warp][lane] = s1;
        s_red[1][warp][lane] = s2;
        __syncthreads();
        if (tid < 32 && c0 + tid < C) {
            float t1 = 0.f, t2 = 0.f;
            for (int w = 0; w < AG_THREADS / 32; ++w) { t1 += s_red[0][w][tid]; t2 += s_red[1][w][tid]; }
            atomicAdd(a.stat_sum + c0 + tid, (double)t1);
            atomicAdd(a.stat_sq + c0 + tid, (double)t2);
        }
    }
}

// dadyn[n,u,w,kc] = sum_t p[n,t,u,kc] * dy[n,t,w,kc]: both operands staged per AG_TCH frames (vector loads, fp32,
// channel fastest); a warp owns target joints w, w+8, ..., w+24 and keeps their 25 source-joint accumulators in registers.
constexpr int DA_TCH = AG_TCH;
constexpr int DA_THREADS = 1024;  // 32 warps, one target joint each (25 accumulators per thread)
template <class T, int V, int NT>
__global__ void __launch_bounds__(NT, 1) agg_dadj_dyn_kernel(dsg_graph_agg_dadj_args a, int vec) {
    DSG_DYN_SMEM(smem_raw);
    float* ps = reinterpret_cast<float*>(smem_raw);        // [DA_TCH][V][32]
    float* ds = ps + DA_TCH * V * 32;                      // [DA_TCH][V][32]
    DSG_SHARED float cf_p[96], cf_d[96];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, kc0 = blockIdx.y * 32;
    agg_stage_coefs(a.p, kc0, a.KC, cf_p);
    agg_stage_coefs(a.dy, kc0, a.KC, cf_d);
    constexpr int NW = NT / 32;
    constexpr int WPW = (V + NW - 1) / NW;                 // target joints per warp
    float acc[WPW][V];
#pragma unroll
    for (int i = 0; i < WPW; ++i)
#pragma unroll
        for (int u = 0; u < V; ++u) acc[i][u] = 0.f;
    for (int t0 = 0; t0 < a.T; t0 += DA_TCH) {
        __syncthreads();
        for (int idx = tid; idx < DA_TCH * V * 4; idx += NT) {
            const int q = idx & 3, rv = idx >> 2;
            const int tt = rv / V;
            float pv[8], dv[8];
            if (t0 + tt < a.T) {
                const long long r = ((long long)n * a.T + t0) * V + rv;
                agg_load8_s<T>(a.p, r, kc0, q, a.KC, vec != 0, cf_p, pv);
                agg_load8_s<T>(a.dy, r, kc0, q, a.KC, vec != 0, cf_d, dv);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) pv[j] = dv[j] = 0.f;
            }
            float4* d0 = reinterpret_cast<float4*>(ps + rv * 32 + q * 8);
            d0[0] = make_float4(pv[0], pv[1], pv[2], pv[3]); d0[1] = make_float4(pv[4], pv[5], pv[6], pv[7]);
            float4* d1 = reinterpret_cast<float4*>(ds + rv * 32 + q * 8);
            d1[0] = make_float4(dv[0], dv[1], dv[2], dv[3]); d1[1] = make_float4(dv[4], dv[5], dv[6], dv[7]);
        }
        __syncthreads();
#pragma unroll 1
        for (int tt = 0; tt < DA_TCH; ++tt) {
            float dw[WPW];
#pragma unroll
            for (int i = 0; i < WPW; ++i) {
                const int w = warp + i * NW;
                dw[i] = w < V ? ds[(tt * V + w) * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const float p = ps[(tt * V + u) * 32 + lane];
#pragma unroll
                for (int i = 0; i < WPW; ++i) acc[i][u] = fmaf(p, dw[i], acc[i][u]);
            }
        }
    }
    if (kc0 + lane < a.KC) {
#pragma unroll
        for (int i = 0; i < WPW; ++i) {
            const int w = warp + i * NW;
            if (w < V) {
#pragma unroll
                for (int u = 0; u < V; ++u)
                    a.dadj[(((long long)n * V + u) * V + w) * a.KC + kc0 + lane] = acc[i][u];
            }
        }
    }
}

// static: dA[k,u,w] += sum_{n,t,c} p[n,t,u,k*C+c] * dy[n,t,w,c].   Simple (not hot: only unit_gcn with a
// learnable A uses it): one CTA per (sample, frame chunk), a thread per (k,u,w).
template <class T>
__global__ void __launch_bounds__(AG_THREADS) agg_dadj_static_kernel(dsg_graph_agg_dadj_args a, int t_chunk) {
    const int n = blockIdx.x;
    const int tbeg = blockIdx.y * t_chunk;
    const int tend = tbeg + t_chunk < a.T ? tbeg + t_chunk : a.T;
    const int V = a.V, C = a.KC;
    for (int idx = threadIdx.x; idx < a.Ksub * V * V; idx += AG_THREADS) {
        int w = idx % V, u = (idx / V) % V, k = idx / (V * V);
        float s = 0.f;
        for (int t = tbeg; t < tend; ++t) {
            long long r0 = ((long long)n * a.T + t) * V;
            for (int c = 0; c < C; ++c) s = fmaf(act_value<T>(a.p, r0 + u, k * C + c), act_value<T>(a.dy, r0 + w, c), s);
        }
        atomicAdd(a.dadj + idx, s);
    }
}

template <class T, int V> static const char* launch_agg_v(const dsg_graph_agg_args& a, dsg_stream_t st) {
    if (a.n_samples <= 0) return nullptr;
    // enough CTAs to fill the GPU: split T when the (sample x channel-slice) grid is small
    int slices = (a.KC + 31) / 32;
    long long base = (long long)a.n_samples * slices;
    int tsplit = (int)((2 * dsg_num_sms() + base - 1) / base);
    int unit = (a.mode <= 1) ? AG_TCH : (AG_THREADS / 32);
    int t_chunk = (a.T + tsplit - 1) / tsplit;
    t_chunk = (t_chunk + unit - 1) / unit * unit;
    dim3 grid(a.n_samples, slices, (a.T + t_chunk - 1) / t_chunk);
    if (a.mode <= 1) {
        size_t smem = (size_t)(V * V + AG_TCH * V) * 32 * sizeof(float);
        int vec = (sizeof(T) == 2) && act8_ok(a.src) ? 1 : 0;
        if (a.has_mask || a.stat_sum) {
            // backward-type calls (mask / partner loads per output): 512 threads, one joint per warp pass: 59 registers, so the
            // two CTAs that fit an SM's shared memory bring 32 warps (measured 4.35 -> 3.55 ms per training step)
            DSG_SET_SMEM((agg_dyn_kernel<T, V, AG_DYN_THREADS, 1>), smem);
            dsg_launch((agg_dyn_kernel<T, V, AG_DYN_THREADS, 1>), grid, dim3(AG_DYN_THREADS), smem, st, a, t_chunk, vec);
        } else {
            // plain forward contraction: 256 threads, two joints per warp pass (one shared load feeds two FMAs)
            DSG_SET_SMEM((agg_dyn_kernel<T, V, AG_THREADS, 2>), smem);
            dsg_launch((agg_dyn_kernel<T, V, AG_THREADS, 2>), grid, dim3(AG_THREADS), smem, st, a, t_chunk, vec);
        }
    } else {
        if (a.mode == 3 && a.stat_sum) return "graph_agg: statistics are not supported in mode 3";
        size_t smem = (size_t)a.Ksub * V * V * sizeof(float);
        DSG_SET_SMEM((agg_static_kernel<T, V>), smem);
        dsg_launch(agg_static_kernel<T, V>, grid, dim3(AG_THREADS), smem, st, a, t_chunk);
    }
    return dsg_launch_error();
}

template <class T> static const char* launch_agg(const dsg_graph_agg_args& a, dsg_stream_t st) {
    switch (a.V) {
        case 17: return launch_agg_v<T, 17>(a, st);
        case 18: return launch_agg_v<T, 18>(a, st);
        case 25: return launch_agg_v<T, 25>(a, st);
        default: return "graph_agg: V must be 17 (coco), 18 (openpose) or 25 (nturgb+d)";
    }
}

template <class T, int V> static const char* launch_dadj_v(const dsg_graph_agg_dadj_args& a, dsg_stream_t st) {
    size_t smem = (size_t)2 * DA_TCH * V * 32 * sizeof(float);
    int vec = (sizeof(T) == 2) && act8_ok(a.p) && act8_ok(a.dy) ? 1 : 0;
    DSG_SET_SMEM((agg_dadj_dyn_kernel<T, V, DA_THREADS>), smem);
    dsg_launch((agg_dadj_dyn_kernel<T, V, DA_THREADS>), dim3(a.n_samples, (a.KC + 31) / 32), dim3(DA_THREADS), smem, st, a, vec);
    return dsg_launch_error();
}

template <class T> static const char* launch_dadj(const dsg_graph_agg_dadj_args& a, dsg_stream_t st) {
    if (a.n_samples <= 0) return nullptr;
    if (a.is_static) {
        int t_chunk = 4;
        dsg_launch(agg_dadj_static_kernel<T>, dim3(a.n_samples, (a.T + t_chunk - 1) / t_chunk), dim3(AG_THREADS), 0, st, a, t_chunk);
        return dsg_launch_error();
    }
    switch (a.V) {
        case 17: return launch_dadj_v<T, 17>(a, st);
        case 18: return launch_dadj_v<T, 18>(a, st);
        case 25: return launch_dadj_v<T, 25>(a, st);
        default: return "graph_agg_dadj: V must be 17, 18 or 25";
    }
}

}  // namespace dsg
