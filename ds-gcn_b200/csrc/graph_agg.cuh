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

template <class T, int V>
__global__ void __launch_bounds__(AG_THREADS) agg_dyn_kernel(dsg_graph_agg_args a, int t_chunk) {
    DSG_DYN_SMEM(smem_raw);
    float* adj = reinterpret_cast<float*>(smem_raw);      // [V*V][32]
    DSG_SHARED float s_red[2][AG_THREADS / 32][32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, kc0 = blockIdx.y * 32;
    const int tbeg = blockIdx.z * t_chunk;
    const int tend = tbeg + t_chunk < a.T ? tbeg + t_chunk : a.T;
    const int ch = kc0 + lane;
    const bool ch_ok = ch < a.KC;
    const T* adyn = reinterpret_cast<const T*>(a.adyn) + (long long)n * V * V * a.KC;
    for (int idx = tid; idx < V * V * 32; idx += AG_THREADS) {
        int l = idx & 31, uw = idx >> 5;
        float v = (kc0 + l < a.KC) ? ldf<T>(adyn + (long long)uw * a.KC + kc0 + l) : 0.f;
        int dst = uw;
        if (a.mode == 1) { int u = uw / V, w = uw - u * V; dst = w * V + u; }
        adj[dst * 32 + l] = v;
    }
    __syncthreads();
    float s1 = 0.f, s2 = 0.f;
    for (int tg = tbeg + warp * AG_TF; tg < tend; tg += (AG_THREADS / 32) * AG_TF) {
        float p[AG_TF][V];
#pragma unroll
        for (int f = 0; f < AG_TF; ++f) {
            const bool ok = ch_ok && (tg + f < tend);
            const long long r0 = ((long long)n * a.T + tg + f) * V;
#pragma unroll
            for (int u = 0; u < V; ++u) p[f][u] = ok ? act_value<T>(a.src, r0 + u, ch) : 0.f;
        }
        for (int w = 0; w < V; ++w) {
            float acc[AG_TF];
#pragma unroll
            for (int f = 0; f < AG_TF; ++f) acc[f] = 0.f;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const float av = adj[(u * V + w) * 32 + lane];
#pragma unroll
                for (int f = 0; f < AG_TF; ++f) acc[f] = fmaf(p[f][u], av, acc[f]);
            }
#pragma unroll
            for (int f = 0; f < AG_TF; ++f)
                if (ch_ok && tg + f < tend) agg_store<T>(a, ((long long)n * a.T + tg + f) * V + w, ch, acc[f], s1, s2);
        }
    }
    if (a.stat_sum) {
        s_red[0][warp][lane] = s1;
        s_red[1][warp][lane] = s2;
        __syncthreads();
        if (tid < 32 && kc0 + tid < a.KC) {
            float t1 = 0.f, t2 = 0.f;
            for (int w = 0; w < AG_THREADS / 32; ++w) { t1 += s_red[0][w][tid]; t2 += s_red[1][w][tid]; }
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

// dadyn[n,u,w,kc] = sum_t p[n,t,u,kc] * dy[n,t,w,kc]
constexpr int DA_TCH = 4;
template <class T, int V>
__global__ void __launch_bounds__(AG_THREADS) agg_dadj_dyn_kernel(dsg_graph_agg_dadj_args a) {
    DSG_DYN_SMEM(smem_raw);
    float* ps = reinterpret_cast<float*>(smem_raw);        // [DA_TCH][V][32]
    float* ds = ps + DA_TCH * V * 32;                      // [DA_TCH][V][32]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x, kc0 = blockIdx.y * 32;
    constexpr int NW = AG_THREADS / 32;
    constexpr int UPW = (V + NW - 1) / NW;                 // u's per warp
    float acc[UPW][V];
#pragma unroll
    for (int i = 0; i < UPW; ++i)
#pragma unroll
        for (int w = 0; w < V; ++w) acc[i][w] = 0.f;
    for (int t0 = 0; t0 < a.T; t0 += DA_TCH) {
        __syncthreads();
        for (int idx = tid; idx < DA_TCH * V * 32; idx += AG_THREADS) {
            int l = idx & 31, rv = idx >> 5;              // rv = tt*V + v
            int tt = rv / V;
            float pv = 0.f, dv = 0.f;
            if (t0 + tt < a.T && kc0 + l < a.KC) {
                long long r = ((long long)n * a.T + t0) * V + rv;
                pv = act_value<T>(a.p, r, kc0 + l);
                dv = act_value<T>(a.dy, r, kc0 + l);
            }
            ps[idx] = pv;
            ds[idx] = dv;
        }
        __syncthreads();
#pragma unroll
        for (int tt = 0; tt < DA_TCH; ++tt) {
            float pu[UPW];
#pragma unroll
            for (int i = 0; i < UPW; ++i) {
                int u = warp + i * NW;
                pu[i] = u < V ? ps[(tt * V + u) * 32 + lane] : 0.f;
            }
#pragma unroll
            for (int w = 0; w < V; ++w) {
                const float d = ds[(tt * V + w) * 32 + lane];
#pragma unroll
                for (int i = 0; i < UPW; ++i) acc[i][w] = fmaf(pu[i], d, acc[i][w]);
            }
        }
    }
    if (kc0 + lane < a.KC) {
#pragma unroll
        for (int i = 0; i < UPW; ++i) {
            int u = warp + i * NW;
            if (u < V) {
#pragma unroll
                for (int w = 0; w < V; ++w)
                    a.dadj[(((long long)n * V + u) * V + w) * a.KC + kc0 + lane] = acc[i][w];
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
    int tsplit = (int)((2 * 148 + base - 1) / base);
    int unit = (a.mode <= 1) ? (AG_THREADS / 32) * AG_TF : (AG_THREADS / 32);
    int t_chunk = (a.T + tsplit - 1) / tsplit;
    t_chunk = (t_chunk + unit - 1) / unit * unit;
    dim3 grid(a.n_samples, slices, (a.T + t_chunk - 1) / t_chunk);
    if (a.mode <= 1) {
        size_t smem = (size_t)V * V * 32 * sizeof(float);
        DSG_SET_SMEM((agg_dyn_kernel<T, V>), smem);
        dsg_launch(agg_dyn_kernel<T, V>, grid, dim3(AG_THREADS), smem, st, a, t_chunk);
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
    DSG_SET_SMEM((agg_dadj_dyn_kernel<T, V>), smem);
    dsg_launch(agg_dadj_dyn_kernel<T, V>, dim3(a.n_samples, (a.KC + 31) / 32), dim3(AG_THREADS), smem, st, a);
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
