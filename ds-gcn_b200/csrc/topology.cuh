// Per-sample dynamic semantic adjacency of dgphgcn1 (reference gcn.py:2239-2337 with the
// north-star flags) and its backward.  One CTA walks samples n = blockIdx.x, += gridDim.x;
// everything for a sample lives in shared memory (V<=32, R<=32).
//
//   x1_k[c,v], x2_k[c,v] (k=0,1: conv1/conv2 halves; k=2: node-type-selected semantic feature, x1_2 == x2_2)
//   arg_0 = x1_0[c,u]-x2_0[c,w];  arg_1 = We[et(u,w)] (x1_1[:,u]-x2_1[:,w]) + be[et(u,w)];  arg_2 = x1_2[c,u]-x1_2[c,w]
//   S_k = softmax_u( sum_c x1_k[c,u] x2_k[c,w] )
//   adyn[u,w,k*R+c] = A[k,u,w] + alpha_k*tanh(arg_k[c,u,w]) + beta_k*S_k[u,w]
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int TP_THREADS = 256;        // forward
constexpr int TP_MAXCNT = 8;           // typed backward path: joints per node type it has room for (NTU: 6, COCO: 5)
constexpr int TP_BWD_THREADS = 1024;   // backward: one CTA per SM (its accumulators fill shared memory), so many warps

struct TopoSmem {
    float* x1;    // [3][R][V]
    float* x2;    // [3][R][V]
    float* S;     // [3][V][V]
    float* G;     // [3][V][V]   fwd: gram; bwd: sum_c dadyn, then dG
    __device__ TopoSmem(float* base, int R, int V) {
        x1 = base;
        x2 = x1 + 3 * R * V;
        S = x2 + 3 * R * V;
        G = S + 3 * V * V;
    }
    __host__ __device__ static size_t floats(int R, int V) { return (size_t)6 * R * V + 6 * V * V; }
};

DSG_D void topo_load_features(const dsg_topology_args& a, const TopoSmem& sm, int n) {
    const int R = a.R, V = a.V;
    const float* h = a.H + (long long)n * V * a.ld_h;
    for (int idx = threadIdx.x; idx < 3 * R * V; idx += blockDim.x) {
        int c = idx % R, v = (idx / R) % V, k = idx / (R * V);
        float f1, f2;
        if (a.variant == 1) {                    // plain DG-GCN unit: H = [conv1 (3R) | conv2 (3R)]
            f1 = h[(long long)v * a.ld_h + k * R + c];
            f2 = h[(long long)v * a.ld_h + 3 * R + k * R + c];
        } else if (k < 2) {
            f1 = h[(long long)v * a.ld_h + k * R + c];
            f2 = h[(long long)v * a.ld_h + 2 * R + k * R + c];
        } else {
            f1 = f2 = h[(long long)v * a.ld_h + 4 * R + c * 5 + a.node_type[v]];
        }
        sm.x1[(k * R + c) * V + v] = f1;
        sm.x2[(k * R + c) * V + v] = f2;
    }
}

// edge_linears weight staged in shared memory with rows padded to R+1 floats: conflict-free both for lanes that walk
// the output channel o (stride R+1) and for lanes that walk the input channel i (stride 1)
DSG_D void topo_stage_we(const dsg_topology_args& a, float* We_s, float* be_s) {
    const int R = a.R;
    for (int idx = threadIdx.x; idx < 15 * R * R; idx += blockDim.x) {
        const int i = idx % R, eo = idx / R;
        We_s[eo * (R + 1) + i] = a.We[idx];
    }
    for (int idx = threadIdx.x; idx < 15 * R; idx += blockDim.x) be_s[idx] = a.be[idx];
}
// subset-1 pre-tanh argument from the staged weights: We[e][o][:] . (x1[1][:,u] - x2[1][:,w]) + be[e][o]
DSG_D float topo_arg1(const TopoSmem& sm, const float* We_s, const float* be_s, int R, int V, int e, int o, int u, int w) {
    const float* wr = We_s + (e * R + o) * (R + 1);
    const float* x1 = sm.x1 + R * V + u;
    const float* x2 = sm.x2 + R * V + w;
    float s = be_s[e * R + o];
    for (int i = 0; i < R; ++i) s = fmaf(wr[i], x1[i * V] - x2[i * V], s);
    return s;
}

// The edge type of a joint pair is a function of the two NODE types (graph.py: rank of the product of the signed type codes), and a
// joint has one of 5 node types — so a source joint u meets at most 5 distinct edge-typed linears, one per node type of the target:
//   We[e(u,w)] (x1[:,u] - x2[:,w]) + be[e(u,w)] = P1[u][type(w)] - P2[w][type(u)]
//   P1[u][t][o] = be[E(type(u),t)][o] + sum_i We[E(type(u),t)][o][i] x1[i,u],   P2[w][t][o] = sum_i We[E(t,type(w))][o][i] x2[i,w]
// = 250 R^2 multiply-adds per sample instead of 625 R^2, and the per-pair work is one subtraction.  topo_type_tables builds
// E[5][5] from one representative joint per type and CHECKS the caller's tables against it (node types in 0..4, every pair's edge
// type equal to E of its node types, in 0..14); a table that is not of that form keeps the per-pair path.
struct TopoTypes { int nt[32]; int rep[8]; int E[25]; int ok; int list[32]; int off[8]; };     // list: joints sorted by type, off[t]..off[t+1]
DSG_D void topo_type_tables(const dsg_topology_args& a, TopoTypes& tt) {
    const int tid = threadIdx.x, V = a.V;
    if (tid < V) tt.nt[tid] = a.node_type[tid];
    if (tid == 0) tt.ok = 1;
    __syncthreads();
    if (tid < V && (tt.nt[tid] < 0 || tt.nt[tid] > 4)) tt.ok = 0;
    if (tid < 5) {
        int r = -1;
        for (int v = V - 1; v >= 0; --v)
            if (tt.nt[v] == tid) r = v;
        tt.rep[tid] = r;
    }
    __syncthreads();
    if (tid < 25) {
        const int ru = tt.rep[tid / 5], rw = tt.rep[tid % 5];
        tt.E[tid] = (tt.ok && ru >= 0 && rw >= 0) ? a.edge_type[ru * V + rw] : 0;
    }
    if (tt.ok) {                                         // counting sort of the joints by node type, a thread per joint / per offset
        if (tid >= 32 && tid < 32 + V) {
            const int v = tid - 32, t = tt.nt[v];
            int pos = 0;
            for (int x = 0; x < V; ++x) pos += (tt.nt[x] < t || (tt.nt[x] == t && x < v)) ? 1 : 0;
            tt.list[pos] = v;
        }
        if (tid >= 64 && tid < 70) {
            const int t = tid - 64;
            int q = 0;
            for (int x = 0; x < V; ++x) q += tt.nt[x] < t ? 1 : 0;
            tt.off[t] = q;
        }

    }
    __syncthreads();
    if (tt.ok)
        for (int idx = tid; idx < V * V; idx += blockDim.x) {
            const int e = a.edge_type[idx];
            if (e < 0 || e > 14 || e != tt.E[tt.nt[idx / V] * 5 + tt.nt[idx % V]]) tt.ok = 0;
        }
    __syncthreads();
}

DSG_D int topo_maxcnt(const TopoTypes& tt) {             // joints of the most frequent node type (valid when tt.ok)
    int mx = 0;
    for (int t = 0; t < 5; ++t) mx = tt.off[t + 1] - tt.off[t] > mx ? tt.off[t + 1] - tt.off[t] : mx;
    return mx;
}

// images of the subset-1 features for the source joints of node type tu: P1t[ul][t][o] (bias included), P2t[w][o]
DSG_D void topo_typed_images(const TopoTypes& tt, const float* We_s, const float* be_s, const float* x1b, const float* x2b, int R, int V, int tu,
                             float* P1t, float* P2t) {
    const int tid = threadIdx.x, NT = blockDim.x;
    const int u0 = tt.off[tu], ntu = tt.off[tu + 1] - u0;
    for (int idx = tid; idx < V * R; idx += NT) {
        const int o = idx % R, w = idx / R;
        const int e = tt.E[tu * 5 + tt.nt[w]];
        const float* wr = We_s + (e * R + o) * (R + 1);
        const float* xr = x2b + w;
        float s0 = 0.f, s1 = 0.f;
        int i = 0;
        for (; i + 2 <= R; i += 2) { s0 = fmaf(wr[i], xr[i * V], s0); s1 = fmaf(wr[i + 1], xr[(i + 1) * V], s1); }
        if (i < R) s0 = fmaf(wr[i], xr[i * V], s0);
        P2t[idx] = s0 + s1;
    }
    for (int idx = tid; idx < ntu * 5 * R; idx += NT) {
        const int o = idx % R, t = (idx / R) % 5, ul = idx / (5 * R);
        const int e = tt.E[tu * 5 + t];
        const float* wr = We_s + (e * R + o) * (R + 1);
        const float* xr = x1b + tt.list[u0 + ul];
        float s0 = be_s[e * R + o], s1 = 0.f;
        int i = 0;
        for (; i + 2 <= R; i += 2) { s0 = fmaf(wr[i], xr[i * V], s0); s1 = fmaf(wr[i + 1], xr[(i + 1) * V], s1); }
        if (i < R) s0 = fmaf(wr[i], xr[i * V], s0);
        P1t[idx] = s0 + s1;
    }
}

// bf16 compute mode: the adjacency is rounded to bf16 (8 mantissa bits) right after, so the hardware tanh (MUFU.TANH, ~2^-11
// relative error, one instruction instead of ~40) is exact enough; the fp32 parity mode keeps tanhf.
template <bool FAST> DSG_D float topo_tanh(float x) {
#ifndef DSG_EMU
    if (FAST) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#endif
    return tanhf(x);
}

template <class T>
__global__ void __launch_bounds__(TP_THREADS) topology_fwd_kernel(dsg_topology_args a) {
    constexpr bool FAST = sizeof(T) == 2;
    DSG_DYN_SMEM(smem_raw);
    const int R = a.R, V = a.V, VV = V * V, KC = 3 * R;
    TopoSmem sm(reinterpret_cast<float*>(smem_raw), R, V);
    float* We_s = reinterpret_cast<float*>(smem_raw) + TopoSmem::floats(R, V);
    float* be_s = We_s + 15 * R * (R + 1);
    const int tid = threadIdx.x;
    const bool plain = a.variant == 1;
    float* P1t = be_s + 15 * R;                          // typed path: [TP_MAXCNT][5][R] images of the source joints of one node type
    float* P2t = P1t + TP_MAXCNT * 5 * R;                //             [V][R] images of the target joints
    DSG_SHARED TopoTypes tt;
    if (!plain) { topo_stage_we(a, We_s, be_s); topo_type_tables(a, tt); }
    // bf16 compute mode only: the fp32 parity mode keeps the reference's association (linear of the difference)
    const bool typed = FAST && !plain && tt.ok && topo_maxcnt(tt) <= TP_MAXCNT;
    const int i1 = a.subset_wise ? 1 : 0, i2 = a.subset_wise ? 2 : 0;    // not subset-wise: alpha[0] / beta[0] scale every subset
    const float al0 = a.alpha[0], al1 = a.alpha[i1], al2 = a.alpha[i2];
    const float be0 = a.beta[0], be1 = a.beta[i1], be2 = a.beta[i2];
    for (int n = blockIdx.x; n < a.n_samples; n += gridDim.x) {
        __syncthreads();
        topo_load_features(a, sm, n);
        __syncthreads();
        for (int idx = tid; idx < 3 * VV; idx += TP_THREADS) {
            int w = idx % V, u = (idx / V) % V, k = idx / VV;
            float s = 0.f;
            for (int c = 0; c < R; ++c) s = fmaf(sm.x1[(k * R + c) * V + u], sm.x2[(k * R + c) * V + w], s);
            sm.G[idx] = s;
        }
        __syncthreads();
        for (int idx = tid; idx < 3 * V; idx += TP_THREADS) {     // column softmax over u
            int w = idx % V, k = idx / V;
            float m = -3.0e38f;
            for (int u = 0; u < V; ++u) m = fmaxf(m, sm.G[(k * V + u) * V + w]);
            float z = 0.f;
            for (int u = 0; u < V; ++u) z += expf(sm.G[(k * V + u) * V + w] - m);
            float iz = 1.f / z;
            for (int u = 0; u < V; ++u) {
                float s = expf(sm.G[(k * V + u) * V + w] - m) * iz;
                sm.S[(k * V + u) * V + w] = s;
                a.S[((long long)n * 3 + k) * VV + u * V + w] = s;
            }
        }
        __syncthreads();
        T* out = reinterpret_cast<T*>(a.adyn) + (long long)n * VV * KC;
        // subsets 0 and 2: tanh of a feature difference; thread = (pair, channel), channel fastest
        const int nd = plain ? 3 : 2;           // subsets whose tanh argument is the plain feature difference
        for (int idx = tid; idx < VV * nd * R; idx += TP_THREADS) {
            const int cc = idx % (nd * R), uw = idx / (nd * R);
            const int kk = cc / R, c = cc - kk * R;
            const int k = plain ? kk : 2 * kk;
            const int u = uw / V, w = uw - u * V;
            const float th = topo_tanh<FAST>(sm.x1[(k * R + c) * V + u] - sm.x2[(k * R + c) * V + w]);
            const float v = a.A[k * VV + uw] + (k == 0 ? al0 : k == 1 ? al1 : al2) * th + (k == 0 ? be0 : k == 1 ? be1 : be2) * sm.S[k * VV + uw];
            stf<T>(out + (long long)uw * KC + k * R + c, v);
        }
        // subset 1: edge-typed linear; consecutive threads = consecutive output channels of one pair
        for (int idx = tid; !plain && !typed && idx < VV * R; idx += TP_THREADS) {
            const int c = idx % R, uw = idx / R;
            const int u = uw / V, w = uw - u * V;
            const float th = topo_tanh<FAST>(topo_arg1(sm, We_s, be_s, R, V, a.edge_type[uw], c, u, w));
            const float v = a.A[VV + uw] + al1 * th + be1 * sm.S[VV + uw];
            stf<T>(out + (long long)uw * KC + R + c, v);
        }
        // typed form (topo_type_tables): per node type of the source joint, the images of the features under the 5 linears that
        // type meets, then one subtraction per (pair, channel)
        for (int tu = 0; typed && tu < 5; ++tu) {
            const int u0 = tt.off[tu], ntu = tt.off[tu + 1] - u0;
            if (ntu == 0) continue;                                 // CTA-uniform
            __syncthreads();                                        // the previous type's readers are done
            topo_typed_images(tt, We_s, be_s, sm.x1 + R * V, sm.x2 + R * V, R, V, tu, P1t, P2t);
            __syncthreads();
            for (int idx = tid; idx < ntu * V * R; idx += TP_THREADS) {
                const int o = idx % R, w = (idx / R) % V, ul = idx / (V * R);
                const int uw = tt.list[u0 + ul] * V + w;
                const float th = topo_tanh<FAST>(P1t[(ul * 5 + tt.nt[w]) * R + o] - P2t[w * R + o]);
                const float v = a.A[VV + uw] + al1 * th + be1 * sm.S[VV + uw];
                stf<T>(out + (long long)uw * KC + R + o, v);
            }
        }
    }
}

// Backward.  Extra shared memory after TopoSmem:
//   dx1,dx2 [3][R][V] each; dA_acc [3][V][V]; dWe_acc [15][R][R]; dbe_acc [15][R]; hbuf [V][R]; dbuf [V][R]; red[8];
//   We_s [15][R][R+1] (row-padded weights); be_s [15][R]
template <bool FAST, int NTB>
__global__ void __launch_bounds__(NTB) topology_bwd_kernel(dsg_topology_args a) {
    DSG_DYN_SMEM(smem_raw);
    const int R = a.R, V = a.V, VV = V * V, KC = 3 * R;
    float* base = reinterpret_cast<float*>(smem_raw);
    TopoSmem sm(base, R, V);
    float* dx1 = base + TopoSmem::floats(R, V);
    float* dx2 = dx1 + 3 * R * V;
    float* dA_acc = dx2 + 3 * R * V;
    float* dWe_acc = dA_acc + 3 * VV;
    float* dbe_acc = dWe_acc + 15 * R * R;
    float* hbuf = dbe_acc + 15 * R;
    float* dbuf = hbuf + V * R;
    float* red = dbuf + V * R;            // [8]: dalpha[3], dbeta[3]
    float* We_s = red + 8;
    float* be_s = We_s + 15 * R * (R + 1);
    const int tid = threadIdx.x, NT = blockDim.x;
    float* d1s = be_s + 15 * R;                                               // [V][R] x1[1][:,u] - x2[1][:,w] of the current source joint
    unsigned char* et_s = reinterpret_cast<unsigned char*>(d1s + V * R);      // [V*V] edge types
    DSG_SHARED unsigned char ord_s[32 * 32];                                  // per source joint u: target joints sorted by edge type
    DSG_SHARED TopoTypes tt;
    float* P1t = d1s + V * R + (VV + 3) / 4;                                  // typed path: [TP_MAXCNT][5][R] images of the source joints of one type
    float* HUt = P1t + TP_MAXCNT * 5 * R;                                     //             [TP_MAXCNT][5][R] sum of h over the targets of one type
    float* P2t = hbuf;                                                        //             [V][R] images of the targets (per source type)
    float* HWt = dbuf;                                                        //             [V][R] sum of h over the sources of the type
    const bool plain = a.variant == 1;
    const int sw = a.subset_wise ? 1 : 0;
    if (!plain) topo_type_tables(a, tt);
    const bool typed = FAST && !plain && tt.ok && topo_maxcnt(tt) <= TP_MAXCNT;      // bf16 compute mode only (the fp32 parity mode keeps the per-pair form)
    if (!plain) {
        topo_stage_we(a, We_s, be_s);
        for (int idx = tid; idx < VV; idx += NT) et_s[idx] = (unsigned char)a.edge_type[idx];
        if (tid < V) {                                                          // counting sort of row `tid` of the edge-type table
            int q = 0;
            for (int e = 0; e < 15; ++e)
                for (int w = 0; w < V; ++w)
                    if (a.edge_type[tid * V + w] == e) ord_s[tid * V + q++] = (unsigned char)w;
            for (int w = 0; w < V; ++w)                                         // (types outside 0..14 never occur; keep the row complete)
                if (a.edge_type[tid * V + w] < 0 || a.edge_type[tid * V + w] > 14) ord_s[tid * V + q++] = (unsigned char)w;
        }
    }
    for (int idx = tid; idx < 3 * VV + 15 * R * R + 15 * R; idx += NT) dA_acc[idx] = 0.f;   // contiguous block
    if (tid < 8) red[tid] = 0.f;

    for (int n = blockIdx.x; n < a.n_samples; n += gridDim.x) {
        __syncthreads();
        topo_load_features(a, sm, n);
        for (int idx = tid; idx < 3 * VV; idx += NT) sm.S[idx] = a.S[(long long)n * 3 * VV + idx];
        for (int idx = tid; idx < 6 * R * V; idx += NT) dx1[idx] = 0.f;   // dx1 and dx2 are contiguous
        const float* g = a.dadyn + (long long)n * VV * KC;
        __syncthreads();
        // (1) sS[k,u,w] = sum_c g[u,w,kR+c]  -> G ; dA accumulates it.  Thread per (pair, subset): R/4 independent
        //     16-byte loads in flight
        const bool g_vec = (R % 4 == 0) && ((uintptr_t)g % 16 == 0);
        for (int idx = tid; idx < 3 * VV; idx += NT) {
            const int k = idx % 3, uw = idx / 3;
            const float* gp = g + (long long)uw * KC + k * R;
            float s = 0.f;
            if (g_vec) {
                for (int c = 0; c < R; c += 4) {
                    const float4 q = *reinterpret_cast<const float4*>(gp + c);
                    s += (q.x + q.y) + (q.z + q.w);
                }
            } else {
                for (int c = 0; c < R; ++c) s += gp[c];
            }
            sm.G[k * VV + uw] = s;
            dA_acc[k * VV + uw] += s;
        }
        __syncthreads();
        // (2) softmax backward per column (k,w): dG = beta*S*(sS - sum_u sS*S); dbeta += sum sS*S
        for (int idx = tid; idx < 3 * V; idx += NT) {
            int w = idx % V, k = idx / V;
            float dot = 0.f;
            for (int u = 0; u < V; ++u) dot = fmaf(sm.G[(k * V + u) * V + w], sm.S[(k * V + u) * V + w], dot);
            atomicAdd(&red[3 + k], dot);
            const float bk = a.beta[sw * k];
            for (int u = 0; u < V; ++u) {
                int i = (k * V + u) * V + w;
                sm.G[i] = bk * sm.S[i] * (sm.G[i] - dot);
            }
        }
        __syncthreads();
        // (3) gram backward + subsets 0 and 2 of the tanh branch; thread per (k,c,v), no atomics:
        //     dx1[k,c,u] = sum_w dG[u,w] x2[c,w] + sum_w h[c,u,w];  dx2[k,c,w] = sum_u dG[u,w] x1[c,u] - sum_u h[c,u,w]
        float my_dalpha0 = 0.f, my_dalpha2 = 0.f, my_dalpha1 = 0.f;
        for (int idx = tid; idx < 3 * R * V; idx += NT) {
            int c = idx % R, v = (idx / R) % V, k = idx / (R * V);
            float s1 = 0.f, s2 = 0.f;
            const float* x1r = sm.x1 + (k * R + c) * V;
            const float* x2r = sm.x2 + (k * R + c) * V;
            for (int o = 0; o < V; ++o) {
                s1 = fmaf(sm.G[(k * V + v) * V + o], x2r[o], s1);      // u=v, w=o
                s2 = fmaf(sm.G[(k * V + o) * V + v], x1r[o], s2);      // u=o, w=v
            }
            if (k != 1 || plain) {
                const float al = a.alpha[sw * k];
                float da = 0.f;
#pragma unroll 5
                for (int o = 0; o < V; ++o) {
                    float th = topo_tanh<FAST>(x1r[v] - x2r[o]);                  // pair (u=v, w=o)
                    float gg = g[(long long)(v * V + o) * KC + k * R + c];
                    da = fmaf(gg, th, da);
                    s1 = fmaf(al * (1.f - th * th), gg, s1);
                    float th2 = topo_tanh<FAST>(x1r[o] - x2r[v]);                 // pair (u=o, w=v)
                    float gg2 = g[(long long)(o * V + v) * KC + k * R + c];
                    s2 = fmaf(-al * (1.f - th2 * th2), gg2, s2);
                }
                if (k == 0) my_dalpha0 += da; else if (k == 1) my_dalpha1 += da; else my_dalpha2 += da;
            }
            dx1[(k * R + c) * V + v] = s1;
            dx2[(k * R + c) * V + v] = s2;
        }
        my_dalpha0 = warp_sum(my_dalpha0);
        my_dalpha2 = warp_sum(my_dalpha2);
        if ((tid & 31) == 0) { atomicAdd(&red[0], my_dalpha0); atomicAdd(&red[2], my_dalpha2); }
        __syncthreads();
        // (4) subset 1 (edge-typed linear), one source joint u at a time, two barriers per joint:
        //     (a) h[w][o] = alpha1 * (1 - tanh^2) * g   (+ the dx1 column of the previous joint, from dbuf)
        //     (b) dWe[e][o][i] += h[w][o] * d1[w][i], dbe[e][o] += h[w][o]   (thread owns (o,i): no atomics)
        //         dd1[w][i] = sum_o We[e][o][i] h[w][o] -> dx2[1][i][w] -= dd1, dbuf = dd1
        //     with d1[w][i] = x1[1][i][u] - x2[1][i][w] formed on the fly
        const float al1 = a.alpha[sw];
        const float* x1b = sm.x1 + R * V;
        const float* x2b = sm.x2 + R * V;
        if (typed) {
            // (4t) subset 1 by node type of the source joint (topo_type_tables): three barriers per type instead of two per joint, every
            //      phase V*R or more items wide, 750 R^2 multiply-adds per sample instead of 1875 R^2
            for (int tu = 0; tu < 5; ++tu) {
                const int u0 = tt.off[tu], ntu = tt.off[tu + 1] - u0;
                if (ntu == 0) continue;                                 // CTA-uniform
                // (a) images of the features under the 5 linears this source type meets
                topo_typed_images(tt, We_s, be_s, x1b, x2b, R, V, tu, P1t, P2t);
                __syncthreads();
                // (b) h = alpha1 (1 - tanh^2) g per (source joint of the type, target joint, channel), summed over the targets of each
                //     node type (HUt) and over the sources of this type (HWt): two passes that recompute h, fixed summation order
                for (int idx = tid; idx < ntu * 5 * R; idx += NT) {
                    const int o = idx % R, t = (idx / R) % 5, ul = idx / (5 * R);
                    const int u = tt.list[u0 + ul];
                    const float p1 = P1t[idx];
                    float hs = 0.f;
                    for (int q = tt.off[t]; q < tt.off[t + 1]; ++q) {
                        const int w = tt.list[q];
                        const float gg = g[(long long)(u * V + w) * KC + R + o];
                        const float th = topo_tanh<FAST>(p1 - P2t[w * R + o]);
                        my_dalpha1 = fmaf(gg, th, my_dalpha1);
                        hs = fmaf(al1 * (1.f - th * th), gg, hs);
                    }
                    HUt[idx] = hs;
                }
                for (int idx = tid; idx < V * R; idx += NT) {
                    const int o = idx % R, w = idx / R, tw = tt.nt[w];
                    const float p2 = P2t[idx];
                    float hs = 0.f;
                    for (int ul = 0; ul < ntu; ++ul) {
                        const float gg = g[(long long)(tt.list[u0 + ul] * V + w) * KC + R + o];
                        const float th = topo_tanh<FAST>(P1t[(ul * 5 + tw) * R + o] - p2);
                        hs = fmaf(al1 * (1.f - th * th), gg, hs);
                    }
                    HWt[idx] = hs;
                }
                __syncthreads();
                // (c) every consumer of the sums writes its own accumulators
                for (int idx = tid; idx < V * R; idx += NT) {           // dx2[1][i][w] -= sum_o We[E(tu,type w)][o][i] HWt[w][o]
                    const int i = idx % R, w = idx / R;
                    const float* wc = We_s + tt.E[tu * 5 + tt.nt[w]] * R * (R + 1) + i;
                    const float* hr = HWt + w * R;
                    float s0 = 0.f, s1 = 0.f;
                    int o = 0;
                    for (; o + 2 <= R; o += 2) { s0 = fmaf(wc[o * (R + 1)], hr[o], s0); s1 = fmaf(wc[(o + 1) * (R + 1)], hr[o + 1], s1); }
                    if (o < R) s0 = fmaf(wc[o * (R + 1)], hr[o], s0);
                    dx2[(R + i) * V + w] -= s0 + s1;
                }
                for (int idx = tid; idx < ntu * R; idx += NT) {         // dx1[1][i][u] += sum_t sum_o We[E(tu,t)][o][i] HUt[u][t][o]
                    const int i = idx % R, ul = idx / R;
                    float s0 = 0.f;
                    for (int t = 0; t < 5; ++t) {
                        const float* wc = We_s + tt.E[tu * 5 + t] * R * (R + 1) + i;
                        const float* hr = HUt + (ul * 5 + t) * R;
                        for (int o = 0; o < R; ++o) s0 = fmaf(wc[o * (R + 1)], hr[o], s0);
                    }
                    dx1[(R + i) * V + tt.list[u0 + ul]] += s0;
                }
                for (int idx = tid; idx < 5 * R * R; idx += NT) {       // dWe, dbe of the 5 edge types {tu, t}
                    const int i = idx % R, o = (idx / R) % R, t = idx / (R * R);
                    if (tt.off[t + 1] == tt.off[t]) continue;
                    const int e = tt.E[tu * 5 + t];
                    float acc = 0.f, hs = 0.f;
                    for (int ul = 0; ul < ntu; ++ul) {
                        const float h = HUt[(ul * 5 + t) * R + o];
                        acc = fmaf(h, x1b[i * V + tt.list[u0 + ul]], acc);
                        hs += h;
                    }
                    for (int q = tt.off[t]; q < tt.off[t + 1]; ++q) {
                        const int w = tt.list[q];
                        acc = fmaf(-HWt[w * R + o], x2b[i * V + w], acc);
                    }
                    dWe_acc[(e * R + o) * R + i] += acc;
                    if (i == 0) dbe_acc[e * R + o] += hs;
                }
                __syncthreads();
            }
        }
        if (!plain && !typed) {                                         // d1 of the first source joint (later ones are formed during (b))
            for (int idx = tid; idx < V * R; idx += NT) {
                const int i = idx % R, w = idx / R;
                d1s[idx] = x1b[i * V] - x2b[i * V + w];
            }
            __syncthreads();
        }
        for (int u = 0; !plain && !typed && u <= V; ++u) {
            if (u > 0 && tid < R) {                                     // dx1[1][i][u-1] += sum_w dd1[w][i]
                float sacc = 0.f;
                for (int ww = 0; ww < V; ++ww) sacc += dbuf[ww * R + tid];
                dx1[(R + tid) * V + u - 1] += sacc;
            }
            if (u == V) break;
            for (int idx = tid; idx < V * R; idx += NT) {       // (a): two shared loads + FMA per term (row-padded weights, staged d1)
                const int o = idx % R, w = idx / R;
                const int e = (int)et_s[u * V + w];
                const float* wr = We_s + (e * R + o) * (R + 1);
                const float* dr = d1s + w * R;
                const float gg = g[(long long)(u * V + w) * KC + R + o];       // issued first: its latency hides behind the dot product
                float a0 = be_s[e * R + o], a1 = 0.f, a2 = 0.f, a3 = 0.f;      // four independent chains (the loop was latency-bound)
                int i = 0;
                for (; i + 4 <= R; i += 4) {
                    a0 = fmaf(wr[i], dr[i], a0);
                    a1 = fmaf(wr[i + 1], dr[i + 1], a1);
                    a2 = fmaf(wr[i + 2], dr[i + 2], a2);
                    a3 = fmaf(wr[i + 3], dr[i + 3], a3);
                }
                for (; i < R; ++i) a0 = fmaf(wr[i], dr[i], a0);
                const float th = topo_tanh<FAST>((a0 + a1) + (a2 + a3));
                my_dalpha1 = fmaf(gg, th, my_dalpha1);
                hbuf[idx] = al1 * (1.f - th * th) * gg;
            }
            __syncthreads();
            // (b) dWe, dbe: the target joints are walked in edge-type order (few distinct types per source joint) with a register
            //     accumulator and ONE read-modify-write of the shared accumulator per type (was one per target joint)
            for (int idx = tid; idx < R * R; idx += NT) {
                const int i = idx % R, o = idx / R;
                const float x1u = x1b[i * V + u];
                const unsigned char* ord = ord_s + u * V;
                int e_cur = (int)et_s[u * V + ord[0]];
                float acc = 0.f, hs = 0.f;
                for (int q = 0; q < V; ++q) {
                    const int w = ord[q];
                    const int e = (int)et_s[u * V + w];
                    if (e != e_cur) {
                        dWe_acc[(e_cur * R + o) * R + i] += acc;
                        if (i == 0) dbe_acc[e_cur * R + o] += hs;
                        acc = hs = 0.f;
                        e_cur = e;
                    }
                    const float h = hbuf[w * R + o];
                    acc = fmaf(h, x1u - x2b[i * V + w], acc);
                    hs += h;
                }
                dWe_acc[(e_cur * R + o) * R + i] += acc;
                if (i == 0) dbe_acc[e_cur * R + o] += hs;
            }
            for (int idx = tid; idx < V * R; idx += NT) {       // (b) dd1, and d1 of the next source joint
                const int i = idx % R, w = idx / R;
                const int e = (int)et_s[u * V + w];
                const float* wc = We_s + e * R * (R + 1) + i;
                float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
                const float* hr = hbuf + w * R;
                int o = 0;
                for (; o + 4 <= R; o += 4) {
                    s0 = fmaf(wc[o * (R + 1)], hr[o], s0);
                    s1 = fmaf(wc[(o + 1) * (R + 1)], hr[o + 1], s1);
                    s2 = fmaf(wc[(o + 2) * (R + 1)], hr[o + 2], s2);
                    s3 = fmaf(wc[(o + 3) * (R + 1)], hr[o + 3], s3);
                }
                for (; o < R; ++o) s0 = fmaf(wc[o * (R + 1)], hr[o], s0);
                const float sacc = (s0 + s1) + (s2 + s3);
                dbuf[idx] = sacc;
                dx2[(R + i) * V + w] -= sacc;
                if (u + 1 < V) d1s[idx] = x1b[i * V + u + 1] - x2b[i * V + w];
            }
            __syncthreads();
        }
        my_dalpha1 = warp_sum(my_dalpha1);
        if ((tid & 31) == 0) atomicAdd(&red[1], my_dalpha1);
        __syncthreads();
        // (5) write dH[n][v][9R]  (plain variant: [v][6R] = dx1 | dx2)
        float* dh = a.dH + (long long)n * V * a.ld_h;
        const int HC = plain ? 6 * R : 9 * R;
        for (int idx = tid; idx < V * HC; idx += NT) {
            int col = idx % HC, v = idx / HC;
            float val;
            if (plain) val = col < 3 * R ? dx1[col * V + v] : dx2[(col - 3 * R) * V + v];
            else if (col < 2 * R) val = dx1[col * V + v];
            else if (col < 4 * R) val = dx2[(col - 2 * R) * V + v];
            else {
                int q = col - 4 * R, c = q / 5, ty = q - c * 5;
                val = (ty == a.node_type[v]) ? dx1[(2 * R + c) * V + v] + dx2[(2 * R + c) * V + v] : 0.f;
            }
            dh[(long long)v * a.ld_h + col] = val;
            if (a.dH_bf16) reinterpret_cast<bf16*>(a.dH_bf16)[((long long)n * V + v) * a.ld_h + col] = __float2bfloat16(val);
        }
    }
    __syncthreads();
    for (int idx = tid; idx < 3 * VV; idx += NT) atomicAdd(a.dA + idx, dA_acc[idx]);
    if (!plain) {
        for (int idx = tid; idx < 15 * R * R; idx += NT) atomicAdd(a.dWe + idx, dWe_acc[idx]);
        for (int idx = tid; idx < 15 * R; idx += NT) atomicAdd(a.dbe + idx, dbe_acc[idx]);
    }
    if (tid < 3) { atomicAdd(a.dalpha + sw * tid, red[tid]); atomicAdd(a.dbeta + sw * tid, red[3 + tid]); }
}

static inline size_t topo_bwd_smem_floats(int R, int V) {
    return TopoSmem::floats(R, V) + (size_t)6 * R * V + 3 * V * V + 15 * R * R + 15 * R + 2 * V * R + 8 + 15 * R * (R + 1) + 15 * R + V * R + (V * V + 3) / 4 +
           (size_t)2 * TP_MAXCNT * 5 * R;
}

static const char* launch_topology(const dsg_topology_args& a, bool bwd, dsg_stream_t st) {
    if (a.V > 32 || a.R > 32 || a.R < 1) return "topology: needs V<=32 and 1<=R<=32";
    if (a.n_samples <= 0) return nullptr;
    int grid = a.n_samples < 2 * dsg_num_sms() ? a.n_samples : 2 * dsg_num_sms();
    if (!bwd) {
        size_t smem = (TopoSmem::floats(a.R, a.V) + (size_t)15 * a.R * (a.R + 1) + 15 * a.R + (size_t)(TP_MAXCNT * 5 + a.V) * a.R) * sizeof(float);
        if (a.adyn_dtype == DSG_BF16) {
            DSG_SET_SMEM(topology_fwd_kernel<bf16>, smem);
            dsg_launch(topology_fwd_kernel<bf16>, dim3(grid), dim3(TP_THREADS), smem, st, a);
        } else {
            DSG_SET_SMEM(topology_fwd_kernel<float>, smem);
            dsg_launch(topology_fwd_kernel<float>, dim3(grid), dim3(TP_THREADS), smem, st, a);
        }
    } else {
        size_t smem = topo_bwd_smem_floats(a.R, a.V) * sizeof(float);
        if (smem > 224 * 1024) return "topology_bwd: shared memory budget exceeded";
        // the kernel is bound by the latency of its dependent global / shared round trips, not by issue slots (ncu: 50 % warps
        // active, 2 % DRAM): where the accumulators leave room (R <= 16), two 512-thread CTAs share an SM and every sample gets
        // its own CTA, so one CTA's barriers and load latencies hide behind the other's work
        const bool small = smem <= 100 * 1024;
        const int cap = small ? 2 * dsg_num_sms() : dsg_num_sms();
        grid = a.n_samples < cap ? a.n_samples : cap;
#define DSG_TOPO_BWD(FAST_, NT_)                                                                         \
        do {                                                                                             \
            DSG_SET_SMEM((topology_bwd_kernel<FAST_, NT_>), smem);                                       \
            dsg_launch((topology_bwd_kernel<FAST_, NT_>), dim3(grid), dim3(NT_), smem, st, a);           \
        } while (0)
        // bf16 compute mode (the caller asks for the bf16 copy of dH exactly then): same tanh as the forward kernel used
        if (a.dH_bf16) { if (small) DSG_TOPO_BWD(true, 512); else DSG_TOPO_BWD(true, TP_BWD_THREADS); }
        else { if (small) DSG_TOPO_BWD(false, 512); else DSG_TOPO_BWD(false, TP_BWD_THREADS); }
#undef DSG_TOPO_BWD
    }
    return dsg_launch_error();
}

}  // namespace dsg
