// Channel-wise topology refinement of CTR-GCN (reference gcn.py:634-666 CTRGC, :882-930 unit_ctrgcn) and its backward.
//
//   x1_k[r,v], x2_k[r,v]   : conv1_k / conv2_k of the temporal mean (H[n, v, 6R]: x1_k at k*R, x2_k at 3R + k*R)
//   th_k[r,u,w]            = tanh(x1_k[r,u] - x2_k[r,w])
//   adyn[n,u,w,k*C+c]      = alpha * (sum_r W4_k[c,r] th_k[r,u,w] + b4_k[c]) + A[k,u,w]          (per OUTPUT channel c)
//
// One CTA per (sample stripe, subset k); pairs (u,w) are processed in chunks of CT_PAIRS: the chunk's tanh values are staged in
// shared memory, then a thread per (pair, channel) — channel fastest, so the adjacency rows are written / read coalesced —
// applies the R -> C lift.  Backward accumulates dW4 / db4 / dA / dalpha of its stripe in shared memory and flushes once.
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int CT_THREADS = 256;
constexpr int CT_PAIRS = 32;

struct CtrSmem {
    float *x1, *x2;      // [R][V]
    float* W4;           // [C][R+1]
    float* b4;           // [C]
    float* TH;           // [CT_PAIRS][R]
    // backward only
    float *dx1, *dx2;    // [R][V]
    float* dW4;          // [C][R]
    float* db4;          // [C]
    float* dA;           // [V*V]
    float* DQ;           // [CT_PAIRS][C]
    float* red;          // [1] dalpha
    __device__ CtrSmem(float* base, int R, int V, int C, bool bwd) {
        x1 = base; x2 = x1 + R * V; W4 = x2 + R * V; b4 = W4 + C * (R + 1); TH = b4 + C;
        dx1 = TH + CT_PAIRS * R; dx2 = dx1 + R * V; dW4 = dx2 + R * V; db4 = dW4 + C * R; dA = db4 + C; DQ = dA + V * V;
        red = DQ + CT_PAIRS * C;
        (void)bwd;
    }
    static size_t floats(int R, int V, int C, bool bwd) {
        size_t f = (size_t)2 * R * V + (size_t)C * (R + 1) + C + (size_t)CT_PAIRS * R;
        if (bwd) f += (size_t)2 * R * V + (size_t)C * R + C + (size_t)V * V + (size_t)CT_PAIRS * C + 4;
        return f;
    }
};

DSG_D void ctr_stage(const dsg_ctr_topology_args& a, const CtrSmem& sm, int k) {
    const int R = a.R, C = a.C;
    for (int idx = threadIdx.x; idx < C * R; idx += blockDim.x) {
        const int r = idx % R, c = idx / R;
        sm.W4[c * (R + 1) + r] = a.W4[((long long)k * C + c) * R + r];
    }
    for (int c = threadIdx.x; c < C; c += blockDim.x) sm.b4[c] = a.b4[k * C + c];
}
DSG_D void ctr_load(const dsg_ctr_topology_args& a, const CtrSmem& sm, int n, int k) {
    const int R = a.R, V = a.V;
    const float* h = a.H + (long long)n * V * a.ld_h;
    for (int idx = threadIdx.x; idx < R * V; idx += blockDim.x) {
        const int r = idx % R, v = idx / R;
        sm.x1[r * V + v] = h[(long long)v * a.ld_h + k * R + r];
        sm.x2[r * V + v] = h[(long long)v * a.ld_h + 3 * R + k * R + r];
    }
}
DSG_D void ctr_tanh_chunk(const CtrSmem& sm, int R, int V, int p0, int np) {
    for (int idx = threadIdx.x; idx < np * R; idx += blockDim.x) {
        const int r = idx % R, p = idx / R, uw = p0 + p;
        const int u = uw / V, w = uw - u * V;
        sm.TH[p * R + r] = tanhf(sm.x1[r * V + u] - sm.x2[r * V + w]);
    }
}

template <class T>
__global__ void __launch_bounds__(CT_THREADS) ctr_topology_fwd_kernel(dsg_ctr_topology_args a) {
    DSG_DYN_SMEM(smem_raw);
    const int R = a.R, V = a.V, C = a.C, VV = V * V, KC = 3 * C, k = blockIdx.y;
    CtrSmem sm(reinterpret_cast<float*>(smem_raw), R, V, C, false);
    ctr_stage(a, sm, k);
    const float alpha = a.alpha[0];
    for (int n = blockIdx.x; n < a.n_samples; n += gridDim.x) {
        __syncthreads();
        ctr_load(a, sm, n, k);
        T* out = reinterpret_cast<T*>(a.adyn) + (long long)n * VV * KC + k * C;
        for (int p0 = 0; p0 < VV; p0 += CT_PAIRS) {
            const int np = VV - p0 < CT_PAIRS ? VV - p0 : CT_PAIRS;
            __syncthreads();
            ctr_tanh_chunk(sm, R, V, p0, np);
            __syncthreads();
            for (int idx = threadIdx.x; idx < np * C; idx += blockDim.x) {
                const int c = idx % C, p = idx / C;
                const float* wr = sm.W4 + c * (R + 1);
                const float* th = sm.TH + p * R;
                float s = sm.b4[c];
                for (int r = 0; r < R; ++r) s = fmaf(wr[r], th[r], s);
                stf<T>(out + (long long)(p0 + p) * KC + c, fmaf(alpha, s, a.A[k * VV + p0 + p]));
            }
        }
    }
}

// Backward: g = dadyn[n,u,w,k*C+c] (fp32).  q = W4 th + b4;  dalpha += g q;  dq = alpha g;  dA[k,u,w] += sum_c g;
// dW4[c,r] += dq th[r];  db4[c] += dq;  dth[r] = sum_c W4[c,r] dq[c];  d(arg) = dth (1 - th^2) -> +dx1[r,u], -dx2[r,w].
__global__ void __launch_bounds__(CT_THREADS) ctr_topology_bwd_kernel(dsg_ctr_topology_args a) {
    DSG_DYN_SMEM(smem_raw);
    const int R = a.R, V = a.V, C = a.C, VV = V * V, KC = 3 * C, k = blockIdx.y;
    const int tid = threadIdx.x, NT = blockDim.x;
    CtrSmem sm(reinterpret_cast<float*>(smem_raw), R, V, C, true);
    ctr_stage(a, sm, k);
    for (int idx = tid; idx < C * R + C + VV; idx += NT) sm.dW4[idx] = 0.f;       // dW4, db4, dA are contiguous
    if (tid == 0) sm.red[0] = 0.f;
    const float alpha = a.alpha[0];
    float my_dalpha = 0.f;
    for (int n = blockIdx.x; n < a.n_samples; n += gridDim.x) {
        __syncthreads();
        ctr_load(a, sm, n, k);
        for (int idx = tid; idx < 2 * R * V; idx += NT) sm.dx1[idx] = 0.f;         // dx1, dx2 contiguous
        const float* g = a.dadyn + (long long)n * VV * KC + k * C;
        for (int p0 = 0; p0 < VV; p0 += CT_PAIRS) {
            const int np = VV - p0 < CT_PAIRS ? VV - p0 : CT_PAIRS;
            __syncthreads();
            ctr_tanh_chunk(sm, R, V, p0, np);
            __syncthreads();
            // (a) per (pair, channel): q, dalpha, dq -> DQ; dA via warp-level partial sums
            for (int idx = tid; idx < np * C; idx += NT) {
                const int c = idx % C, p = idx / C;
                const float* wr = sm.W4 + c * (R + 1);
                const float* th = sm.TH + p * R;
                float q = sm.b4[c];
                for (int r = 0; r < R; ++r) q = fmaf(wr[r], th[r], q);
                const float gg = g[(long long)(p0 + p) * KC + c];
                my_dalpha = fmaf(gg, q, my_dalpha);
                sm.DQ[p * C + c] = alpha * gg;
                atomicAdd(&sm.dA[p0 + p], gg);
            }
            __syncthreads();
            // (b) dW4 / db4: the thread that owns (c, r) walks the chunk's pairs (no atomics: fixed ownership)
            for (int idx = tid; idx < C * R; idx += NT) {
                const int r = idx % R, c = idx / R;
                float s = 0.f, sb = 0.f;
                for (int p = 0; p < np; ++p) {
                    const float dq = sm.DQ[p * C + c];
                    s = fmaf(dq, sm.TH[p * R + r], s);
                    sb += dq;
                }
                sm.dW4[idx] += s;
                if (r == 0) sm.db4[c] += sb;
            }
            // (c) dth -> dx1 / dx2
            for (int idx = tid; idx < np * R; idx += NT) {
                const int r = idx % R, p = idx / R, uw = p0 + p;
                const int u = uw / V, w = uw - u * V;
                float s = 0.f;
                for (int c = 0; c < C; ++c) s = fmaf(sm.W4[c * (R + 1) + r], sm.DQ[p * C + c], s);
                const float th = sm.TH[p * R + r];
                const float d = s * (1.f - th * th);
                atomicAdd(&sm.dx1[r * V + u], d);
                atomicAdd(&sm.dx2[r * V + w], -d);
            }
        }
        __syncthreads();
        float* dh = a.dH + (long long)n * V * a.ld_h;
        for (int idx = tid; idx < 2 * R * V; idx += NT) {
            const int r = idx % R, v = (idx / R) % V, which = idx / (R * V);
            const float val = which ? sm.dx2[r * V + v] : sm.dx1[r * V + v];
            const int col = which * 3 * R + k * R + r;
            dh[(long long)v * a.ld_h + col] = val;
            if (a.dH_bf16) reinterpret_cast<bf16*>(a.dH_bf16)[((long long)n * V + v) * a.ld_h + col] = __float2bfloat16(val);
        }
    }
    my_dalpha = warp_sum(my_dalpha);
    if ((tid & 31) == 0) atomicAdd(&sm.red[0], my_dalpha);
    __syncthreads();
    for (int idx = tid; idx < C * R; idx += NT) atomicAdd(a.dW4 + (long long)k * C * R + idx, sm.dW4[idx]);
    for (int idx = tid; idx < C; idx += NT) atomicAdd(a.db4 + k * C + idx, sm.db4[idx]);
    for (int idx = tid; idx < VV; idx += NT) atomicAdd(a.dA + k * VV + idx, sm.dA[idx]);
    if (tid == 0) atomicAdd(a.dalpha, sm.red[0]);
}

static const char* launch_ctr_topology(const dsg_ctr_topology_args& a, bool bwd, dsg_stream_t st) {
    if (a.V > 32 || a.R > 64 || a.R < 1 || a.C < 1) return "ctr_topology: needs V<=32, 1<=R<=64";
    if (a.n_samples <= 0) return nullptr;
    const size_t smem = CtrSmem::floats(a.R, a.V, a.C, bwd) * sizeof(float);
    if (smem > 200 * 1024) return "ctr_topology: shared memory budget exceeded (C*R too large)";
    const int gx = a.n_samples < dsg_num_sms() ? a.n_samples : dsg_num_sms();
    if (!bwd) {
        if (a.adyn_dtype == DSG_BF16) {
            DSG_SET_SMEM(ctr_topology_fwd_kernel<bf16>, smem);
            dsg_launch(ctr_topology_fwd_kernel<bf16>, dim3(gx, 3), dim3(CT_THREADS), smem, st, a);
        } else {
            DSG_SET_SMEM(ctr_topology_fwd_kernel<float>, smem);
            dsg_launch(ctr_topology_fwd_kernel<float>, dim3(gx, 3), dim3(CT_THREADS), smem, st, a);
        }
    } else {
        DSG_SET_SMEM(ctr_topology_bwd_kernel, smem);
        dsg_launch(ctr_topology_bwd_kernel, dim3(gx, 3), dim3(CT_THREADS), smem, st, a);
    }
    return dsg_launch_error();
}

}  // namespace dsg
