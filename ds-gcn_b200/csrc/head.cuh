// Classification head of RecognizerGCN in two small kernels per direction (the step after the backbone):
//   forward : logits = pooled @ W^T + b (heads/simple_head.py:93-96), cross-entropy per sample (losses/cross_entropy_loss.py:77-80,
//             F.cross_entropy), top-1 / top-5 hits (core/evaluation top_k_accuracy) — one CTA per HD_SPB samples, a warp per sample
//   backward: dlogits = (softmax - onehot) * gscale, dpooled = dlogits @ W (same CTA layout); dW[k,:] = sum_n dlogits[n,k] pooled[n,:],
//             db[k] = sum_n dlogits[n,k] (one CTA per class, written not accumulated... accumulated: the caller pre-zeroes / owns the sink)
// fp32 throughout (the pooled feature is the fp32 result of the temporal-mean kernel); num_classes <= 1024.
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int HD_THREADS = 256;
constexpr int HD_SPB = HD_THREADS / 32;      // samples per CTA (one warp each)

// per-sample outputs: stats[n*3 + {0,1,2}] = {cross-entropy, top-1 hit, top-5 hit}
__global__ void __launch_bounds__(HD_THREADS) head_ce_fwd_kernel(const float* pooled, const float* W, const float* b, const long long* label,
                                                                 int N, int C, int K, float* logits, float* stats) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * HD_SPB + warp;
    if (n >= N) return;
    const float* x = pooled + (long long)n * C;
    float* lg = logits + (long long)n * K;
    // logits: lane-strided classes, each a C-long dot product (W rows stream from L2; x stays in L1)
    for (int k = lane; k < K; k += 32) {
        const float* w = W + (long long)k * C;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
        int c = 0;
        for (; c + 4 <= C; c += 4) {
            s0 = fmaf(x[c], w[c], s0); s1 = fmaf(x[c + 1], w[c + 1], s1);
            s2 = fmaf(x[c + 2], w[c + 2], s2); s3 = fmaf(x[c + 3], w[c + 3], s3);
        }
        for (; c < C; ++c) s0 = fmaf(x[c], w[c], s0);
        lg[k] = (s0 + s1) + (s2 + s3) + (b ? b[k] : 0.f);
    }
    __syncwarp();
    if (!stats) return;
    const int y = (int)label[n];
    float m = -3.0e38f;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, lg[k]);
    m = warp_max(m);
    float z = 0.f;
    const float ly = (y >= 0 && y < K) ? lg[y] : 0.f;
    int above = 0;                                  // classes scored strictly above the label's (ties resolve like a stable sort: lower index first)
    for (int k = lane; k < K; k += 32) {
        const float v = lg[k];
        z += expf(v - m);
        above += (v > ly || (v == ly && k < y)) ? 1 : 0;
    }
    z = warp_sum(z);
    above = (int)warp_sum((float)above);
    if (lane == 0) {
        stats[n * 3 + 0] = (m + logf(z)) - ly;
        stats[n * 3 + 1] = above < 1 ? 1.f : 0.f;
        stats[n * 3 + 2] = above < (K < 5 ? K : 5) ? 1.f : 0.f;
    }
}

// dlogits (written to `dlogits`, [N,K]) and dpooled [N,C]; gscale = upstream gradient * loss_weight / N read from device memory
__global__ void __launch_bounds__(HD_THREADS) head_ce_bwd_kernel(const float* logits, const long long* label, const float* W, const float* gscale,
                                                                 int N, int C, int K, float* dlogits, float* dpooled) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.x * HD_SPB + warp;
    if (n >= N) return;
    const float* lg = logits + (long long)n * K;
    float* dl = dlogits + (long long)n * K;
    const int y = (int)label[n];
    const float gs = gscale[0];
    float m = -3.0e38f;
    for (int k = lane; k < K; k += 32) m = fmaxf(m, lg[k]);
    m = warp_max(m);
    float z = 0.f;
    for (int k = lane; k < K; k += 32) z += expf(lg[k] - m);
    z = warp_sum(z);
    const float iz = 1.f / z;
    for (int k = lane; k < K; k += 32) dl[k] = (expf(lg[k] - m) * iz - (k == y ? 1.f : 0.f)) * gs;
    __syncwarp();
    if (!dpooled) return;
    for (int c = lane; c < C; c += 32) {            // consecutive lanes = consecutive channels: coalesced reads of W rows
        float s = 0.f;
        for (int k = 0; k < K; ++k) s = fmaf(dl[k], W[(long long)k * C + c], s);
        dpooled[(long long)n * C + c] = s;
    }
}

// dW[k, :] += sum_n dlogits[n,k] * pooled[n,:],  db[k] += sum_n dlogits[n,k]     (one CTA per class, a thread per channel)
__global__ void __launch_bounds__(HD_THREADS) head_wgrad_kernel(const float* dlogits, const float* pooled, int N, int C, int K, float* dW, float* db) {
    const int k = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += HD_THREADS) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s = fmaf(dlogits[(long long)n * K + k], pooled[(long long)n * C + c], s);
        dW[(long long)k * C + c] += s;
    }
    if (threadIdx.x == 0 && db) {
        float s = 0.f;
        for (int n = 0; n < N; ++n) s += dlogits[(long long)n * K + k];
        db[k] += s;
    }
}

static const char* launch_head_ce_fwd(const float* pooled, const float* W, const float* b, const long long* label, int N, int C, int K,
                                      float* logits, float* stats, dsg_stream_t st) {
    if (N <= 0) return nullptr;
    if (C < 1 || K < 1 || !pooled || !W || !logits || (stats && !label)) return "head_ce_fwd: bad arguments";
    dsg_launch(head_ce_fwd_kernel, dim3((N + HD_SPB - 1) / HD_SPB), dim3(HD_THREADS), 0, st, pooled, W, b, label, N, C, K, logits, stats);
    return dsg_launch_error();
}
static const char* launch_head_ce_bwd(const float* logits, const long long* label, const float* pooled, const float* W, const float* gscale,
                                      int N, int C, int K, float* dlogits, float* dpooled, float* dW, float* db, dsg_stream_t st) {
    if (N <= 0) return nullptr;
    if (C < 1 || K < 1 || !logits || !label || !W || !gscale || !dlogits) return "head_ce_bwd: bad arguments";
    dsg_launch(head_ce_bwd_kernel, dim3((N + HD_SPB - 1) / HD_SPB), dim3(HD_THREADS), 0, st, logits, label, W, gscale, N, C, K, dlogits, dpooled);
    if (const char* e = dsg_launch_error()) return e;
    if (dW) {
        if (!pooled) return "head_ce_bwd: the weight gradient needs the pooled feature";
        dsg_launch(head_wgrad_kernel, dim3(K), dim3(HD_THREADS), 0, st, (const float*)dlogits, pooled, N, C, K, dW, db);
    }
    return dsg_launch_error();
}

}  // namespace dsg
