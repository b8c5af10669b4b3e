// Row-per-thread tcgen05 engine for the hot 1x1 convolutions (taps == 1, everything 16-byte aligned).
//
// 128 threads = the 128 rows of the M tile = the 128 TMEM lanes.  A thread loads *its own* row (K/8 independent
// 16-byte loads in flight), applies the fused prologue and stores bf16 chunks into the K-major no-swizzle UMMA tile;
// after the MMAs it reads *its own* accumulator row from TMEM (tcgen05.ld 32x32b: lane == row) and runs the fused tail
// on registers: no fp32 staging tile, no per-element index arithmetic, 24-64 KB of shared memory per CTA so several
// CTAs overlap on one SM.  BatchNorm statistics use a transpose-reduce over the warp (16 shuffles per 16 columns).
// The reduction axis is walked in passes of 128 channels.  Shapes this kernel does not take (temporal taps, odd
// widths, unaligned slices) run on conv_gemm_tc_kernel.
#pragma once
#include "conv_gemm_tc.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc {

constexpr int T2_THREADS = 128;
constexpr int T2_KPASS = 128;
constexpr int T2_BN = 128;

// per-column sums over the 32 lanes of a warp for 16 columns held by every lane: returns, in lane l, the total of
// column (l >> 1) (both lanes of a pair hold it)
DSG_D float warp_colsum16(const float* v, int lane) {
    float w8[8], w4[4], w2[2], w1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float send = (lane & 16) ? v[i] : v[i + 8], keep = (lane & 16) ? v[i + 8] : v[i];
        w8[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float send = (lane & 8) ? w8[i] : w8[i + 4], keep = (lane & 8) ? w8[i + 4] : w8[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const float send = (lane & 4) ? w4[i] : w4[i + 2], keep = (lane & 4) ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    {
        const float send = (lane & 2) ? w2[0] : w2[1], keep = (lane & 2) ? w2[1] : w2[0];
        w1 = keep + __shfl_xor_sync(0xffffffffu, send, 2);
    }
    return w1 + __shfl_xor_sync(0xffffffffu, w1, 1);
}

template <int NB, bool DUAL>
DSG_D void t2_stage_a(const dsg_conv_gemm_args& a, long long sr, int tid, int kv0, int nch, unsigned char* Abase,
                      const float* cf_a1, const float* cf_b, const float* cf_a2) {
    const bf16* x1 = reinterpret_cast<const bf16*>(a.src.x1) + sr * a.src.ld1;
    const bf16* x2 = DUAL ? reinterpret_cast<const bf16*>(a.src.x2) + sr * a.src.ld2 : nullptr;
    for (int kc0 = 0; kc0 < nch; kc0 += NB) {
        Act8Raw raw[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int k = kv0 + (kc0 + b) * 8;
            raw[b].a = raw[b].b = make_uint4(0u, 0u, 0u, 0u);
            if (kc0 + b < nch && sr >= 0 && k < a.K) {
                raw[b].a = *reinterpret_cast<const uint4*>(x1 + k);
                if (DUAL) raw[b].b = *reinterpret_cast<const uint4*>(x2 + k);
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int kc = kc0 + b, k = kv0 + kc * 8;
            if (kc >= nch) continue;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (sr >= 0 && k < a.K) {
                float v[8];
                finish_smem(raw[b], DUAL, a.src.relu, cf_a1 + k, cf_b + k, cf_a2 + k, v);
                pk = pack8(v);
            }
            *reinterpret_cast<uint4*>(Abase + op_off(tid, kc, nch)) = pk;
        }
    }
}

// ---- pre-packed weights: bf16 tiles in the K-major no-swizzle UMMA layout, one per (128-column tile j, K pass p):
//      tile (j, p) starts at byte j * 128 * Kp * 2 + 128 * (p * T2_KPASS) * 2 and holds Ntp(j) * kv_len(p) * 2 bytes
static inline long long conv_wpack_bytes(int K, int N) {
    const long long Kp = (K + 15) & ~15, ntiles = (N + T2_BN - 1) / T2_BN;
    return ntiles * 128LL * Kp * 2;
}
__global__ void __launch_bounds__(256) conv_wpack_kernel(const float* W, long long ws_n, long long ws_k, int K, int N, unsigned char* out) {
    const int Kp = (K + 15) & ~15;
    const int kpass = Kp < T2_KPASS ? Kp : T2_KPASS;
    const int n0 = blockIdx.x * T2_BN, kv0 = blockIdx.y * kpass;
    const int Nt = N - n0 < T2_BN ? N - n0 : T2_BN, Ntp = (Nt + 15) & ~15;
    const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass, nch = kv_len >> 3;
    unsigned char* dst = out + (long long)blockIdx.x * 128 * Kp * 2 + 128LL * kv0 * 2;
    for (int idx = threadIdx.x; idx < Ntp * nch; idx += 256) {
        int n, kc;
        if (ws_k == 1) { kc = idx % nch; n = idx / nch; } else { n = idx % Ntp; kc = idx / Ntp; }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kv0 + kc * 8 + e;
            v[e] = (n < Nt && k < K) ? W[(long long)(n0 + n) * ws_n + (long long)k * ws_k] : 0.f;
        }
        *reinterpret_cast<uint4*>(dst + op_off(n, kc, nch)) = pack8(v);
    }
}
DSG_D void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DSG_D void bulk_g2s(void* dst_smem, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst_smem)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// wmode 0: W[n][k] rows contiguous in k; 1: contiguous in n (MN-major B); 3: pre-packed tiles in a.wpack (bulk copy)
__global__ void __launch_bounds__(T2_THREADS) conv_gemm_tc2_kernel(dsg_conv_gemm_args a, int wmode) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar, wbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float ext_s[2][8][16];                    // joint-mean accumulator rows of up to 8 frames, double-buffered
    __shared__ float s_acc[2][4][T2_BN];                 // per-warp column sums
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rpf = a.Vin + a.ext_in;
    const int Fr = 128 / rpf;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const long long f0 = (long long)blockIdx.x * Fr;
    const int n0 = blockIdx.y * T2_BN;
    const int Nt = a.N - n0 < T2_BN ? a.N - n0 : T2_BN;
    const int Ntp = (Nt + 15) & ~15;
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Ntp) tmem_cols <<= 1;
    const int Kp = (a.K + 15) & ~15;
    const int kpass = Kp < T2_KPASS ? Kp : T2_KPASS;
    unsigned char* Abase = smem;
    unsigned char* Bbase = smem + (size_t)128 * kpass * 2;
    // per-channel coefficients staged once per CTA (defaults 1/0 filled in), so the per-chunk prologue never waits on
    // a dependent global load: src: a1,b1(+b2),a2 over Kp channels; tail: bias, mask a1,b1(+b2),a2 over the 128-column tile
    float* cf_a1 = reinterpret_cast<float*>(smem + (size_t)(128 + T2_BN) * kpass * 2);
    float* cf_b = cf_a1 + Kp;
    float* cf_a2 = cf_b + Kp;
    float* tl_bias = cf_a2 + Kp;
    float* tl_ma1 = tl_bias + T2_BN;
    float* tl_mb = tl_ma1 + T2_BN;
    float* tl_ma2 = tl_mb + T2_BN;
    for (int k = tid; k < Kp; k += T2_THREADS) {
        const bool in = k < a.K;
        cf_a1[k] = (in && a.src.a1) ? a.src.a1[k] : 1.f;
        cf_b[k] = ((in && a.src.b1) ? a.src.b1[k] : 0.f) + ((in && a.src.b2) ? a.src.b2[k] : 0.f);
        cf_a2[k] = (in && a.src.a2) ? a.src.a2[k] : 1.f;
    }
    {
        const int cch = n0 + tid;
        const bool in = cch < a.N;
        tl_bias[tid] = (in && a.bias) ? a.bias[cch] : 0.f;
        tl_ma1[tid] = (in && a.has_mask && a.mask.a1) ? a.mask.a1[cch] : 1.f;
        tl_mb[tid] = ((in && a.has_mask && a.mask.b1) ? a.mask.b1[cch] : 0.f) + ((in && a.has_mask && a.mask.b2) ? a.mask.b2[cch] : 0.f);
        tl_ma2[tid] = (in && a.has_mask && a.mask.a2) ? a.mask.a2[cch] : 1.f;
    }

    // ---- this thread's row
    const int fl = tid / rpf, j = tid - fl * rpf;
    const long long f = f0 + fl;
    const bool row_ok = fl < Fr && f < n_frames;
    long long sr = -1;                                   // source row (-2: joint-mean row)
    if (row_ok) {
        FrameMap fm{1, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};
        const long long sf = src_frame(fm, f, 0);
        if (sf >= 0) sr = (j < a.Vin) ? sf * a.Vin + j : -2;
    }
    const int Vout = rpf - a.contract_ext;
    const bool is_ext_row = a.contract_ext && j == rpf - 1;
    const long long orow = (row_ok && !is_ext_row) ? f * Vout + j : -1;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        mbar_init(&wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    for (int i = tid; i < 2 * 4 * T2_BN; i += T2_THREADS) (&s_acc[0][0][0])[i] = 0.f;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = wmode == 1 ? make_idesc_bmn(128, Ntp) : make_idesc(128, Ntp);

    uint32_t phase = 0, wphase = 0;
    int first = 1;
    for (int kv0 = 0; kv0 < Kp; kv0 += kpass) {
        const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass;
        const int nch = kv_len >> 3;
        if (!first) mbar_wait(&mbar, phase ^ 1);
        if (wmode == 3 && tid == 0) {                        // this pass's weight tile: one bulk copy, overlapped with the A staging
            const uint32_t bytes = (uint32_t)(Ntp * kv_len * 2);
            mbar_expect_tx(&wbar, bytes);
            bulk_g2s(Bbase, reinterpret_cast<const unsigned char*>(a.wpack) + (long long)blockIdx.y * 128 * Kp * 2 + 128LL * kv0 * 2, bytes, &wbar);
        }
        // ---- A: my row; NB independent 16-byte loads in flight per source (8 with one source, 4 + 4 with two)
        if (a.src.x2 == nullptr) t2_stage_a<8, false>(a, sr, tid, kv0, nch, Abase, cf_a1, cf_b, cf_a2);
        else t2_stage_a<4, true>(a, sr, tid, kv0, nch, Abase, cf_a1, cf_b, cf_a2);
        // ---- B: weights fp32 -> bf16
        if (wmode == 0) {
            if (tid < Ntp) {
                const float* wrow = a.W + (long long)(n0 + tid) * a.ws_n;
                for (int kc0 = 0; kc0 < nch; kc0 += 4) {
                    float4 x[4], y[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int k = kv0 + (kc0 + b) * 8;
                        if (kc0 + b < nch && tid < Nt && k < a.K) {
                            x[b] = *reinterpret_cast<const float4*>(wrow + k);
                            y[b] = *reinterpret_cast<const float4*>(wrow + k + 4);
                        } else x[b] = y[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (kc0 + b >= nch) continue;
                        float w[8] = {x[b].x, x[b].y, x[b].z, x[b].w, y[b].x, y[b].y, y[b].z, y[b].w};
                        *reinterpret_cast<uint4*>(Bbase + op_off(tid, kc0 + b, nch)) = pack8(w);
                    }
                }
            }
        } else if (wmode == 1) {                          // MN-major: my k row, all 8-channel groups
            const int gN = Ntp >> 3;
            if (tid < kv_len) {
                const int k = kv0 + tid;
                const float* wrow = a.W + (long long)k * a.ws_k + n0;
                for (int g0 = 0; g0 < gN; g0 += 4) {
                    float4 x[4], y[4];
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const int n = (g0 + b) * 8;
                        if (g0 + b < gN && k < a.K && n < Nt) {
                            x[b] = *reinterpret_cast<const float4*>(wrow + n);
                            y[b] = *reinterpret_cast<const float4*>(wrow + n + 4);
                        } else x[b] = y[b] = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        if (g0 + b >= gN) continue;
                        float w[8] = {x[b].x, x[b].y, x[b].z, x[b].w, y[b].x, y[b].y, y[b].z, y[b].w};
                        *reinterpret_cast<uint4*>(Bbase + mn_off(g0 + b, tid, gN)) = pack8(w);
                    }
                }
            }
        }
        if (a.ext_in) {
            __syncthreads();
            for (int idx = tid; idx < Fr * nch; idx += T2_THREADS) {      // joint-mean rows, averaged in fp32
                const int kc = idx % nch, ff = idx / nch;
                if (f0 + ff >= n_frames) continue;
                float s[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] = 0.f;
                for (int v = 0; v < a.Vin; ++v) {
                    float t[8];
                    unpack8(*reinterpret_cast<const uint4*>(Abase + op_off(ff * rpf + v, kc, nch)), t);
#pragma unroll
                    for (int e = 0; e < 8; ++e) s[e] += t[e];
                }
                const float inv = 1.f / (float)a.Vin;
#pragma unroll
                for (int e = 0; e < 8; ++e) s[e] *= inv;
                *reinterpret_cast<uint4*>(Abase + op_off(ff * rpf + a.Vin, kc, nch)) = pack8(s);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t sbo = (uint32_t)nch * 128u;
            const uint32_t a0 = smem_u32(Abase), b0 = smem_u32(Bbase);
            const uint32_t gN = (uint32_t)(Ntp >> 3);
            if (wmode == 3) { mbar_wait(&wbar, wphase); wphase ^= 1; }
            for (int ks = 0; ks < (kv_len >> 4); ++ks) {
                const uint64_t ad = make_desc(a0 + ks * 256u, 128u, sbo);
                const uint64_t bd = wmode == 1 ? make_desc(b0 + ks * 2u * gN * 128u, gN * 128u, 128u) : make_desc(b0 + ks * 256u, 128u, sbo);
                umma_f16(tmem_d, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
            }
            umma_commit(&mbar);
        }
        first = 0;
        phase ^= 1;
    }
    mbar_wait(&mbar, phase ^ 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- fused tail on my accumulator row, 16 columns at a time
    const float inv_ext = a.contract_ext ? 1.f / (float)(rpf - 1) : 0.f;
    const int samp = row_ok ? (int)(f / a.T_out) : 0;
    bf16* out = reinterpret_cast<bf16*>(a.out);
    const bf16* addp = reinterpret_cast<const bf16*>(a.add);
    const bf16* add2p = reinterpret_cast<const bf16*>(a.add2);
    const bf16* partp = reinterpret_cast<const bf16*>(a.partner);
    for (int c16 = 0; c16 < Ntp; c16 += 16) {
        float v[16];
        tmem_ld16(tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)c16, v);
        const int c = n0 + c16;
        const bool live0 = c < a.N, live1 = c + 8 < a.N;             // N % 8 == 0: the chunk is live in halves
        // issue this row's loads early
        uint4 ra[2], ra2[2], rp[2];
        Act8Raw rm[2];
        if (orow >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!(h ? live1 : live0)) continue;
                if (addp) ra[h] = *reinterpret_cast<const uint4*>(addp + orow * a.ld_add + c + h * 8);
                if (add2p) ra2[h] = *reinterpret_cast<const uint4*>(add2p + orow * a.ld_add2 + c + h * 8);
                if (partp) rp[h] = *reinterpret_cast<const uint4*>(partp + orow * a.ld_partner + c + h * 8);
                if (a.has_mask) rm[h] = act8_issue(a.mask, orow, c + h * 8);
            }
        }
        if (a.contract_ext) {                                           // fold the joint-mean row back into its frame
            const int buf = (c16 >> 4) & 1;
            if (row_ok && is_ext_row && fl < 8) {
#pragma unroll
                for (int e = 0; e < 16; ++e) ext_s[buf][fl][e] = v[e];
            }
            __syncthreads();
            if (orow >= 0) {
#pragma unroll
                for (int e = 0; e < 16; ++e) v[e] = fmaf(ext_s[buf][fl][e], inv_ext, v[e]);
            }
        }
        float s1[16], s2[16];
#pragma unroll
        for (int e = 0; e < 16; ++e) s1[e] = s2[e] = 0.f;
        if (orow >= 0) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!(h ? live1 : live0)) continue;
                float* vv = v + h * 8;
                const int ch = c + h * 8;
                add8(vv, tl_bias + c16 + h * 8);
                if (addp) { float t[8]; unpack8(ra[h], t);
#pragma unroll
                    for (int e = 0; e < 8; ++e) vv[e] += t[e]; }
                if (add2p) { float t[8]; unpack8(ra2[h], t);
#pragma unroll
                    for (int e = 0; e < 8; ++e) vv[e] += t[e]; }
                if (a.bcast) {
                    float t[8];
                    load8f(a.bcast + ((long long)samp * Vout + j) * a.N + ch, t, 0.f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) vv[e] = fmaf(t[e], a.bcast_scale, vv[e]);
                }
                if (a.has_mask) {
                    float m[8];
                    finish_smem(rm[h], a.mask.x2 != nullptr, 0, tl_ma1 + c16 + h * 8, tl_mb + c16 + h * 8, tl_ma2 + c16 + h * 8, m);
#pragma unroll
                    for (int e = 0; e < 8; ++e) vv[e] = m[e] > 0.f ? vv[e] : 0.f;
                }
                if (a.stat_sum) {
                    float p[8];
                    if (partp) unpack8(rp[h], p);
#pragma unroll
                    for (int e = 0; e < 8; ++e) { s1[h * 8 + e] = vv[e]; s2[h * 8 + e] = vv[e] * (partp ? p[e] : vv[e]); }
                }
                *reinterpret_cast<uint4*>(out + orow * a.ld_out + ch) = pack8(vv);
            }
        }
        if (a.stat_sum) {                                               // warp-uniform: every lane takes part in the shuffles
            const float t1 = warp_colsum16(s1, lane), t2 = warp_colsum16(s2, lane);
            if ((lane & 1) == 0) {
                s_acc[0][warp][c16 + (lane >> 1)] = t1;
                s_acc[1][warp][c16 + (lane >> 1)] = t2;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
    if (a.stat_sum && tid < Nt) {
        const float t1 = s_acc[0][0][tid] + s_acc[0][1][tid] + s_acc[0][2][tid] + s_acc[0][3][tid];
        const float t2 = s_acc[1][0][tid] + s_acc[1][1][tid] + s_acc[1][2][tid] + s_acc[1][3][tid];
        atomicAdd(a.stat_sum + n0 + tid, (double)t1);
        atomicAdd(a.stat_sq + n0 + tid, (double)t2);
    }
}

static const char* launch_conv_gemm_tc2(const dsg_conv_gemm_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    const int rpf = a.Vin + a.ext_in;
    if (a.dtype != DSG_BF16 || a.taps != 1 || rpf > 128 || a.K % 8 != 0 || a.N % 8 != 0) return nullptr;
    if (a.contract_ext && (a.ext_in || a.Vin < 2 || 128 / rpf > 8)) return nullptr;
    if (!act8_ok(a.src) || (!a.wpack && (uintptr_t)a.W % 16 != 0)) return nullptr;
    int wmode;
    if (a.wpack && (uintptr_t)a.wpack % 128 == 0) wmode = 3;
    else if (a.ws_k == 1 && a.ws_n % 4 == 0) wmode = 0;
    else if (a.ws_n == 1 && a.ws_k % 4 == 0) wmode = 1;
    else return nullptr;
    auto al16 = [](const void* p, long long ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 8 == 0); };
    if (!(al16(a.out, a.ld_out) && al16(a.add, a.ld_add) && al16(a.add2, a.ld_add2) && al16(a.partner, a.ld_partner) &&
          (!a.has_mask || act8_ok(a.mask)) && (!a.bias || (uintptr_t)a.bias % 16 == 0) && (!a.bcast || (uintptr_t)a.bcast % 16 == 0)))
        return nullptr;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0 || a.N <= 0) { *handled = true; return nullptr; }
    const int Kp = (a.K + 15) & ~15;
    const int kpass = Kp < T2_KPASS ? Kp : T2_KPASS;
    const size_t smem = (size_t)(128 + T2_BN) * kpass * 2 + (size_t)(3 * Kp + 4 * T2_BN) * sizeof(float);
    const int Fr = 128 / rpf;
    dim3 grid((unsigned)((n_frames + Fr - 1) / Fr), (unsigned)((a.N + T2_BN - 1) / T2_BN));
    if (wmode == 3) {
        conv_wpack_kernel<<<dim3(grid.y, (unsigned)((Kp + kpass - 1) / kpass)), dim3(256), 0, st>>>(a.W, a.ws_n, a.ws_k, a.K, a.N, reinterpret_cast<unsigned char*>(a.wpack));
        if (const char* e = dsg_launch_error()) return e;
    }
    cudaFuncSetAttribute(conv_gemm_tc2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_gemm_tc2_kernel<<<grid, dim3(T2_THREADS), smem, st>>>(a, wmode);
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc
}  // namespace dsg
#endif
