// Channels-last (taps x 1) temporal-convolution GEMM with fused prologue / epilogue, and its
// weight gradient.  fp32-accurate CUDA-core engine (FFMA): this is the exact path used for the
// fp32 parity mode and for every shape; the bf16 tensor-core engine lives in conv_gemm_tc.cuh.
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int CG_BM = 128;   // GEMM rows per CTA (whole frames)
constexpr int CG_BN = 64;    // output channels per CTA
constexpr int CG_BK = 16;
constexpr int CG_THREADS = 256;

struct FrameMap {
    int taps, tap_step, tap_off, t_mul, t_div, T_in, T_out, Vin, ext_in;
};

// source frame index for (output frame f, tap) or -1
DSG_D long long src_frame(const FrameMap& m, long long f, int tap) {
    long long n = f / m.T_out;
    int t = (int)(f - n * m.T_out);
    int num = t * m.t_mul + tap * m.tap_step + m.tap_off;
    if (num < 0 || (num % m.t_div) != 0) return -1;
    int ti = num / m.t_div;
    if (ti >= m.T_in) return -1;
    return n * m.T_in + ti;
}

// Fused epilogue tail shared by the CUDA-core and the tcgen05 kernels: the fp32 accumulator tile is staged in
// shared memory (Cs[row][col], pitch ldc); thread (c = tid % BN, row group = tid / BN) walks output rows.
template <class T, int BN>
DSG_D void gemm_tail(const dsg_conv_gemm_args& a, const float* Cs, int ldc, float* s_red /* [2][THREADS/BN][BN] */,
                     long long f0, int Fr, int rpf, int n0, long long n_frames) {
    constexpr int RG = CG_THREADS / BN;
    const int tid = threadIdx.x;
    const int Vout = rpf - a.contract_ext;
    const int c = tid % BN, rgrp = tid / BN;
    const int cg = n0 + c;
    float s1 = 0.f, s2 = 0.f;
    if (cg < a.N) {
        const float bias = a.bias ? a.bias[cg] : 0.f;
        const float inv_ext = a.contract_ext ? 1.f / (float)(rpf - 1) : 0.f;
        T* out = reinterpret_cast<T*>(a.out);
        for (int lr = rgrp; lr < Fr * Vout; lr += RG) {
            int fl = lr / Vout, j = lr - fl * Vout;
            long long f = f0 + fl;
            if (f >= n_frames) break;
            float v = Cs[(fl * rpf + j) * ldc + c];
            if (a.contract_ext) v += Cs[(fl * rpf + rpf - 1) * ldc + c] * inv_ext;
            v += bias;
            long long orow = f * Vout + j;
            if (a.add) v += ldf<T>(reinterpret_cast<const T*>(a.add) + orow * a.ld_add + cg);
            if (a.add2) v += ldf<T>(reinterpret_cast<const T*>(a.add2) + orow * a.ld_add2 + cg);
            if (a.bcast) {
                long long n = f / a.T_out;
                v += a.bcast[(n * Vout + j) * a.N + cg] * a.bcast_scale;
            }
            if (a.has_mask && !(act_value<T>(a.mask, orow, cg) > 0.f)) v = 0.f;
            if (a.stat_sum) {
                float p = a.partner ? ldf<T>(reinterpret_cast<const T*>(a.partner) + orow * a.ld_partner + cg) : v;
                s1 += v;
                s2 += v * p;
            }
            stf<T>(out + orow * a.ld_out + cg, v);
        }
    }
    if (a.stat_sum) {
        s_red[(0 * RG + rgrp) * BN + c] = s1;
        s_red[(1 * RG + rgrp) * BN + c] = s2;
        __syncthreads();
        if (tid < BN && n0 + tid < a.N) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int g = 0; g < RG; ++g) { t1 += s_red[(0 * RG + g) * BN + tid]; t2 += s_red[(1 * RG + g) * BN + tid]; }
            atomicAdd(a.stat_sum + n0 + tid, (double)t1);
            atomicAdd(a.stat_sq + n0 + tid, (double)t2);
        }
    }
}

template <class T>
__global__ void __launch_bounds__(CG_THREADS) conv_gemm_kernel(dsg_conv_gemm_args a) {
    // shared memory: operand chunks, re-used as the fp32 output staging tile in the epilogue
    DSG_SHARED __align__(16) float smem_f[CG_BM * (CG_BN + 1)];
    DSG_SHARED long long rowsrc[CG_BM];
    DSG_SHARED float s_red[2][4][CG_BN];
    float* As = smem_f;                               // [BK][BM+4]
    float* Bs = smem_f + CG_BK * (CG_BM + 4);         // [BK][BN+4]
    float* Cs = smem_f;                               // [BM][BN+1]

    const int tid = threadIdx.x;
    const int rpf = a.Vin + a.ext_in;                 // GEMM rows per frame
    const int Fr = CG_BM / rpf > 0 ? CG_BM / rpf : 1; // frames per tile
    const int rows_tile = Fr * rpf;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const long long f0 = (long long)blockIdx.x * Fr;
    const int n0 = blockIdx.y * CG_BN;
    FrameMap fm{a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};

    const int ty = tid / 16, tx = tid % 16;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

    for (int tap = 0; tap < a.taps; ++tap) {
        __syncthreads();
        if (tid < CG_BM) {
            long long sr = -1;
            if (tid < rows_tile) {
                int fl = tid / rpf, j = tid - fl * rpf;
                long long f = f0 + fl;
                if (f < n_frames) {
                    long long sf = src_frame(fm, f, tap);
                    if (sf >= 0) sr = (j < a.Vin) ? sf * a.Vin + j : -2;   // -2: joint-mean row
                }
            }
            rowsrc[tid] = sr;
        }
        __syncthreads();
        for (int k0 = 0; k0 < a.K; k0 += CG_BK) {
            // ---- A chunk (prologue fused) ----
            for (int idx = tid; idx < CG_BM * CG_BK; idx += CG_THREADS) {
                int k = idx % CG_BK, row = idx / CG_BK;
                long long sr = rowsrc[row];
                float v = 0.f;
                if (sr >= 0 && k0 + k < a.K) v = act_value<T>(a.src, sr, k0 + k);
                As[k * (CG_BM + 4) + row] = v;
            }
            // ---- B chunk ----
            for (int idx = tid; idx < CG_BN * CG_BK; idx += CG_THREADS) {
                int k = idx % CG_BK, n = idx / CG_BK;
                float w = 0.f;
                if (n0 + n < a.N && k0 + k < a.K)
                    w = a.W[(long long)(n0 + n) * a.ws_n + (long long)(k0 + k) * a.ws_k + (long long)tap * a.ws_tap];
                Bs[k * (CG_BN + 4) + n] = w;
            }
            __syncthreads();
            if (a.ext_in) {
                for (int idx = tid; idx < Fr * CG_BK; idx += CG_THREADS) {
                    int k = idx % CG_BK, fl = idx / CG_BK;
                    int mr = fl * rpf + a.Vin;
                    if (rowsrc[mr] == -2) {
                        float s = 0.f;
                        for (int j = 0; j < a.Vin; ++j) s += As[k * (CG_BM + 4) + fl * rpf + j];
                        As[k * (CG_BM + 4) + mr] = s / (float)a.Vin;
                    }
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < CG_BK; ++k) {
                float av[8], bv[4];
#pragma unroll
                for (int i = 0; i < 8; ++i) av[i] = As[k * (CG_BM + 4) + ty * 8 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) bv[j] = Bs[k * (CG_BN + 4) + tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
    }
    // ---- epilogue: stage the tile, then row-wise fused tail ----
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) Cs[(ty * 8 + i) * (CG_BN + 1) + tx * 4 + j] = acc[i][j];
    __syncthreads();

    gemm_tail<T, CG_BN>(a, Cs, CG_BN + 1, &s_red[0][0][0], f0, Fr, rpf, n0, n_frames);
}

// ---------------------------------------------------------------------------------------------
// Skinny outputs (N <= 8: the data gradient towards the 3-channel network input): a GEMM tile would idle, the op is a
// pure stream over the input rows.  Thread = output row; weights and prologue coefficients in shared memory; bf16 rows
// are read with 4 independent 16-byte loads in flight.  Epilogue: bias, addends, per-sample broadcast.
constexpr int SK_KMAX = 512;
template <class T>
__global__ void __launch_bounds__(256) conv_gemm_skinny_kernel(dsg_conv_gemm_args a, int vec) {
    DSG_SHARED __align__(16) float Ws[8 * SK_KMAX];          // [n][k]
    DSG_SHARED __align__(16) float cf[3][SK_KMAX];           // a1, b1 + b2, a2
    const int tid = threadIdx.x, K = a.K, N = a.N;
    for (int idx = tid; idx < N * K; idx += 256) {
        const int k = idx % K, n = idx / K;
        Ws[n * K + k] = a.W[(long long)n * a.ws_n + (long long)k * a.ws_k];
    }
    for (int k = tid; k < K; k += 256) {
        cf[0][k] = a.src.a1 ? a.src.a1[k] : 1.f;
        cf[1][k] = (a.src.b1 ? a.src.b1[k] : 0.f) + (a.src.b2 ? a.src.b2[k] : 0.f);
        cf[2][k] = a.src.a2 ? a.src.a2[k] : 1.f;
    }
    __syncthreads();
    const long long rows_out = (long long)a.n_samples * a.T_out * a.Vin;
    const long long row = (long long)blockIdx.x * 256 + tid;
    if (row >= rows_out) return;
    const long long f = row / a.Vin;
    const int j = (int)(row - f * a.Vin);
    FrameMap fm{1, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, 0};
    const long long sf = src_frame(fm, f, 0);
    float acc[8];
#pragma unroll
    for (int n = 0; n < 8; ++n) acc[n] = (n < N && a.bias) ? a.bias[n] : 0.f;
    if (sf >= 0) {
        const long long sr = sf * a.Vin + j;
        int k0 = 0;
        if (vec) {                                               // bf16, 16-byte aligned rows
            const bf16* x1 = reinterpret_cast<const bf16*>(a.src.x1) + sr * a.src.ld1;
            const bf16* x2 = a.src.x2 ? reinterpret_cast<const bf16*>(a.src.x2) + sr * a.src.ld2 : nullptr;
            for (; k0 + 8 <= K; k0 += 32) {
                uint4 r1[4], r2[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int k = k0 + b * 8;
                    if (k + 8 <= K) {
                        r1[b] = *reinterpret_cast<const uint4*>(x1 + k);
                        if (x2) r2[b] = *reinterpret_cast<const uint4*>(x2 + k);
                    }
                }
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int k = k0 + b * 8;
                    if (k + 8 > K) continue;
                    float x[8], v[8];
                    unpack8(r1[b], x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], cf[0][k + e], cf[1][k + e]);
                    if (x2) {
                        unpack8(r2[b], x);
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], cf[2][k + e], v[e]);
                    }
                    if (a.src.relu) {
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
                    }
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        if (n >= N) break;
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[n] = fmaf(v[e], Ws[n * K + k + e], acc[n]);
                    }
                }
            }
            k0 = K & ~7;                                         // leftover channels (K % 8) on the scalar path
        }
        for (int k = k0; k < K; ++k) {
            const float v = act_value<T>(a.src, sr, k);
#pragma unroll
            for (int n = 0; n < 8; ++n)
                if (n < N) acc[n] = fmaf(v, Ws[n * K + k], acc[n]);
        }
    }
    const int samp = (int)(f / a.T_out);
#pragma unroll
    for (int n = 0; n < 8; ++n) {
        if (n >= N) break;
        float v = acc[n];
        if (a.add) v += ldf<T>(reinterpret_cast<const T*>(a.add) + row * a.ld_add + n);
        if (a.add2) v += ldf<T>(reinterpret_cast<const T*>(a.add2) + row * a.ld_add2 + n);
        if (a.bcast) v = fmaf(a.bcast[((long long)samp * a.Vin + j) * N + n], a.bcast_scale, v);
        stf<T>(reinterpret_cast<T*>(a.out) + row * a.ld_out + n, v);
    }
}

// ---------------------------------------------------------------------------------------------
constexpr int WG_BK = 64;    // dW tile: input channels
constexpr int WG_BN = 64;    // dW tile: output channels
constexpr int WG_BR = 64;    // GEMM rows per inner step (whole frames)

template <class T>
__global__ void __launch_bounds__(CG_THREADS) conv_wgrad_kernel(dsg_conv_wgrad_args a, int frames_per_cta) {
    DSG_SHARED float At[WG_BR][WG_BK + 1];
    DSG_SHARED float Bt[WG_BR][WG_BN + 1];
    DSG_SHARED long long rowsrc[WG_BR];
    DSG_SHARED long long rowdst[WG_BR];
    const int tid = threadIdx.x;
    const int rpf = a.Vin + a.ext_in;
    const int Fr = WG_BR / rpf > 0 ? WG_BR / rpf : 1;
    const int rows_tile = Fr * rpf;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const int ktiles = (a.K + WG_BK - 1) / WG_BK;
    const int kt = blockIdx.y % ktiles, tap = blockIdx.y / ktiles;
    const int k0 = kt * WG_BK, n0 = blockIdx.z * WG_BN;
    const long long fbeg = (long long)blockIdx.x * frames_per_cta;
    long long fend = fbeg + frames_per_cta;
    if (fend > n_frames) fend = n_frames;
    FrameMap fm{a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};
    const int ty = tid / 16, tx = tid % 16;   // 4 k x 4 n per thread
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    float bsum = 0.f;    // db partial: thread tid < WG_BN owns column n0+tid (only in the kt==0, tap==0 CTAs)
    const bool do_bias = (a.db != nullptr) && kt == 0 && tap == 0;

    for (long long f0 = fbeg; f0 < fend; f0 += Fr) {
        __syncthreads();
        if (tid < WG_BR) {
            long long sr = -1, dr = -1;
            if (tid < rows_tile) {
                int fl = tid / rpf, j = tid - fl * rpf;
                long long f = f0 + fl;
                if (f < fend) {
                    dr = f * rpf + j;
                    long long sf = src_frame(fm, f, tap);
                    if (sf >= 0) sr = (j < a.Vin) ? sf * a.Vin + j : -2;
                }
            }
            rowsrc[tid] = sr;
            rowdst[tid] = dr;
        }
        __syncthreads();
        for (int idx = tid; idx < WG_BR * WG_BK; idx += CG_THREADS) {
            int k = idx % WG_BK, row = idx / WG_BK;
            long long sr = rowsrc[row];
            float v = 0.f;
            if (sr >= 0 && k0 + k < a.K) v = act_value<T>(a.A, sr, k0 + k);
            At[row][k] = v;
        }
        for (int idx = tid; idx < WG_BR * WG_BN; idx += CG_THREADS) {
            int n = idx % WG_BN, row = idx / WG_BN;
            long long dr = rowdst[row];
            float v = 0.f;
            if (dr >= 0 && n0 + n < a.N) v = act_value<T>(a.B, dr, n0 + n);
            Bt[row][n] = v;
        }
        __syncthreads();
        if (a.ext_in) {
            for (int idx = tid; idx < Fr * WG_BK; idx += CG_THREADS) {
                int k = idx % WG_BK, fl = idx / WG_BK;
                int mr = fl * rpf + a.Vin;
                if (rowsrc[mr] == -2) {
                    float s = 0.f;
                    for (int j = 0; j < a.Vin; ++j) s += At[fl * rpf + j][k];
                    At[mr][k] = s / (float)a.Vin;
                }
            }
            __syncthreads();
        }
        for (int r = 0; r < rows_tile; ++r) {
            float av[4], bv[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) av[i] = At[r][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) bv[j] = Bt[r][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (do_bias && tid < WG_BN) {
            for (int r = 0; r < rows_tile; ++r) bsum += Bt[r][tid];
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int k = k0 + ty * 4 + i;
        if (k >= a.K) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int n = n0 + tx * 4 + j;
            if (n < a.N) atomicAdd(a.dW + (long long)n * a.ws_n + (long long)k * a.ws_k + (long long)tap * a.ws_tap, acc[i][j]);
        }
    }
    if (do_bias && tid < WG_BN && n0 + tid < a.N) atomicAdd(a.db + n0 + tid, bsum);
}

// shapes the skinny kernel takes
static inline bool conv_gemm_skinny_ok(const dsg_conv_gemm_args& a) {
    return a.N <= 8 && a.K <= SK_KMAX && a.taps == 1 && !a.ext_in && !a.contract_ext && !a.has_mask && !a.stat_sum && !a.partner;
}
template <class T> static const char* launch_conv_gemm_skinny(const dsg_conv_gemm_args& a, dsg_stream_t st) {
    const long long rows_out = (long long)a.n_samples * a.T_out * a.Vin;
    if (rows_out <= 0 || a.N <= 0) return nullptr;
    const int vec = (sizeof(T) == 2) && act8_ok(a.src) ? 1 : 0;
    dsg_launch(conv_gemm_skinny_kernel<T>, dim3((unsigned)((rows_out + 255) / 256)), dim3(256), 0, st, a, vec);
    return dsg_launch_error();
}

template <class T> static const char* launch_conv_gemm(const dsg_conv_gemm_args& a, dsg_stream_t st) {
    int rpf = a.Vin + a.ext_in;
    if (rpf > CG_BM) return "conv_gemm: more than 128 rows per frame";
    if (a.contract_ext && (a.ext_in || a.Vin < 2)) return "conv_gemm: contract_ext needs Vin>=2 rows and no ext_in";
    int Fr = CG_BM / rpf;
    long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0 || a.N <= 0) return nullptr;
    dim3 grid((unsigned)((n_frames + Fr - 1) / Fr), (unsigned)((a.N + CG_BN - 1) / CG_BN));
    dsg_launch(conv_gemm_kernel<T>, grid, dim3(CG_THREADS), 0, st, a);
    return dsg_launch_error();
}

template <class T> static const char* launch_conv_wgrad(const dsg_conv_wgrad_args& a, dsg_stream_t st) {
    int rpf = a.Vin + a.ext_in;
    if (rpf > WG_BR) return "conv_wgrad: more than 64 rows per frame";
    long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0) return nullptr;
    int ktiles = (a.K + WG_BK - 1) / WG_BK, ntiles = (a.N + WG_BN - 1) / WG_BN;
    int Fr = WG_BR / rpf;
    // aim for ~4 CTAs per SM in total; each CTA walks a contiguous frame range
    long long want = (4 * dsg_num_sms() + (long long)ktiles * ntiles * a.taps - 1) / ((long long)ktiles * ntiles * a.taps);
    long long fpc = (n_frames + want - 1) / want;
    fpc = (fpc + Fr - 1) / Fr * Fr;
    if (fpc < Fr) fpc = Fr;
    dim3 grid((unsigned)((n_frames + fpc - 1) / fpc), (unsigned)(ktiles * a.taps), (unsigned)ntiles);
    dsg_launch(conv_wgrad_kernel<T>, grid, dim3(CG_THREADS), 0, st, a, (int)fpc);
    return dsg_launch_error();
}

}  // namespace dsg
