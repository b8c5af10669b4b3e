// bf16 tensor-core engine of dsg_conv_gemm for sm_100a: tcgen05.mma (cta_group::1, kind::f16, M=128) with the
// fp32 accumulator in TMEM.  Same frame map, fused prologue and fused epilogue tail as the CUDA-core kernel in
// conv_gemm.cuh; only the inner product differs.
//
// Per CTA: one 128-row tile (whole frames) x one <=128-column tile.  The (tap, k) reduction axis is walked in
// passes of <= 256 "virtual k": for each pass the CTA
//   1. builds the A operand in shared memory: rows are loaded from HBM (16-byte vectors when aligned), the fused
//      prologue (BN affine, residual, ReLU) is applied in registers, and bf16 values are written in the canonical
//      K-major no-swizzle UMMA layout (8x16B core matrices; LBO = 128 B between the two k-halves of an MMA,
//      SBO = 128 B * chunks between 8-row groups) — cute::UMMA::LayoutType::SWIZZLE_NONE, Major::K;
//   2. builds the B operand (weights fp32 -> bf16) the same way;
//   3. fence.proxy.async + barrier, then ONE thread issues kv/16 tcgen05.mma and a tcgen05.commit to an mbarrier.
// After the last pass every warp reads its TMEM lane quarter with tcgen05.ld (32x32b.x16), stages the fp32 tile in
// shared memory and runs the shared fused tail (bias, addends, mask, BN statistics, bf16 store).
#pragma once
#include "conv_gemm.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc {

constexpr int TC_BM = 128;
constexpr int TC_BN = 128;
constexpr int TC_KPASS = 256;       // virtual-k per pass
constexpr int TC_MAX_TAPS = 9;

DSG_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

DSG_D void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
DSG_D bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
DSG_D void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();     // never hang the GPU: a lost arrive becomes a launch error
    }
}
DSG_D void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
DSG_D void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
DSG_D void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
DSG_D void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_D void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 | LBO>>4 <<16 |
// SBO>>4 <<32 | version=1 <<46 | layout_type=0 <<61
DSG_D uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// cute::UMMA::InstrDescriptor: D=f32 (1<<4), A=bf16 (1<<7), B=bf16 (1<<10), K-major A/B, N>>3 at [17,23), M>>4 at [24,29)
DSG_D uint32_t make_idesc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct Pass {
    int kv0, kv_len;    // virtual-k range [kv0, kv0+kv_len), kv_len % 16 == 0
};

// element (row r, 16-byte chunk kc) of an operand tile with `nchunks` chunks per row
DSG_D uint32_t op_off(int r, int kc, int nchunks) { return (uint32_t)(((r >> 3) * nchunks + kc) * 128 + (r & 7) * 16); }

// MN-major operand tile: element (channel group g8, reduction index kk) -> byte offset; `ngroups` groups per tile
DSG_D uint32_t mn_off(int g8, int kk, int ngroups) { return (uint32_t)(((kk >> 3) * ngroups + g8) * 128 + (kk & 7) * 16); }
DSG_D uint32_t make_idesc_bmn(int M, int N) { return make_idesc(M, N) | (1u << 16); }      // B operand MN-major

DSG_D void mul8(float* v, const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] *= a.x; v[1] *= a.y; v[2] *= a.z; v[3] *= a.w; v[4] *= b.x; v[5] *= b.y; v[6] *= b.z; v[7] *= b.w;
}
DSG_D void add8(float* v, const float* p) {
    float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
}

// Raw 16-byte loads of an activation source, issued early so several are in flight per thread
struct Act8Raw { uint4 a, b; };
DSG_D Act8Raw act8_issue(const ActSrc& s, long long row, int c) {
    Act8Raw r;
    r.a = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x1) + row * s.ld1 + c);
    r.b = s.x2 ? *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(s.x2) + row * s.ld2 + c) : make_uint4(0u, 0u, 0u, 0u);
    return r;
}
DSG_D void act8_finish_f(const ActSrc& s, const Act8Raw& r, int c, float* v) {
    unpack8(r.a, v);
    if (s.a1) mul8(v, s.a1 + c);
    if (s.b1) add8(v, s.b1 + c);
    if (s.x2) {
        float w[8];
        unpack8(r.b, w);
        if (s.a2) mul8(w, s.a2 + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] += w[j];
    }
    if (s.b2) add8(v, s.b2 + c);
    if (s.relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
}
DSG_D uint4 act8_finish(const ActSrc& s, const Act8Raw& r, int c) {
    float v[8];
    act8_finish_f(s, r, c, v);
    return pack8(v);
}

// 8 consecutive channels of an activation source at (row, c) -> packed bf16x8 (vector path when aligned)
DSG_D uint4 load_act8(const ActSrc& s, long long row, int c, int C, int vec_ok) {
    if (vec_ok && c + 8 <= C) return act8_finish(s, act8_issue(s, row, c), c);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = (c + j < C) ? act_value<bf16>(s, row, c + j) : 0.f;
    return pack8(v);
}

// v = relu?( x1*a1 + b + x2*a2 ) with the coefficients read from shared memory (16-byte broadcast loads)
DSG_D void finish_smem(const Act8Raw& r, bool has_x2, int relu, const float* a1, const float* b, const float* a2, float* v) {
    float x[8], c1[8], cb[8];
    unpack8(r.a, x);
    load8f(a1, c1, 1.f);
    load8f(b, cb, 0.f);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = fmaf(x[j], c1[j], cb[j]);
    if (has_x2) {
        float y[8], c2[8];
        unpack8(r.b, y);
        load8f(a2, c2, 1.f);
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaf(y[j], c2[j], v[j]);
    }
    if (relu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
    }
}

constexpr int TC_LDC = TC_BN + 4;      // fp32 staging pitch: 33 x 16 B, conflict-free for 16-byte row-wise and column-wise access

// Vectorised fused tail: thread = (8-column chunk, row lane); rows come from a per-CTA table (no per-element div/mod).
DSG_D void gemm_tail_vec(const dsg_conv_gemm_args& a, const float* Cs, const long long* orow, const int* osamp, int n_out_rows,
                         int rpf, int n0, float* s_red /* [2][8][TC_BN] */) {
    const int tid = threadIdx.x, cc = tid & 15, rl = tid >> 4;        // 16 chunks x 16 row lanes
    const int Vout = rpf - a.contract_ext;
    const int c = n0 + cc * 8;
    const bool live = c < a.N;                                        // N % 8 == 0 on this path
    float s1[8], s2[8], bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s1[j] = s2[j] = 0.f;
    if (live) load8f(a.bias ? a.bias + c : nullptr, bias, 0.f);
    const float inv_ext = a.contract_ext ? 1.f / (float)(rpf - 1) : 0.f;
    bf16* out = reinterpret_cast<bf16*>(a.out);
    if (live) {
        constexpr int RB = 2;                                 // rows in flight per thread
        const bf16* addp = reinterpret_cast<const bf16*>(a.add);
        const bf16* add2p = reinterpret_cast<const bf16*>(a.add2);
        const bf16* partp = reinterpret_cast<const bf16*>(a.partner);
        for (int lr0 = rl; lr0 < n_out_rows; lr0 += 16 * RB) {
            long long rr[RB];
            uint4 ra[RB], ra2[RB], rp[RB];
            Act8Raw rm[RB];
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const int lr = lr0 + b * 16;
                rr[b] = lr < n_out_rows ? orow[lr] : -1;
                if (rr[b] < 0) continue;
                if (addp) ra[b] = *reinterpret_cast<const uint4*>(addp + rr[b] * a.ld_add + c);
                if (add2p) ra2[b] = *reinterpret_cast<const uint4*>(add2p + rr[b] * a.ld_add2 + c);
                if (partp) rp[b] = *reinterpret_cast<const uint4*>(partp + rr[b] * a.ld_partner + c);
                if (a.has_mask) rm[b] = act8_issue(a.mask, rr[b], c);
            }
#pragma unroll
            for (int b = 0; b < RB; ++b) {
                const long long r = rr[b];
                if (r < 0) continue;
                const int lr = lr0 + b * 16;
                const int fl = lr / Vout, j0 = lr - fl * Vout;
                const float* cp = Cs + (fl * rpf + j0) * TC_LDC + cc * 8;
                float v[8];
                {
                    float4 x = *reinterpret_cast<const float4*>(cp), y = *reinterpret_cast<const float4*>(cp + 4);
                    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w; v[4] = y.x; v[5] = y.y; v[6] = y.z; v[7] = y.w;
                }
                if (a.contract_ext) {
                    const float* ep = Cs + (fl * rpf + rpf - 1) * TC_LDC + cc * 8;
                    float4 x = *reinterpret_cast<const float4*>(ep), y = *reinterpret_cast<const float4*>(ep + 4);
                    v[0] = fmaf(x.x, inv_ext, v[0]); v[1] = fmaf(x.y, inv_ext, v[1]); v[2] = fmaf(x.z, inv_ext, v[2]); v[3] = fmaf(x.w, inv_ext, v[3]);
                    v[4] = fmaf(y.x, inv_ext, v[4]); v[5] = fmaf(y.y, inv_ext, v[5]); v[6] = fmaf(y.z, inv_ext, v[6]); v[7] = fmaf(y.w, inv_ext, v[7]);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] += bias[j];
                if (addp) {
                    float t[8];
                    unpack8(ra[b], t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += t[j];
                }
                if (add2p) {
                    float t[8];
                    unpack8(ra2[b], t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] += t[j];
                }
                if (a.bcast) {
                    const float* bp = a.bcast + ((long long)osamp[lr] * Vout + j0) * a.N + c;
                    float4 x = *reinterpret_cast<const float4*>(bp), y = *reinterpret_cast<const float4*>(bp + 4);
                    v[0] = fmaf(x.x, a.bcast_scale, v[0]); v[1] = fmaf(x.y, a.bcast_scale, v[1]); v[2] = fmaf(x.z, a.bcast_scale, v[2]);
                    v[3] = fmaf(x.w, a.bcast_scale, v[3]); v[4] = fmaf(y.x, a.bcast_scale, v[4]); v[5] = fmaf(y.y, a.bcast_scale, v[5]);
                    v[6] = fmaf(y.z, a.bcast_scale, v[6]); v[7] = fmaf(y.w, a.bcast_scale, v[7]);
                }
                if (a.has_mask) {
                    float m[8];
                    act8_finish_f(a.mask, rm[b], c, m);
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = m[j] > 0.f ? v[j] : 0.f;
                }
                if (a.stat_sum) {
                    float p[8];
                    if (partp) unpack8(rp[b], p);
#pragma unroll
                    for (int j = 0; j < 8; ++j) { s1[j] += v[j]; s2[j] += v[j] * (partp ? p[j] : v[j]); }
                }
                *reinterpret_cast<uint4*>(out + r * a.ld_out + c) = pack8(v);
            }
        }
    }
    if (a.stat_sum) {
        // lanes l and l^16 hold the same chunk (row lanes 2w, 2w+1): fold them, then one partial per warp
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s1[j] += __shfl_xor_sync(0xffffffffu, s1[j], 16);
            s2[j] += __shfl_xor_sync(0xffffffffu, s2[j], 16);
        }
        const int warp = tid >> 5;
        if ((tid & 31) < 16) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s_red[(0 * 8 + warp) * TC_BN + cc * 8 + j] = s1[j];
                s_red[(1 * 8 + warp) * TC_BN + cc * 8 + j] = s2[j];
            }
        }
        __syncthreads();
        if (tid < TC_BN && n0 + tid < a.N) {
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { t1 += s_red[(0 * 8 + w) * TC_BN + tid]; t2 += s_red[(1 * 8 + w) * TC_BN + tid]; }
            atomicAdd(a.stat_sum + n0 + tid, (double)t1);
            atomicAdd(a.stat_sq + n0 + tid, (double)t2);
        }
    }
}

// wmode: 0 = weights contiguous along k (ws_k == 1, 16-byte aligned rows): K-major B, vector loads
//        1 = weights contiguous along n (ws_n == 1): MN-major B, vector loads
//        2 = anything else (temporal taps): K-major B, scalar loads
__global__ void __launch_bounds__(CG_THREADS, 2) conv_gemm_tc_kernel(dsg_conv_gemm_args a, int Kp, int vec_ok, int wmode, int vec_tail) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_red[2 * 8 * TC_BN];
    __shared__ long long orow[TC_BM];
    __shared__ int osamp[TC_BM];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rpf = a.Vin + a.ext_in;
    const int Fr = TC_BM / rpf;
    const int rows_tile = Fr * rpf;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const long long f0 = (long long)blockIdx.x * Fr;
    const int n0 = blockIdx.y * TC_BN;
    const int Nt = a.N - n0 < TC_BN ? a.N - n0 : TC_BN;    // live columns of this tile
    const int Ntp = (Nt + 15) & ~15;                       // MMA N (multiple of 16 for M=128)
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Ntp) tmem_cols <<= 1;
    FrameMap fm{a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};

    // shared memory carve-up: [rowsrc taps*128 x i64][A chunk][B chunk]; the fp32 staging tile re-uses A/B
    long long* rowsrc = reinterpret_cast<long long*>(smem);
    const int KVtot = a.taps * Kp;
    const int kpass = KVtot < TC_KPASS ? KVtot : TC_KPASS;
    unsigned char* Abase = smem + (size_t)a.taps * TC_BM * sizeof(long long);
    unsigned char* Bbase = Abase + (size_t)TC_BM * kpass * 2;
    float* Cs = reinterpret_cast<float*>(Abase);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    for (int idx = tid; idx < a.taps * TC_BM; idx += CG_THREADS) {
        int tap = idx / TC_BM, row = idx - tap * TC_BM;
        long long sr = -1;
        if (row < rows_tile) {
            int fl = row / rpf, j = row - fl * rpf;
            long long f = f0 + fl;
            if (f < n_frames) {
                long long sf = src_frame(fm, f, tap);
                if (sf >= 0) sr = (j < a.Vin) ? sf * a.Vin + j : -2;
            }
        }
        rowsrc[idx] = sr;
    }
    const int Vout = rpf - a.contract_ext;
    if (tid < TC_BM) {        // output-row table of the fused tail
        long long r = -1;
        int sm = 0;
        if (tid < Fr * Vout) {
            int fl = tid / Vout, j = tid - fl * Vout;
            long long f = f0 + fl;
            if (f < n_frames) { r = f * Vout + j; sm = (int)(f / a.T_out); }
        }
        orow[tid] = r;
        osamp[tid] = sm;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = wmode == 1 ? make_idesc_bmn(TC_BM, Ntp) : make_idesc(TC_BM, Ntp);

    uint32_t phase = 0;
    int first = 1;
    for (int kv0 = 0; kv0 < KVtot; kv0 += kpass) {
        const int kv_len = KVtot - kv0 < kpass ? KVtot - kv0 : kpass;
        const int nch = kv_len >> 3;                       // 16-byte chunks per row in this pass
        if (!first) mbar_wait(&mbar, phase ^ 1);           // previous pass's MMAs have finished reading A/B
        // ---- A operand: item = (row group of 8, 4 chunks) per warp-iteration; lane = (chunk%4)*8 + row%8
        const int groups4 = (nch + 3) >> 2;
        const int n_it = (TC_BM / 8) * groups4;
        constexpr int AB = 4;                               // items in flight per thread
        for (int it0 = warp; it0 < n_it; it0 += (CG_THREADS / 32) * AB) {
            Act8Raw raw[AB];
            long long srs[AB];
            int ks[AB], offs[AB];
#pragma unroll
            for (int b = 0; b < AB; ++b) {
                const int it = it0 + b * (CG_THREADS / 32);
                offs[b] = -1; srs[b] = -1; ks[b] = 0;
                if (it < n_it) {
                    const int rg = it / groups4, g4 = it - rg * groups4;
                    const int r = rg * 8 + (lane & 7), kc = g4 * 4 + (lane >> 3);
                    if (kc < nch) {
                        const int kv = kv0 + kc * 8;
                        const int tap = kv / Kp, k = kv - tap * Kp;
                        offs[b] = (int)op_off(r, kc, nch);
                        ks[b] = k;
                        const long long sr = rowsrc[tap * TC_BM + r];
                        if (sr >= 0 && k < a.K) {
                            srs[b] = sr;
                            if (vec_ok && k + 8 <= a.K) raw[b] = act8_issue(a.src, sr, k);
                        }
                    }
                }
            }
#pragma unroll
            for (int b = 0; b < AB; ++b) {
                if (offs[b] < 0) continue;
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (srs[b] >= 0) {
                    if (vec_ok && ks[b] + 8 <= a.K) pk = act8_finish(a.src, raw[b], ks[b]);
                    else pk = load_act8(a.src, srs[b], ks[b], a.K, 0);
                }
                *reinterpret_cast<uint4*>(Abase + offs[b]) = pk;
            }
        }
        // ---- B operand: weights fp32 -> bf16
        if (wmode == 0) {              // rows = output channels, 8 consecutive k per 16-byte item
            for (int idx = tid; idx < Ntp * nch; idx += CG_THREADS) {
                const int n = idx % Ntp, kc = idx / Ntp;
                const int k = kv0 + kc * 8;                 // taps == 1 on this path
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (n < Nt && k < a.K) {
                    const float* wp = a.W + (long long)(n0 + n) * a.ws_n + k;
                    float w[8];
                    if (k + 8 <= a.K) {
                        float4 x = *reinterpret_cast<const float4*>(wp), y = *reinterpret_cast<const float4*>(wp + 4);
                        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = (k + j < a.K) ? wp[j] : 0.f;
                    }
                    pk = pack8(w);
                }
                *reinterpret_cast<uint4*>(Bbase + op_off(n, kc, nch)) = pk;
            }
        } else if (wmode == 1) {       // MN-major: item = (k, 8 consecutive output channels)
            const int gN = Ntp >> 3;
            for (int idx = tid; idx < kv_len * gN; idx += CG_THREADS) {
                const int g8 = idx % gN, kk = idx / gN;
                const int k = kv0 + kk, n = g8 * 8;
                uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                if (k < a.K && n < Nt) {
                    const float* wp = a.W + (long long)k * a.ws_k + n0 + n;
                    float w[8];
                    if (n + 8 <= Nt) {
                        float4 x = *reinterpret_cast<const float4*>(wp), y = *reinterpret_cast<const float4*>(wp + 4);
                        w[0] = x.x; w[1] = x.y; w[2] = x.z; w[3] = x.w; w[4] = y.x; w[5] = y.y; w[6] = y.z; w[7] = y.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 8; ++j) w[j] = (n + j < Nt) ? wp[j] : 0.f;
                    }
                    pk = pack8(w);
                }
                *reinterpret_cast<uint4*>(Bbase + mn_off(g8, kk, gN)) = pk;
            }
        } else {
            for (int idx = tid; idx < Ntp * nch; idx += CG_THREADS) {
                const int n = idx % Ntp, kc = idx / Ntp;
                const int kv = kv0 + kc * 8;
                const int tap = kv / Kp, k = kv - tap * Kp;
                float w[8];
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    w[j] = (n < Nt && k + j < a.K)
                               ? a.W[(long long)(n0 + n) * a.ws_n + (long long)(k + j) * a.ws_k + (long long)tap * a.ws_tap] : 0.f;
                *reinterpret_cast<uint4*>(Bbase + op_off(n, kc, nch)) = pack8(w);
            }
        }
        if (a.ext_in) {
            __syncthreads();
            // joint-mean rows: one (frame, 8-channel chunk) per thread-iteration, averaged in fp32 over the V staged rows
            for (int idx = tid; idx < Fr * nch; idx += CG_THREADS) {
                const int kc = idx % nch, fl = idx / nch;
                const int tap = (kv0 + kc * 8) / Kp;
                const int mr = fl * rpf + a.Vin;
                if (rowsrc[tap * TC_BM + mr] != -2) continue;
                float s[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] = 0.f;
                for (int v = 0; v < a.Vin; ++v) {
                    float t[8];
                    unpack8(*reinterpret_cast<const uint4*>(Abase + op_off(fl * rpf + v, kc, nch)), t);
#pragma unroll
                    for (int j = 0; j < 8; ++j) s[j] += t[j];
                }
                const float inv = 1.f / (float)a.Vin;
#pragma unroll
                for (int j = 0; j < 8; ++j) s[j] *= inv;
                *reinterpret_cast<uint4*>(Abase + op_off(mr, kc, nch)) = pack8(s);
            }
        }
        // ---- make the generic-proxy writes visible to the tensor core, then one thread issues the MMAs
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t sbo = (uint32_t)nch * 128u;
            const uint32_t a0 = smem_u32(Abase), b0 = smem_u32(Bbase);
            const uint32_t gN = (uint32_t)(Ntp >> 3);
            for (int ks = 0; ks < (kv_len >> 4); ++ks) {
                const uint64_t ad = make_desc(a0 + ks * 256u, 128u, sbo);
                const uint64_t bd = wmode == 1 ? make_desc(b0 + ks * 2u * gN * 128u, gN * 128u, 128u)
                                               : make_desc(b0 + ks * 256u, 128u, sbo);
                umma_f16(tmem_d, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
            }
            umma_commit(&mbar);
        }
        first = 0;
        phase ^= 1;
    }
    mbar_wait(&mbar, phase ^ 1);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- TMEM -> registers -> fp32 staging tile (warp w owns lanes 32*(w%4).., column half w/4)
    {
        const int lq = warp & 3, half = warp >> 2;
        const int row = lq * 32 + lane;
        const int nc16 = Ntp >> 4;                                    // 16-column groups; warps 0-3 take the first half
        const int gbeg = half ? (nc16 + 1) >> 1 : 0, gend = half ? nc16 : (nc16 + 1) >> 1;
        for (int g = gbeg; g < gend; ++g) {
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(g * 16), v);
            float4* dst = reinterpret_cast<float4*>(Cs + row * TC_LDC + g * 16);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            dst[2] = make_float4(v[8], v[9], v[10], v[11]);
            dst[3] = make_float4(v[12], v[13], v[14], v[15]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
    if (vec_tail) gemm_tail_vec(a, Cs, orow, osamp, Fr * Vout, rpf, n0, s_red);
    else gemm_tail<bf16, TC_BN>(a, Cs, TC_LDC, s_red, f0, Fr, rpf, n0, n_frames);
}

// ------------------------------------------------------------------------------------------------------------
// Weight gradient on tcgen05: dW[n,k,tap] += sum_rows B[row,n] * A[row(tap),k].  The reduction axis is the row
// axis, so both operands are MN-major (channels contiguous): cute::UMMA Major::MN, SWIZZLE_NONE — 8(k) x 8(mn)
// core matrices of 128 B, SBO = 128 B between 8-channel groups, LBO = 128 B * groups between 8-row groups.
// D (TMEM, fp32) = [128 output channels] x [<=256 input channels], accumulated over all row passes of the CTA's
// frame range, then added to dW with atomics (one CTA per SM-slot => ~1e2 partial sums per weight).
constexpr int WT_ROWS = 128;       // rows (reduction length) per pass
constexpr int WT_BN = 128;         // output channels per CTA   (MMA M)
constexpr int WT_U = 4;            // independent operand loads in flight per thread while staging

DSG_D void red_add_v4(float* p, float x, float y, float z, float w) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
constexpr int WT_BK = 128;         // input channels per CTA    (MMA N)

DSG_D uint32_t make_idesc_mn(int M, int N) { return make_idesc(M, N) | (1u << 15) | (1u << 16); }

__global__ void __launch_bounds__(CG_THREADS, 3) conv_wgrad_tc_kernel(dsg_conv_wgrad_args a, int frames_per_cta, int vecA, int vecB, int vecW) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rpf = a.Vin + a.ext_in;
    const int Fr = WT_ROWS / rpf;
    const int rows_tile = Fr * rpf;
    const int rows_p = (rows_tile + 15) & ~15;                 // reduction length per pass (MMA K multiple of 16)
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const int ktiles = (a.K + WT_BK - 1) / WT_BK;
    const int kt = blockIdx.z % ktiles, tap = blockIdx.z / ktiles;
    const int k0 = kt * WT_BK, n0 = blockIdx.y * WT_BN;
    const int Kt = a.K - k0 < WT_BK ? a.K - k0 : WT_BK;        // live input channels of this tile
    const int Ktp = (Kt + 15) & ~15;
    const int Nt = a.N - n0 < WT_BN ? a.N - n0 : WT_BN;
    const int gM = WT_BN / 8, gN = Ktp / 8;                    // channel groups of the two operands
    uint32_t tmem_cols = 32;
    while ((int)tmem_cols < Ktp) tmem_cols <<= 1;
    const long long fbeg = (long long)blockIdx.x * frames_per_cta;
    long long fend = fbeg + frames_per_cta;
    if (fend > n_frames) fend = n_frames;
    FrameMap fm{a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};
    const bool do_bias = (a.db != nullptr) && kt == 0 && tap == 0;
    const bool fastA = vecA && (a.K % 8 == 0), fastB = vecB && (a.N % 8 == 0);

    __shared__ __align__(16) float cfA[3][WT_BK], cfB[3][WT_BN];      // staged per-channel coefficients (a1, b1+b2, a2) of both operands
    for (int i = tid; i < WT_BK; i += CG_THREADS) {
        const int ch = k0 + i;
        const bool in = ch < a.K;
        cfA[0][i] = (in && a.A.a1) ? a.A.a1[ch] : 1.f;
        cfA[1][i] = ((in && a.A.b1) ? a.A.b1[ch] : 0.f) + ((in && a.A.b2) ? a.A.b2[ch] : 0.f);
        cfA[2][i] = (in && a.A.a2) ? a.A.a2[ch] : 1.f;
    }
    for (int i = tid; i < WT_BN; i += CG_THREADS) {
        const int ch = n0 + i;
        const bool in = ch < a.N;
        cfB[0][i] = (in && a.B.a1) ? a.B.a1[ch] : 1.f;
        cfB[1][i] = ((in && a.B.b1) ? a.B.b1[ch] : 0.f) + ((in && a.B.b2) ? a.B.b2[ch] : 0.f);
        cfB[2][i] = (in && a.B.a2) ? a.B.a2[ch] : 1.f;
    }
    long long* rowsrc = reinterpret_cast<long long*>(smem);            // [WT_ROWS]
    long long* rowdst = rowsrc + WT_ROWS;                              // [WT_ROWS]
    unsigned char* Mop = smem + 2 * WT_ROWS * sizeof(long long);      // B side: [128 ch][rows_p]
    unsigned char* Nop = Mop + (size_t)WT_BN * WT_ROWS * 2;           // A side: [Ktp ch][rows_p]

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;
    const uint32_t idesc = make_idesc_mn(WT_BN, Ktp);

    uint32_t phase = 0;
    int first = 1;
    float bs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // bias-gradient partials of this thread's 8 output channels
    for (long long f0 = fbeg; f0 < fend; f0 += Fr) {
        if (!first) mbar_wait(&mbar, phase ^ 1);               // MMAs of the previous pass are done with Mop/Nop
        if (tid < WT_ROWS) {
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
        // ---- stage both operands: lane = (row % 8) + 8 * (channel group % 4).  Fast path (16-byte loads, channel counts
        //      that are multiples of 8): WT_U independent loads in flight per thread.  The B side has 16 channel groups, so
        //      a thread keeps the same group on every iteration and the bias gradient accumulates in registers.
        {
            const int g8 = (warp & 3) * 4 + (lane >> 3);
            const int cB = n0 + g8 * 8;
            const int nI = rows_p >> 4;                        // iterations per warp: row group (warp >> 2) + 2 * i
            if (fastB) {
                for (int i0 = 0; i0 < nI; i0 += WT_U) {
                    Act8Raw q[WT_U];
                    bool ok[WT_U];
#pragma unroll
                    for (int u = 0; u < WT_U; ++u) {
                        const int kk = ((warp >> 2) + 2 * (i0 + u)) * 8 + (lane & 7);
                        ok[u] = false;
                        q[u].a = q[u].b = make_uint4(0u, 0u, 0u, 0u);
                        if (i0 + u < nI) {
                            const long long dr = rowdst[kk];
                            ok[u] = kk < rows_tile && dr >= 0 && cB < a.N;
                            if (ok[u]) q[u] = act8_issue(a.B, dr, cB);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < WT_U; ++u) {
                        const int kk = ((warp >> 2) + 2 * (i0 + u)) * 8 + (lane & 7);
                        if (i0 + u < nI) {
                            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                            if (ok[u]) {
                                float v[8];
                                finish_smem(q[u], a.B.x2 != nullptr, a.B.relu, &cfB[0][g8 * 8], &cfB[1][g8 * 8], &cfB[2][g8 * 8], v);
                                pk = pack8(v);
                                if (do_bias) {
                                    unpack8(pk, v);
#pragma unroll
                                    for (int j = 0; j < 8; ++j) bs[j] += v[j];
                                }
                            }
                            *reinterpret_cast<uint4*>(Mop + mn_off(g8, kk, gM)) = pk;
                        }
                    }
                }
            } else {
                for (int i = 0; i < nI; ++i) {
                    const int kk = ((warp >> 2) + 2 * i) * 8 + (lane & 7);
                    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                    const long long dr = rowdst[kk];
                    if (kk < rows_tile && dr >= 0 && cB < a.N) {
                        pk = load_act8(a.B, dr, cB, a.N, 0);
                        if (do_bias) {
                            float v[8];
                            unpack8(pk, v);
#pragma unroll
                            for (int j = 0; j < 8; ++j) bs[j] += v[j];
                        }
                    }
                    *reinterpret_cast<uint4*>(Mop + mn_off(g8, kk, gM)) = pk;
                }
            }
        }
        {
            const int g4n = (gN + 3) / 4;
            const int total = (rows_p / 8) * g4n;
            if (fastA) {
                for (int it0 = warp; it0 < total; it0 += (CG_THREADS / 32) * WT_U) {
                    Act8Raw q[WT_U];
                    bool ok[WT_U];
#pragma unroll
                    for (int u = 0; u < WT_U; ++u) {
                        const int it = it0 + u * (CG_THREADS / 32);
                        const int rg = it / g4n, g8 = (it - rg * g4n) * 4 + (lane >> 3);
                        const int kk = rg * 8 + (lane & 7);
                        ok[u] = false;
                        q[u].a = q[u].b = make_uint4(0u, 0u, 0u, 0u);
                        if (it < total) {
                            const long long sr = rowsrc[kk];
                            ok[u] = g8 < gN && kk < rows_tile && sr >= 0 && k0 + g8 * 8 < a.K;
                            if (ok[u]) q[u] = act8_issue(a.A, sr, k0 + g8 * 8);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < WT_U; ++u) {
                        const int it = it0 + u * (CG_THREADS / 32);
                        const int rg = it / g4n, g8 = (it - rg * g4n) * 4 + (lane >> 3);
                        const int kk = rg * 8 + (lane & 7);
                        if (it < total && g8 < gN) {
                            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                            if (ok[u]) {
                                float v[8];
                                finish_smem(q[u], a.A.x2 != nullptr, a.A.relu, &cfA[0][g8 * 8], &cfA[1][g8 * 8], &cfA[2][g8 * 8], v);
                                pk = pack8(v);
                            }
                            *reinterpret_cast<uint4*>(Nop + mn_off(g8, kk, gN)) = pk;
                        }
                    }
                }
            } else {
                for (int it = warp; it < total; it += CG_THREADS / 32) {
                    const int rg = it / g4n, g8 = (it - rg * g4n) * 4 + (lane >> 3);
                    const int kk = rg * 8 + (lane & 7);
                    if (g8 >= gN) continue;
                    uint4 pk = make_uint4(0u, 0u, 0u, 0u);
                    const long long sr = rowsrc[kk];
                    if (kk < rows_tile && sr >= 0 && k0 + g8 * 8 < a.K) pk = load_act8(a.A, sr, k0 + g8 * 8, a.K, 0);
                    *reinterpret_cast<uint4*>(Nop + mn_off(g8, kk, gN)) = pk;
                }
            }
        }
        if (a.ext_in) {
            // joint-mean row of every frame, from the staged tile: one thread per (frame, 8-channel group)
            __syncthreads();
            for (int idx = tid; idx < Fr * gN; idx += CG_THREADS) {
                const int g = idx % gN, fl = idx / gN;
                const int mr = fl * rpf + a.Vin;
                if (rowsrc[mr] != -2) continue;
                float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                for (int j = 0; j < a.Vin; ++j) {
                    float v[8];
                    unpack8(*reinterpret_cast<const uint4*>(Nop + mn_off(g, fl * rpf + j, gN)), v);
#pragma unroll
                    for (int c = 0; c < 8; ++c) s[c] += v[c];
                }
                const float inv = 1.f / (float)a.Vin;
#pragma unroll
                for (int c = 0; c < 8; ++c) s[c] *= inv;
                *reinterpret_cast<uint4*>(Nop + mn_off(g, mr, gN)) = pack8(s);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t m0 = smem_u32(Mop), q0 = smem_u32(Nop);
            for (int ks = 0; ks < (rows_p >> 4); ++ks) {
                // MN-major: SBO = distance between 8-channel groups (128 B), LBO = distance between 8-row groups
                const uint64_t ad = make_desc(m0 + ks * 2u * gM * 128u, (uint32_t)gM * 128u, 128u);
                const uint64_t bd = make_desc(q0 + ks * 2u * gN * 128u, (uint32_t)gN * 128u, 128u);
                umma_f16(tmem_d, ad, bd, idesc, (first && ks == 0) ? 0u : 1u);
            }
            umma_commit(&mbar);
        }
        first = 0;
        phase ^= 1;
    }
    if (!first) {
        mbar_wait(&mbar, phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int lq = warp & 3, half = warp >> 2;
        const int n = lq * 32 + lane;
        const int nc16 = Ktp >> 4;
        const int gbeg = half ? (nc16 + 1) >> 1 : 0, gend = half ? nc16 : (nc16 + 1) >> 1;
        for (int g = gbeg; g < gend; ++g) {
            float v[16];
            tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(g * 16), v);
            if (n < Nt) {
                float* dst = a.dW + (long long)(n0 + n) * a.ws_n + (long long)(k0 + g * 16) * a.ws_k + (long long)tap * a.ws_tap;
                if (vecW && g * 16 + 16 <= Kt) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4) red_add_v4(dst + j, v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (g * 16 + j < Kt) atomicAdd(dst + (long long)j * a.ws_k, v[j]);
                }
            }
        }
        if (do_bias) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 1);
                bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 2);
                bs[j] += __shfl_xor_sync(0xffffffffu, bs[j], 4);
            }
            const int g8 = (warp & 3) * 4 + (lane >> 3);
            if ((lane & 7) == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (g8 * 8 + j < Nt) atomicAdd(a.db + n0 + g8 * 8 + j, bs[j]);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, tmem_cols);
}

static const char* launch_conv_wgrad_tc(const dsg_conv_wgrad_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    int rpf = a.Vin + a.ext_in;
    if (a.dtype != DSG_BF16 || rpf > WT_ROWS) return nullptr;
    long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0) { *handled = true; return nullptr; }
    auto vec = [](const ActSrc& s) { return (int)act8_ok(s); };
    int ktiles = (a.K + WT_BK - 1) / WT_BK, ntiles = (a.N + WT_BN - 1) / WT_BN;
    int Fr = WT_ROWS / rpf;
    long long per = (long long)ktiles * ntiles * a.taps;
    int Ktp = a.K < WT_BK ? (a.K + 15) & ~15 : WT_BK;
    size_t smem = 2 * WT_ROWS * sizeof(long long) + (size_t)(WT_BN + Ktp) * WT_ROWS * 2;
    int occ = (int)((200 * 1024) / (smem + 4096));                // CTAs one SM holds (TMEM: <=128 columns each)
    if (occ > 3) occ = 3;                                         // 80 registers x 256 threads: three CTAs per SM
    if (occ < 1) occ = 1;
    long long want = ((long long)occ * dsg_num_sms() + per - 1) / per;      // one wave of resident CTAs
    long long fpc = (n_frames + want - 1) / want;
    fpc = (fpc + Fr - 1) / Fr * Fr;
    if (fpc < Fr) fpc = Fr;
    int vecW = a.ws_k == 1 && a.ws_n % 4 == 0 && a.ws_tap % 4 == 0 && (uintptr_t)a.dW % 16 == 0;
    dim3 grid((unsigned)((n_frames + fpc - 1) / fpc), (unsigned)ntiles, (unsigned)(ktiles * a.taps));
    cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_wgrad_tc_kernel<<<grid, dim3(CG_THREADS), smem, st>>>(a, (int)fpc, vec(a.A), vec(a.B), vecW);
    *handled = true;
    return dsg_launch_error();
}

static inline size_t tc_smem_bytes(int taps, int Kp) {
    int KVtot = taps * Kp;
    int kpass = KVtot < TC_KPASS ? KVtot : TC_KPASS;
    size_t ops = (size_t)2 * TC_BM * kpass * 2;
    size_t stage = (size_t)TC_BM * TC_LDC * sizeof(float);
    return (size_t)taps * TC_BM * sizeof(long long) + (ops > stage ? ops : stage);
}

// returns nullptr on success, an error string on failure, or the sentinel "" when the shape is not eligible
static const char* launch_conv_gemm_tc(const dsg_conv_gemm_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    int rpf = a.Vin + a.ext_in;
    if (a.dtype != DSG_BF16 || rpf > TC_BM || a.taps > TC_MAX_TAPS) return nullptr;
    if (a.contract_ext && (a.ext_in || a.Vin < 2)) return "conv_gemm: contract_ext needs Vin>=2 rows and no ext_in";
    long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0 || a.N <= 0) { *handled = true; return nullptr; }
    int Kp = (a.K + 15) & ~15;
    int vec_ok = act8_ok(a.src);
    int wmode = 2;
    if (a.taps == 1 && (uintptr_t)a.W % 16 == 0) {
        if (a.ws_k == 1 && a.ws_n % 4 == 0) wmode = 0;
        else if (a.ws_n == 1 && a.ws_k % 4 == 0) wmode = 1;
    }
    auto al16 = [](const void* p, long long ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 8 == 0); };
    int vec_tail = (a.N % 8 == 0) && al16(a.out, a.ld_out) && al16(a.add, a.ld_add) && al16(a.add2, a.ld_add2) &&
                   al16(a.partner, a.ld_partner) && (!a.has_mask || act8_ok(a.mask)) &&
                   (!a.bias || (uintptr_t)a.bias % 16 == 0) && (!a.bcast || (uintptr_t)a.bcast % 16 == 0);
    size_t smem = tc_smem_bytes(a.taps, Kp);
    if (smem > 200 * 1024) return nullptr;
    int Fr = TC_BM / rpf;
    dim3 grid((unsigned)((n_frames + Fr - 1) / Fr), (unsigned)((a.N + TC_BN - 1) / TC_BN));
    cudaFuncSetAttribute(conv_gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    conv_gemm_tc_kernel<<<grid, dim3(CG_THREADS), smem, st>>>(a, Kp, vec_ok, wmode, vec_tail);
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc
}  // namespace dsg
#endif  // !DSG_EMU
