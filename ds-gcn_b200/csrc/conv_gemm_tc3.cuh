// Persistent tcgen05 engine for the hot 1x1 convolutions (same shapes as conv_gemm_tc2_kernel, packed weights required).
// One CTA = 256 threads = two groups of 128; two shared-memory stages, two TMEM accumulators, weights resident.
// Default schedule (pp = 1, "ping-pong"): each group produces AND drains its own tiles (tile i -> group i & 1), the two
// groups running out of phase, so the loads of one tile overlap the MMAs and the tail of the other and no warp idles
// whichever side is the bottleneck (measured 1 % better over the step than fixed roles; DSG_TC3_PP=0 selects those):
//
//   producers (warps 0-3)  row-per-thread: load the thread's own input row of tile i (8 independent 16-byte loads in
//                          flight), fused prologue, bf16 chunks into stage i&1 of the K-major no-swizzle UMMA tile;
//                          thread 0 then issues the tile's MMAs (the whole reduction axis: the weight tile of this
//                          CTA's 128 output columns arrives ONCE by bulk copy and stays resident) into accumulator i&1
//   epilogue  (warps 4-7)  row-per-thread: tcgen05.ld of the thread's own accumulator row, fused tail, 16-byte stores,
//                          BatchNorm statistics kept per warp in shared memory across all tiles of the CTA
//
// so the loads of tile i+1 overlap the MMAs and the tail of tile i.  mbarriers: a_free[s] (MMAs done reading stage s),
// acc_full[b] (accumulator b complete), acc_free[b] (128 epilogue threads done reading accumulator b), wbar (weights).
// Per-CTA setup (coefficients, weights, TMEM, barriers) is paid once per ~tiles/296 instead of once per tile.
#pragma once
#include "conv_gemm_tc2.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc {

constexpr int T3_THREADS = 256;

DSG_D void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
DSG_D void group_sync(int id) {               // named barrier over one 128-thread role group
    asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory");
}

// K-major SWIZZLE_128B operand: rows of 128 bytes (64 bf16 of the reduction axis), 8-row groups 1024 bytes apart, the
// 16-byte chunk index XOR-ed with (row & 7); one "atom" = 128 rows x 64 channels = 16 KB, 1024-byte aligned.
DSG_D uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(1024u >> 4) << 32;                   // SBO: next 8-row group
    d |= (uint64_t)1 << 46;                              // descriptor version (sm_100)
    d |= (uint64_t)2 << 61;                              // layout type SWIZZLE_128B
    return d;
}
DSG_D uint32_t sw128_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// Coalesced staging of one 64-channel atom: lane = (row % 4, chunk) so a warp reads 4 whole 128-byte rows per load and
// writes 4 whole swizzled rows per store (conflict-free both ways); NB loads in flight per source.
template <int NB, bool DUAL>
DSG_D void t3_stage_sw128(const dsg_conv_gemm_args& a, const long long* rowsrc, int rtid, int k0, unsigned char* atom,
                          const float* cf_a1, const float* cf_b, const float* cf_a2) {
    const int chunk = rtid & 7, r0 = rtid >> 3;          // rows r0, r0 + 16, ...
    const int k = k0 + chunk * 8;
    const bf16* x1 = reinterpret_cast<const bf16*>(a.src.x1) + k;
    const bf16* x2 = DUAL ? reinterpret_cast<const bf16*>(a.src.x2) + k : nullptr;
    const bool kin = k < a.K;
    for (int i0 = 0; i0 < 8; i0 += NB) {
        Act8Raw raw[NB];
        long long sr[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            sr[b] = rowsrc[(i0 + b) * 16 + r0];
            raw[b].a = raw[b].b = make_uint4(0u, 0u, 0u, 0u);
            if (sr[b] >= 0 && kin) {
                raw[b].a = *reinterpret_cast<const uint4*>(x1 + sr[b] * a.src.ld1);
                if (DUAL) raw[b].b = *reinterpret_cast<const uint4*>(x2 + sr[b] * a.src.ld2);
            }
        }
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const int row = (i0 + b) * 16 + r0;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (sr[b] >= 0 && kin) {
                float v[8];
                finish_smem(raw[b], DUAL, a.src.relu, cf_a1 + k, cf_b + k, cf_a2 + k, v);
                pk = pack8(v);
            }
            *reinterpret_cast<uint4*>(atom + sw128_off(row, chunk)) = pk;
        }
    }
}

__global__ void __launch_bounds__(T3_THREADS) conv_gemm_tc3_kernel(dsg_conv_gemm_args a, int n_tiles, int pp) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t a_free[2], acc_full[2], acc_free[2], wbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float ext_sg[2][2][8][16];                // per group: joint-mean accumulator rows of up to 8 frames, double-buffered
    __shared__ float s_acc[2][8][T2_BN];                 // per-warp column sums, accumulated over all tiles of this CTA
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int role = tid >> 7, rtid = tid & 127;         // 0: producer, 1: epilogue; rtid = row of the tile = TMEM lane
    const int rpf = a.Vin + a.ext_in;
    const int Fr = 128 / rpf;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    const int n0 = blockIdx.y * T2_BN;
    const int Nt = a.N - n0 < T2_BN ? a.N - n0 : T2_BN;
    const int Ntp = (Nt + 15) & ~15;
    uint32_t acc_cols = 32;
    while ((int)acc_cols < Ntp) acc_cols <<= 1;
    const int Kp = (a.K + 15) & ~15;
    const int kpass = Kp < T2_KPASS ? Kp : T2_KPASS;
    const size_t a_bytes = (size_t)128 * Kp * 2;         // one stage: the pass tiles back to back
    __shared__ long long rowsrc_s[2][128];               // per group: source row of every tile row (-1 zero row, -2 joint-mean row)
    const bool sw = (Kp & 63) == 0;                      // reduction axis in whole 64-channel atoms: coalesced swizzled staging
    unsigned char* smem_al = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);      // SWIZZLE_128B atoms need 1024-byte alignment
    unsigned char* Wbase = smem_al;                      // [pass][Ntp x kv_len]
    unsigned char* Abase0 = smem_al + (size_t)128 * Kp * 2;
    float* cf_a1 = reinterpret_cast<float*>(Abase0 + 2 * a_bytes);
    float* cf_b = cf_a1 + Kp;
    float* cf_a2 = cf_b + Kp;
    float* tl_bias = cf_a2 + Kp;
    float* tl_ma1 = tl_bias + T2_BN;
    float* tl_mb = tl_ma1 + T2_BN;
    float* tl_ma2 = tl_mb + T2_BN;
    for (int k = tid; k < Kp; k += T3_THREADS) {
        const bool in = k < a.K;
        cf_a1[k] = (in && a.src.a1) ? a.src.a1[k] : 1.f;
        cf_b[k] = ((in && a.src.b1) ? a.src.b1[k] : 0.f) + ((in && a.src.b2) ? a.src.b2[k] : 0.f);
        cf_a2[k] = (in && a.src.a2) ? a.src.a2[k] : 1.f;
    }
    if (tid < T2_BN) {
        const int cch = n0 + tid;
        const bool in = cch < a.N;
        tl_bias[tid] = (in && a.bias) ? a.bias[cch] : 0.f;
        tl_ma1[tid] = (in && a.has_mask && a.mask.a1) ? a.mask.a1[cch] : 1.f;
        tl_mb[tid] = ((in && a.has_mask && a.mask.b1) ? a.mask.b1[cch] : 0.f) + ((in && a.has_mask && a.mask.b2) ? a.mask.b2[cch] : 0.f);
        tl_ma2[tid] = (in && a.has_mask && a.mask.a2) ? a.mask.a2[cch] : 1.f;
    }
    for (int i = tid; i < 2 * 8 * T2_BN; i += T3_THREADS) (&s_acc[0][0][0])[i] = 0.f;

    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&a_free[i], 1);
            mbar_init(&acc_full[i], 1);
            mbar_init(&acc_free[i], 128);
        }
        mbar_init(&wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 2 * acc_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // this thread's position inside every tile
    const int fl = rtid / rpf, j = rtid - fl * rpf;
    const int Vout = rpf - a.contract_ext;
    const bool is_ext_row = a.contract_ext && j == rpf - 1;

    // pp == 0: fixed roles (group 0 produces every tile, group 1 drains it).  pp == 1: ping-pong — each group produces
    // AND drains its own tiles (it = group, group + 2, ...), the two groups running out of phase, so no warp idles when
    // one side of the pipeline is the bottleneck.  Stage / accumulator index it & 1 and use count it >> 1 hold for both.
    float (*ext_s)[8][16] = ext_sg[role];
    const bool do_prod = pp || role == 0, do_epi = pp || role == 1;
    const int bar_p = pp ? 1 + role : 1, bar_e = pp ? 3 + role : 2;
    {
        if (tid == 0) {                                   // the CTA's weight tile, all passes: one bulk copy
            const uint32_t bytes = (uint32_t)(Ntp * Kp * 2);
            mbar_expect_tx(&wbar, bytes);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.wpack) + (long long)blockIdx.y * 128 * Kp * 2;
            if (Ntp == 128 || Kp <= kpass) bulk_g2s(Wbase, src, bytes, &wbar);       // pass tiles are contiguous in wpack
            else {
                uint32_t done = 0;
                for (int kv0 = 0; kv0 < Kp; kv0 += kpass) {                          // narrower last column tile: one copy per pass
                    const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass;
                    const uint32_t b = (uint32_t)(Ntp * kv_len * 2);
                    bulk_g2s(Wbase + done, src + 128LL * kv0 * 2, b, &wbar);
                    done += b;
                }
            }
        }
    }
    const uint32_t idesc = make_idesc(128, Ntp);
    FrameMap fm{1, a.tap_step, a.tap_off, a.t_mul, a.t_div, a.T_in, a.T_out, a.Vin, a.ext_in};
    const int ew = warp & 3;
    const float inv_ext = a.contract_ext ? 1.f / (float)(rpf - 1) : 0.f;
    bf16* out = reinterpret_cast<bf16*>(a.out);
    const bf16* addp = reinterpret_cast<const bf16*>(a.add);
    const bf16* add2p = reinterpret_cast<const bf16*>(a.add2);
    const bf16* partp = reinterpret_cast<const bf16*>(a.partner);
    const int it0 = pp ? role : 0, its = pp ? 2 : 1;
    for (int it = it0; ; it += its) {
        const long long tile_ll = (long long)blockIdx.x + (long long)it * gridDim.x;
        if (tile_ll >= n_tiles) break;
        const int tile = (int)tile_ll;
        const int st = it & 1, use = it >> 1;
        if (do_prod) {
            // ================================ produce ================================
            unsigned char* Abase = Abase0 + st * a_bytes;
            const long long f = (long long)tile * Fr + fl;
            long long sr = -1;                               // source row (-2: joint-mean row)
            if (fl < Fr && f < n_frames) {
                const long long sf = src_frame(fm, f, 0);
                if (sf >= 0) sr = (j < a.Vin) ? sf * a.Vin + j : -2;
            }
            if (use > 0) mbar_wait(&a_free[st], (uint32_t)((use - 1) & 1));     // the MMAs that read this stage are done
            if (sw) {
                rowsrc_s[role][rtid] = sr;
                group_sync(bar_p);
                for (int k0 = 0; k0 < Kp; k0 += 64) {
                    unsigned char* atom = Abase + (size_t)128 * k0 * 2;
                    if (a.src.x2 == nullptr) t3_stage_sw128<8, false>(a, rowsrc_s[role], rtid, k0, atom, cf_a1, cf_b, cf_a2);
                    else t3_stage_sw128<4, true>(a, rowsrc_s[role], rtid, k0, atom, cf_a1, cf_b, cf_a2);
                }
            } else {
                for (int kv0 = 0; kv0 < Kp; kv0 += kpass) {
                    const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass;
                    unsigned char* Ap = Abase + (size_t)128 * kv0 * 2;
                    if (a.src.x2 == nullptr) t2_stage_a<8, false>(a, sr, rtid, kv0, kv_len >> 3, Ap, cf_a1, cf_b, cf_a2);
                    else t2_stage_a<4, true>(a, sr, rtid, kv0, kv_len >> 3, Ap, cf_a1, cf_b, cf_a2);
                }
            }
            if (a.ext_in && sw) {
                group_sync(bar_p);
                const int nchT = Kp >> 3;
                for (int idx = rtid; idx < Fr * nchT; idx += 128) {        // joint-mean rows, averaged in fp32
                    const int kcT = idx % nchT, ff = idx / nchT;
                    if ((long long)tile * Fr + ff >= n_frames) continue;
                    unsigned char* atom = Abase + (size_t)128 * ((kcT >> 3) * 64) * 2;
                    const int kc = kcT & 7;
                    float s8[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] = 0.f;
                    for (int v = 0; v < a.Vin; ++v) {
                        float t[8];
                        unpack8(*reinterpret_cast<const uint4*>(atom + sw128_off(ff * rpf + v, kc)), t);
#pragma unroll
                        for (int e = 0; e < 8; ++e) s8[e] += t[e];
                    }
                    const float inv = 1.f / (float)a.Vin;
#pragma unroll
                    for (int e = 0; e < 8; ++e) s8[e] *= inv;
                    *reinterpret_cast<uint4*>(atom + sw128_off(ff * rpf + a.Vin, kc)) = pack8(s8);
                }
            } else if (a.ext_in) {
                group_sync(bar_p);
                const int nchT = Kp >> 3;
                for (int idx = rtid; idx < Fr * nchT; idx += 128) {        // joint-mean rows, averaged in fp32
                    const int kcT = idx % nchT, ff = idx / nchT;
                    if ((long long)tile * Fr + ff >= n_frames) continue;
                    const int kv0 = (kcT * 8 / kpass) * kpass;
                    const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass;
                    const int kc = kcT - (kv0 >> 3), nch = kv_len >> 3;
                    unsigned char* Ap = Abase + (size_t)128 * kv0 * 2;
                    float s[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) s[e] = 0.f;
                    for (int v = 0; v < a.Vin; ++v) {
                        float t[8];
                        unpack8(*reinterpret_cast<const uint4*>(Ap + op_off(ff * rpf + v, kc, nch)), t);
#pragma unroll
                        for (int e = 0; e < 8; ++e) s[e] += t[e];
                    }
                    const float inv = 1.f / (float)a.Vin;
#pragma unroll
                    for (int e = 0; e < 8; ++e) s[e] *= inv;
                    *reinterpret_cast<uint4*>(Ap + op_off(ff * rpf + a.Vin, kc, nch)) = pack8(s);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            group_sync(bar_p);
            if (rtid == 0) {
                if (use == 0) mbar_wait(&wbar, 0);
                if (use > 0) mbar_wait(&acc_free[st], (uint32_t)((use - 1) & 1));   // the epilogue has drained this accumulator
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_d + (uint32_t)st * acc_cols;
                uint32_t woff = 0;
                int firstmma = 1;
                for (int kv0 = 0; kv0 < Kp; kv0 += kpass) {
                    const int kv_len = Kp - kv0 < kpass ? Kp - kv0 : kpass;
                    const uint32_t sbo = (uint32_t)(kv_len >> 3) * 128u;
                    const uint32_t a0 = smem_u32(Abase + (size_t)128 * kv0 * 2), b0 = smem_u32(Wbase + woff);
                    for (int ks = 0; ks < (kv_len >> 4); ++ks) {
                        // swizzled A: atom (ks >> 2) of this pass, 32 bytes per K = 16 step inside the atom's 128-byte rows
                        const uint64_t ad = sw ? make_desc_sw128(a0 + (uint32_t)(ks >> 2) * 16384u + (uint32_t)(ks & 3) * 32u)
                                               : make_desc(a0 + ks * 256u, 128u, sbo);
                        umma_f16(acc, ad, make_desc(b0 + ks * 256u, 128u, sbo), idesc, firstmma ? 0u : 1u);
                        firstmma = 0;
                    }
                    woff += (uint32_t)(Ntp * kv_len * 2);
                }
                umma_commit(&a_free[st]);
                umma_commit(&acc_full[st]);
            }
        }
        if (do_epi) {
            // ================================ drain ================================
            const long long f = (long long)tile * Fr + fl;
            const bool row_ok = fl < Fr && f < n_frames;
            const long long orow = (row_ok && !is_ext_row) ? f * Vout + j : -1;
            const int samp = row_ok ? (int)(f / a.T_out) : 0;
            mbar_wait(&acc_full[st], (uint32_t)(use & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t acc = tmem_d + (uint32_t)st * acc_cols;
            for (int c16 = 0; c16 < Ntp; c16 += 16) {
                const int c = n0 + c16;
                const bool live0 = c < a.N, live1 = c + 8 < a.N;             // N % 8 == 0: the chunk is live in halves
                // this row's tail operands first: they travel while the accumulator chunk is read
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
                float v[16];
                tmem_ld16(acc + ((uint32_t)(ew * 32) << 16) + (uint32_t)c16, v);
                if (a.contract_ext) {                                           // fold the joint-mean row back into its frame
                    const int buf = (c16 >> 4) & 1;
                    if (row_ok && is_ext_row && fl < 8) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) ext_s[buf][fl][e] = v[e];
                    }
                    group_sync(bar_e);
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
                        s_acc[0][warp][c16 + (lane >> 1)] += t1;
                        s_acc[1][warp][c16 + (lane >> 1)] += t2;
                    }
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acc_free[st]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, 2 * acc_cols);
    if (a.stat_sum && tid < Nt) {
        float t1 = 0.f, t2 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { t1 += s_acc[0][w][tid]; t2 += s_acc[1][w][tid]; }
        atomicAdd(a.stat_sum + n0 + tid, (double)t1);
        atomicAdd(a.stat_sq + n0 + tid, (double)t2);
    }
}

static const char* launch_conv_gemm_tc3(const dsg_conv_gemm_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    const int rpf = a.Vin + a.ext_in;
    if (a.dtype != DSG_BF16 || a.taps != 1 || rpf > 128 || a.K % 8 != 0 || a.N % 8 != 0) return nullptr;
    if (a.contract_ext && (a.ext_in || a.Vin < 2 || 128 / rpf > 8)) return nullptr;
    if (!act8_ok(a.src) || !a.wpack || (uintptr_t)a.wpack % 128 != 0) return nullptr;
    auto al16 = [](const void* p, long long ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 8 == 0); };
    if (!(al16(a.out, a.ld_out) && al16(a.add, a.ld_add) && al16(a.add2, a.ld_add2) && al16(a.partner, a.ld_partner) &&
          (!a.has_mask || act8_ok(a.mask)) && (!a.bias || (uintptr_t)a.bias % 16 == 0) && (!a.bcast || (uintptr_t)a.bcast % 16 == 0)))
        return nullptr;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0 || a.N <= 0) { *handled = true; return nullptr; }
    const int Kp = (a.K + 15) & ~15;
    const int kpass = Kp < T2_KPASS ? Kp : T2_KPASS;
    const size_t smem = (size_t)3 * 128 * Kp * 2 + (size_t)(3 * Kp + 4 * T2_BN) * sizeof(float) + 1024;   // + alignment slack
    if (smem > 200 * 1024) return nullptr;
    // Measured on B200 (profiles/): with one producer group per CTA this engine wins while the reduction axis is short
    // (Kp <= 192: the weight tile and two stages leave room for two CTAs per SM or the tail dominates) and the CTA's
    // columns are the whole output (one column tile) or the rows are cheap to re-read (Kp <= 96); wider shapes are
    // load-bound on the producers and run faster on conv_gemm_tc2_kernel (3 CTAs per SM, every thread loads).
    static const int all_shapes = [] { const char* e = getenv("DSG_TC3_ALL"); return (e && e[0] == '1') ? 1 : 0; }();      // experiments
    if (!all_shapes && (Kp > 192 || (a.N > T2_BN && Kp > 96))) return nullptr;
    const int Fr = 128 / rpf;
    const long long tiles = (n_frames + Fr - 1) / Fr;
    if (tiles > 0x7fffffff) return nullptr;
    int per_sm = (int)((220 * 1024) / (smem + 8 * 1024));
    if (per_sm > 2) per_sm = 2;                          // 2 x 256 TMEM columns, 2 x 256 threads x ~120 registers
    if (per_sm < 1) per_sm = 1;
    const unsigned ny = (unsigned)((a.N + T2_BN - 1) / T2_BN);
    long long gx = ((long long)per_sm * dsg_num_sms() + ny - 1) / ny;
    if (gx > tiles) gx = tiles;
    if (gx < 1) gx = 1;
    conv_wpack_kernel<<<dim3(ny, (unsigned)((Kp + kpass - 1) / kpass)), dim3(256), 0, st>>>(a.W, a.ws_n, a.ws_k, a.K, a.N, reinterpret_cast<unsigned char*>(a.wpack));
    if (const char* e = dsg_launch_error()) return e;
    cudaFuncSetAttribute(conv_gemm_tc3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    static const int pp = [] { const char* e = getenv("DSG_TC3_PP"); return (e && e[0] == '0') ? 0 : 1; }();             // 0: fixed roles
    conv_gemm_tc3_kernel<<<dim3((unsigned)gx, ny), dim3(T3_THREADS), smem, st>>>(a, (int)tiles, pp);
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc
}  // namespace dsg
#endif
