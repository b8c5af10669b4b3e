// tc4 temporal weight gradient: the weight / bias gradients of ALL dilated (3 x 1) conv branches of a multi-scale temporal unit
// (tcn.py:383-391, backward) on the TMA-fed tcgen05 engine — the backward twin of tc4_tconv.cuh.
//
//   dW_b[co, ci, tap] = sum_{n, t', r} dO[n, t', r, lo_b + co] * H[n, s*t' + (tap-1)*d_b, r, lo_b + ci]      (zero padding in t)
//   db_b[co]          = sum_{n, t', r} dO[n, t', r, lo_b + co]
//
// H = relu(bn(B)) (the forward input of the convs) and dO (the gradient w.r.t. every branch output, joint-mean row included)
// are both plain bf16 tensors the tap-shifted path has materialised anyway, so nothing is staged by a thread: the reduction
// runs over ROWS, so the atoms the TMA unit writes ([F frames x Vr rows] x 64 channels, SWIZZLE_128B) are read by the tensor
// core as MN-major operands (as in tc4_wgrad.cuh).  A CTA owns ONE branch (blockIdx.y) and walks (sample, 4 output frames) tiles:
//
//   stage = [dO window | H at tap 0 | H at tap 1 | H at tap 2]: four atoms (each ONE 4-D TMA box of F frames x Vr rows, densely
//           packed: a reduction over rows needs no frame slots) whose channel window starts at the branch's 8-aligned
//           first channel lo8 (so the branch sits at columns [off, off + w) of its own atoms and no operand starts mid-atom);
//           a tap is just another frame coordinate of the 4-D tensor map (frames outside the sample are zero-filled by the TMA
//           unit = the convolution's zero padding; a temporal stride selects a parity-plane tensor map)
//   MMA     D_tap[64 x Kp] (+)= dO_window^T x H_tap  (M = 64 channels of the dO window, N = Kp = roundup16(off + w), K = 16 rows)
//           and the column sums of dO through a "ones" operand; three accumulators + sums live in TMEM for the whole kernel
//   final   epilogue once per CTA: rows / columns [off, off + w) of the accumulators -> red.global.add into dW, db
//
// The four branch-CTAs of a tile index run side by side, so the overlapping channel windows and the tap re-reads are L2 hits.
#pragma once
#include "tc4_common.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc4 {

constexpr int TW_THREADS = 32 * 6;            // warp 0 producer, warp 1 MMA issuer, warps 2-5 final epilogue
constexpr int TW_MAX_STAGES = 3;

struct TWPlan {
    int nb;
    int lo8[8], off[8], w[8], Kp[8], dil[8];
    int F, slot, Vr, stride, tps, n_tiles, ksteps, T_out;
    int S;
    unsigned stage_bytes, off_ones, smem_total;
    int tmem_cols;
};

struct TWBars {
    uint64_t full[TW_MAX_STAGES], empty[TW_MAX_STAGES], done;
};

__global__ void __launch_bounds__(TW_THREADS, 1)
tc4_twgrad_kernel(const __grid_constant__ CUtensorMap mapD, const __grid_constant__ CUtensorMap mapH0, const __grid_constant__ CUtensorMap mapH1,
                  const dsg_ms_conv_args a, const TWPlan p) {
    DSG_DYN_SMEM(smem_raw);
    __shared__ TWBars bars;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = sm;
    unsigned char* ones = sm + p.off_ones;
    const int b = blockIdx.y;
    const int Kp = p.Kp[b], d = p.dil[b], s = p.stride;
    const int colS = 3 * Kp;

    // ---- one-time setup: padding rows of the frame slots are never written by the TMA unit and must read as zero
    {
        const unsigned n16 = (unsigned)p.S * p.stage_bytes / 16u;
        for (unsigned i = tid; i < n16; i += TW_THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    for (int i = tid; i < 256; i += TW_THREADS) reinterpret_cast<uint16_t*>(ones)[i] = 0x3F80;      // bf16 1.0
    if (tid == 0) {
        for (int st = 0; st < TW_MAX_STAGES; ++st) { mbar_init(&bars.full[st], 1); mbar_init(&bars.empty[st], 1); }
        mbar_init(&bars.done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int n_my = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            prefetch_map(&mapD);
            prefetch_map(&mapH0);
            if (s == 2) prefetch_map(&mapH1);
            const uint32_t tx = (uint32_t)(4 * p.F * p.Vr * 128);
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                const int smp = tile / p.tps, q0 = (tile - smp * p.tps) * p.F;
                mbar_wait(&bars.empty[stage], ph ^ 1);
                mbar_expect_tx(&bars.full[stage], tx);
                unsigned char* st = ring + (size_t)stage * p.stage_bytes;
                tma_load_4d(st, &mapD, p.lo8[b], 0, q0, smp, &bars.full[stage]);
                for (int tap = 0; tap < 3; ++tap) {
                    const int o = (tap - 1) * d;
                    const int par = s == 2 ? (o & 1) : 0;                         // parity plane of s*t' + o
                    const int sh = s == 2 ? (o - par) / 2 : o;                    // frame offset inside the plane
                    const CUtensorMap* m = par ? &mapH1 : &mapH0;
                    unsigned char* dst = st + (size_t)(1 + tap) * ATOM_BYTES;
                    tma_load_4d(dst, m, p.lo8[b], 0, q0 + sh, smp, &bars.full[stage]);
                }
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================================== MMA issuer ==================================================
        if (lane == 0 && n_my > 0) {
            const uint32_t idesc_w = idesc_major(64, Kp, 1, 1);
            const uint32_t idesc_s = idesc_major(64, 8, 1, 0);
            const uint32_t ones_d = smem_u32(ones);
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                mbar_wait(&bars.full[stage], ph);
                tc_fence_after();
                const uint32_t d0 = smem_u32(ring + (size_t)stage * p.stage_bytes);
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    const uint32_t acc = (i | ks) ? 1u : 0u;
                    const uint64_t dd = desc_mn_sw128(d0 + ks * 2048u, ATOM_BYTES);
                    for (int tap = 0; tap < 3; ++tap)
                        umma_f16(tmem + (uint32_t)(tap * Kp), dd, desc_mn_sw128(d0 + (uint32_t)(1 + tap) * ATOM_BYTES + ks * 2048u, ATOM_BYTES), idesc_w, acc);
                    umma_f16(tmem + (uint32_t)colS, dd, desc_ones(ones_d), idesc_s, acc);
                }
                umma_commit(&bars.empty[stage]);
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
            umma_commit(&bars.done);
        }
    } else {
        // ================================================ final epilogue ================================================
        if (n_my > 0) {
            const int q = warp & 3;                       // TMEM lane quarter this warp may read
            const uint32_t lanes = (uint32_t)(q * 32) << 16;
            mbar_wait(&bars.done, 0);
            tc_fence_after();
            const int m = lane < 16 ? q * 16 + lane : -1;                         // accumulator row of an M = 64 MMA held by this lane
            const int co = m - p.off[b];
            const bool row_ok = m >= 0 && co >= 0 && co < p.w[b];
            const int w = p.w[b];
            float* dW = a.br[b].dW;
            for (int tap = 0; tap < 3; ++tap)
                for (int g = 0; g < (Kp >> 4); ++g) {
                    float v[16];
                    tmem_ld16(tmem + (uint32_t)(tap * Kp + g * 16) + lanes, v);
                    if (!row_ok || !dW) continue;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int ci = g * 16 + e - p.off[b];
                        if (ci >= 0 && ci < w) atomicAdd(dW + ((long long)co * w + ci) * 3 + tap, v[e]);
                    }
                }
            float s8[8];
            tmem_ld8(tmem + (uint32_t)colS + lanes, s8);
            if (row_ok && a.br[b].db) atomicAdd(a.br[b].db + co, s8[0]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

static const char* launch_ms_conv_wgrad_tc4(const dsg_ms_conv_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (!tc4_enabled() || !encode_fn()) return nullptr;
    if (a.n_branches < 1 || a.n_branches > 8 || a.stride < 1 || a.stride > 2 || a.Vr < 1 || a.Vr > 32) return nullptr;
    if (!tma_ptr_ok(a.src, a.ld_src) || !tma_ptr_ok(a.out, a.ld_out)) return nullptr;
    if (a.n_samples <= 0 || a.T_out <= 0) { *handled = true; return nullptr; }
    TWPlan p{};
    p.nb = a.n_branches;
    int hi_all = 0;
    for (int b = 0; b < p.nb; ++b) {
        const dsg_ms_branch& br = a.br[b];
        if (br.kind != 0 || br.hi <= br.lo || br.dilation < 1) return nullptr;
        p.lo8[b] = br.lo & ~7;
        p.off[b] = br.lo - p.lo8[b];
        p.w[b] = br.hi - br.lo;
        p.Kp[b] = (p.off[b] + p.w[b] + 15) & ~15;
        p.dil[b] = br.dilation;
        if (p.Kp[b] > ATOM_CH) return nullptr;            // the branch must fit one 64-channel atom from its aligned start
        if (br.hi > hi_all) hi_all = br.hi;
    }
    p.Vr = a.Vr;
    p.slot = a.Vr;                                        // dense rows: F frames of Vr rows per atom, the tail rows stay zero
    p.F = ATOM_ROWS / a.Vr;
    p.ksteps = (p.F * a.Vr + 15) / 16;
    p.stride = a.stride;
    p.T_out = a.T_out;
    p.tps = (a.T_out + p.F - 1) / p.F;
    const long long nt = (long long)a.n_samples * p.tps;
    if (nt > 0x3fffffff) return nullptr;
    p.n_tiles = (int)nt;
    p.stage_bytes = 4u * ATOM_BYTES;
    p.S = TW_MAX_STAGES;
    p.off_ones = (unsigned)p.S * p.stage_bytes;
    p.smem_total = p.off_ones + 1024u + 1024u;
    p.tmem_cols = 256;                                    // 3 * Kp (<= 192) + 8
    // tensor maps: channel extent = what the tensors really hold (windows past it are zero-filled)
    const int Csrc = (int)(a.ld_src < hi_all + 64 ? a.ld_src : hi_all + 64);
    const int Cout = (int)(a.ld_out < hi_all + 64 ? a.ld_out : hi_all + 64);
    CUtensorMap mD, mH0, mH1;
    bool ok = make_map_4d(&mD, a.out, a.n_samples, a.T_out, a.Vr, Cout, a.ld_out, 0, 1, p.F);
    ok = ok && make_map_4d(&mH0, a.src, a.n_samples, a.T_in, a.Vr, Csrc, a.ld_src, 0, a.stride, p.F);
    mH1 = mH0;
    if (ok && a.stride == 2 && a.T_in > 1) ok = make_map_4d(&mH1, a.src, a.n_samples, a.T_in, a.Vr, Csrc, a.ld_src, 1, 2, p.F);
    if (!ok) return nullptr;
    int gx = num_sms() / p.nb;
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    cudaFuncSetAttribute(tc4_twgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_total);
    tc4_twgrad_kernel<<<dim3((unsigned)gx, (unsigned)p.nb), dim3(TW_THREADS), p.smem_total, st>>>(mD, mH0, mH1, a, p);
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc4
}  // namespace dsg
#endif
