// tc4: TMA-fed, warp-specialised, persistent tcgen05 engine of dsg_conv_gemm for the 1x1 convolutions (taps == 1, no frame
// remap) — north-star kernels (a)/(c): "1x1 channel GEMM on tcgen05 tensor cores fed by TMA".
//
//   warp 0        TMA producer: one elected thread streams the A operand as SWIZZLE_128B atoms ([128 rows x 64 channels],
//                 cp.async.bulk.tensor, S-stage mbarrier ring, 64-96 KB in flight per SM) and brings the CTA's packed weight
//                 tile once (cp.async.bulk).  No thread ever touches an input row: out-of-bounds rows/channels are
//                 zero-filled by the TMA unit.
//   warp 1        MMA issuer: tcgen05.mma kind::f16 (M = 128, N <= 128) straight from the TMA-written atoms into one of two
//                 TMEM accumulators; tcgen05.commit frees the stage / publishes the accumulator.  It also issues the
//                 BatchNorm-statistics MMAs (below).
//   warps 2-5     transform warps (only when the operand is not a plain tensor): BN-affine + ReLU prologue in place, the
//                 joint-mean row of dgmstcn (tcn.py:409) or its gradient fold-back, then fence.proxy.async + arrive.
//   warps 6-13    epilogue: thread = accumulator row.  tcgen05.ld, bias / addends / per-sample broadcast / ReLU mask, bf16 pack,
//                 16-byte conflict-free stores into a SWIZZLE_128B out tile in shared memory.  They never wait on anything but
//                 "accumulator full" and "out tile free".
//   warp 14       drain: one thread issues the TMA store of the finished out tile (boundary clipping by the TMA unit) and the
//                 BatchNorm-statistics MMAs over it, waits for both to have read the tile and hands it back.
//
// Two algebraic moves keep CUDA cores off the data path:
//  * BatchNorm-backward prologue dy = ca*e + cb*y + cc (dsg_act_src with two tensors, no ReLU) is folded into the GEMM:
//    A = [e | y] (two tensor maps, K concatenated), W' = [diag(ca) W ; diag(cb) W], bias' = cc^T W — computed by the weight
//    pack kernel in fp32 and rounded once to bf16; e and y go from HBM to the tensor core untouched.
//  * BatchNorm statistics run on the tensor core over the bf16 out tile (A operand, MN-major: reduction over the tile rows):
//    sum v = tile^T x ones (N = 8 "ones" B operand); sum v^2 = diagonal of the Gram matrix tile^T x tile (forward), or
//    sum v*partner = column sums of a bf16 product tile the epilogue writes next to it (BatchNorm backward).  The sums
//    accumulate in TMEM over all tiles of the CTA; one fp64 atomic per channel per CTA at the end.  (The shuffle
//    transpose-reduce of the older engines cost ~150 instructions per 16 columns per thread.)
#pragma once
#include "tc4_common.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc4 {

constexpr int G4_XF_WARPS = 4, G4_EPI_WARPS = 8;
constexpr int G4_DRAIN_WARP = 2 + G4_XF_WARPS + G4_EPI_WARPS;            // warp 14: TMA stores + statistics MMAs
constexpr int G4_THREADS = 32 * (G4_DRAIN_WARP + 1);                     // 480
constexpr int G4_XF_T0 = 64, G4_EPI_T0 = 64 + 32 * G4_XF_WARPS;         // first thread of the transform / epilogue groups
constexpr int G4_EPI_THREADS = 32 * G4_EPI_WARPS;                       // 256
constexpr int G4_MAX_ATOMS = 12, G4_MAX_STAGES = 6, G4_MAX_OPS = 24, G4_MAX_OB = 4;
// experiment knobs (environment, read once): cap of the operand ring depth / of the out-tile ring depth, transform mapping
static inline int tc4_env_int(const char* name, int dflt) { const char* e = getenv(name); return (e && e[0]) ? atoi(e) : dflt; }
static inline int tc4_s_cap() { static const int v = tc4_env_int("DSG_TC4_S", G4_MAX_STAGES); return v < 2 ? 2 : (v > G4_MAX_STAGES ? G4_MAX_STAGES : v); }
static inline int tc4_ob_max() { static const int v = tc4_env_int("DSG_TC4_OB", 2); return v < 1 ? 1 : (v > G4_MAX_OB ? G4_MAX_OB : v); }
static inline int tc4_drain_defer() { static const int v = tc4_env_int("DSG_TC4_DEFER", 0); return v; }
static inline int tc4_dbg() {             // timing-only ablation (skips work: WRONG results): compiled in only with -DDSG_TC4_ABLATION
#ifdef DSG_TC4_ABLATION
    static const int v = tc4_env_int("DSG_TC4_DBG", 0);
    return v;
#else
    return 0;
#endif
}
static inline int tc4_tail_tma() { static const int v = tc4_env_int("DSG_TC4_TAILTMA", 1); return v; }
static inline int tc4_gram_ones() { static const int v = tc4_env_int("DSG_TC4_GRAMONES", 1); return v; }
static inline int tc4_xf_map() { static const int v = tc4_env_int("DSG_TC4_XFMAP", 1); return v; }
constexpr int G4_BAR_XF = 1, G4_BAR_EPI = 2;                            // named barriers

struct G4Plan {
    int mode;                    // 0: plain rows; 1: ext_in (joint-mean row appended per frame); 2: contract_ext (mean row folded back)
    int V, slot, F;              // joints per frame (without the mean row), rows per frame slot in the tile (8-aligned), frames per tile
    int n_tiles;
    long long rows_out, n_frames;
    int natoms, natoms1;         // A atoms per tile; atoms [0, natoms1) come from x1, the rest from x2
    int ksteps[G4_MAX_ATOMS];    // K = 16 MMA steps per atom (live channels only)
    int K1p;                     // natoms1 * 64
    int Ntile, S, OB;            // output columns per CTA, A stages, out/stat buffers
    // ---- staged tails: the epilogue's tail operands (addends, partner, mask sources) of a tile are brought in by the TMA unit with the
    //      geometry of the out store (same boxes, same SWIZZLE_128B atoms) into a TB-deep ring, [tensor][out atom] per slot; the epilogue
    //      reads them at the offsets it writes the out tile at.  tn = 0: per-thread 16-byte global loads (row-strided: 32 L1 tags per
    //      warp instruction — the tails were bound by the LSU tag stage, bench_gemm bwd64: 119 us, 57 us without the tail work).
    int tn, TB;                  // distinct tail tensors staged (<= 4), ring depth
    int t_add, t_add2, t_part, t_m1, t_m2;      // ring tensor index of each operand (-1: absent)
    unsigned off_tail, tail_stage_bytes;
    int gram_ones;               // 64-wide tiles: the Gram MMA's B operand is [tile | ones atom] (N = 80), so sum x sits next to the Gram block
    unsigned off_ones16;         //   and the 8 ones-MMAs per tile are not issued
    int dbg;                     // timing experiments only (DSG_TC4_DBG; results are wrong): 1 skip statistics MMAs, 2 skip out stores, 4 skip epilogue math
    int drain_defer;             // drain warp retires tile i-1 after issuing tile i (pipelined store / statistics completion)
    int xfmap;                   // transform-warp mapping of the affine prologue: 1 = thread owns a 16-byte chunk column (coefficients in registers)
    int xf, act, stats;          // stats: 0 none, 1 Gram (sum v, sum v^2), 2 product tile (sum v, sum v*partner)
    unsigned off_w, off_a, off_out, off_stat, off_ones, off_cf;       // byte offsets from the 1024-aligned base
    unsigned w_tile_bytes, out_bytes, smem_total;
    int acc_cols, stat_col, sum_col, tmem_cols;   // stat_col: Gram / product sums; sum_col: plain sums (8 columns)
    // ---- mode 3: the dilated temporal convolutions of a multi-scale unit (tcn.py:383-396) as tap-shifted atoms.  A tile is F
    //      frames of ONE sample; atom = (branch, tap): the branch's 64-channel window of the source at frame + shift (4-D tensor
    //      map: frames outside the sample are zero-filled = the convolution's zero padding), multiplied into the branch's
    //      column window of the accumulator.  Per column tile y: t_n[y] atoms.
    int tps, qs, qp, Tq, Tdst;   // tiles per sample; frame stride / parity of the destination plane; its frames; T of the destination
    // ---- mode 4: the adjacency contraction fused in front of the `post` 1x1 convolution (gcn.py:2350-2363).  Tiles are F frames of
    //      one sample (as mode 3), walked in CONSECUTIVE order by a CTA so the sample's adjacency slice adyn[n] stays in shared
    //      memory; a stage is two atoms: the raw `pre` tile the TMA unit writes and the contracted tile the tensor core reads.
    int tpc;                     // tiles per CTA (consecutive)
    unsigned off_adj, adj_bytes; // shared-memory slice [V*V][K] bf16
    int store_y;                 // also write the contracted tile to HBM (training: operand of the weight gradient)
    int cvar;                    // (V, K) variant of the contraction code: {25,17,18} x {24,48}
    //      An "atom" of mode 3 is a WINDOW of t_nfr consecutive source frames of one 64-channel column (t_nfr = F: a single tap;
    //      t_nfr = F + halo: every tap of every branch that lives in the column reads it at a row offset — the frames are loaded
    //      once instead of once per (branch, tap)); its MMA ops are o_*[t_op0 .. t_op0 of the next atom).
    int t_n[2];
    short t_map[2][G4_MAX_ATOMS], t_tsh[2][G4_MAX_ATOMS], t_nfr[2][G4_MAX_ATOMS], t_op0[2][G4_MAX_ATOMS + 1];
    int t_c0[2][G4_MAX_ATOMS];
    short o_row[2][G4_MAX_OPS], o_k0[2][G4_MAX_OPS], o_ks[2][G4_MAX_OPS], o_ncol[2][G4_MAX_OPS], o_nw[2][G4_MAX_OPS];
    unsigned o_woff[2][G4_MAX_OPS], t_wbytes[2], t_stage;
};

// ---- packed weights: per column tile j, per atom a: [Ntile rows (n) x 64 k] bf16 in the K-major SWIZZLE_128B layout, scaled
//      per k when a BatchNorm-backward / affine prologue is folded in; cbias[n] = bias[n] + sum_k (c1[k] + c2[k]) W[n,k]
__global__ void __launch_bounds__(256) tc4_wpack_kernel(const float* W, long long ws_n, long long ws_k, int K, int N, int natoms1, int natoms,
                                                        int Ntile, const float* s1, const float* s2, const float* c1, const float* c2,
                                                        const float* bias, unsigned char* out, float* cbias, float fold_scale = 1.f) {
    const int j = blockIdx.x, a = blockIdx.y;
    if (a == natoms) {                                  // folded bias for this column tile
        for (int nl = threadIdx.x; nl < Ntile; nl += 256) {
            const int n = j * Ntile + nl;
            if (n >= N) continue;
            float acc = 0.f;
            if (c1 || c2)
                for (int k = 0; k < K; ++k) {
                    const float c = (c1 ? c1[k] : 0.f) + (c2 ? c2[k] : 0.f);
                    acc = fmaf(c, W[(long long)n * ws_n + (long long)k * ws_k], acc);
                }
            // contract_ext: every joint row receives its own folded constant AND 1/V of the mean row's (fold_scale = 1 + 1/V)
            cbias[n] = fmaf(acc, fold_scale, bias ? bias[n] : 0.f);
        }
        return;
    }
    const bool second = a >= natoms1;
    const int k0 = (second ? a - natoms1 : a) * ATOM_CH;
    const float* sc = second ? s2 : s1;
    unsigned char* dst = out + ((size_t)j * natoms + a) * (size_t)Ntile * 128;
    for (int idx = threadIdx.x; idx < Ntile * 8; idx += 256) {
        int nl, ch;
        if (ws_k == 1) { ch = idx & 7; nl = idx >> 3; } else { nl = idx % Ntile; ch = idx / Ntile; }
        const int n = j * Ntile + nl;
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = k0 + ch * 8 + e;
            float w = (n < N && k < K) ? W[(long long)n * ws_n + (long long)k * ws_k] : 0.f;
            if (sc && k < K) w *= sc[k];
            v[e] = w;
        }
        *reinterpret_cast<uint4*>(dst + atom_off(nl, ch)) = pack8(v);
    }
}

// one (frame, channel) column of the fused adjacency contraction: VV source joints in registers; VV and KK (channels of the
// adjacency slice) are compile-time, so every shared-memory operand address is base + immediate — the inner loop is load, bf16
// unpack, FMA and nothing else (the first version spent three IMADs per FMA on addresses: cuobjdump, 1549 IMAD vs 495 FFMA)
template <int VV, int KK, int J>
DSG_D void xf_contract_group(const float* pv, const bf16* aw, unsigned char* ycol, int row, int chunk) {
    float acc[J];
#pragma unroll
    for (int j = 0; j < J; ++j) acc[j] = 0.f;
#pragma unroll
    for (int u = 0; u < VV; ++u) {
#pragma unroll
        for (int j = 0; j < J; ++j) acc[j] = fmaf(pv[u], __bfloat162float(aw[(u * VV + j) * KK]), acc[j]);
    }
#pragma unroll
    for (int j = 0; j < J; ++j) *reinterpret_cast<bf16*>(ycol + atom_off(row + j, chunk)) = __float2bfloat16(acc[j]);
}
template <int VV, int KK>
DSG_D void xf_contract_col(const unsigned char* rcol, unsigned char* ycol, const bf16* ac, int row0, int chunk, float ka, float kb, bool relu) {
    float pv[VV];
#pragma unroll
    for (int u = 0; u < VV; ++u) {
        const float v = fmaf(__bfloat162float(*reinterpret_cast<const bf16*>(rcol + atom_off(row0 + u, chunk))), ka, kb);
        pv[u] = relu ? fmaxf(v, 0.f) : v;
    }
#pragma unroll 1
    for (int w0 = 0; w0 + 4 <= VV; w0 += 4) xf_contract_group<VV, KK, 4>(pv, ac + w0 * KK, ycol, row0 + w0, chunk);
    constexpr int TAIL = VV % 4;
    if (TAIL) xf_contract_group<VV, KK, TAIL ? TAIL : 1>(pv, ac + (VV - TAIL) * KK, ycol, row0 + VV - TAIL, chunk);
}

constexpr int G4_MAX_TAILS = 4, G4_TB = 2;
struct G4TailMaps { CUtensorMap m[G4_MAX_TAILS]; };

struct G4Bars {
    uint64_t full[G4_MAX_STAGES], empty[G4_MAX_STAGES], ready[G4_MAX_STAGES];
    uint64_t wbar, acc_full[2], acc_free[2], out_ready[G4_MAX_OB], stat_done[G4_MAX_OB], out_free[G4_MAX_OB];
    uint64_t tfull[G4_TB], tempty[G4_TB];
};

// XF: transform warps active; TAILS: any of add / add2 / bcast / mask / partner; STATS: 0 none, 1 Gram, 2 product tile
template <bool XF, bool TAILS, int STATS>
__global__ void __launch_bounds__(G4_THREADS, 1)
tc4_gemm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1, const __grid_constant__ CUtensorMap mapO,
                const dsg_conv_gemm_args a, const G4Plan p, const float* __restrict__ cbias, const __grid_constant__ G4TailMaps tm) {
    // (mode 4 passes the tensor map of the contracted-tile side output Y as mapA1)
    DSG_DYN_SMEM(smem_raw);
    __shared__ G4Bars bars;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* Wsm = sm + p.off_w;
    unsigned char* Asm = sm + p.off_a;
    unsigned char* Osm = sm + p.off_out;
    unsigned char* Ssm = sm + p.off_stat;
    unsigned char* Tsm = sm + p.off_tail;
    unsigned char* ones = sm + p.off_ones;
    float* tl_bias = reinterpret_cast<float*>(sm + p.off_cf);
    float* tl_ma1 = tl_bias + 128;
    float* tl_mb = tl_ma1 + 128;
    float* tl_ma2 = tl_mb + 128;
    float* cf_a = tl_ma2 + 128;                         // [K1p] prologue coefficients of x1 (act mode)
    float* cf_b = cf_a + p.K1p;
    const int n0 = blockIdx.y * p.Ntile;
    const int Nt = a.N - n0 < p.Ntile ? a.N - n0 : p.Ntile;
    const int Ntp = (Nt + 15) & ~15;
    const int n_oatoms = (Ntp + ATOM_CH - 1) / ATOM_CH;

    // ---- one-time setup
    if (tid < 128) {
        const int cch = n0 + tid;
        const bool in = cch < a.N;
        tl_bias[tid] = in ? cbias[cch] : 0.f;
        tl_ma1[tid] = (in && a.has_mask && a.mask.a1) ? a.mask.a1[cch] : 1.f;
        tl_mb[tid] = ((in && a.has_mask && a.mask.b1) ? a.mask.b1[cch] : 0.f) + ((in && a.has_mask && a.mask.b2) ? a.mask.b2[cch] : 0.f);
        tl_ma2[tid] = (in && a.has_mask && a.mask.a2) ? a.mask.a2[cch] : 1.f;
    }
    if (XF && p.act)
        for (int k = tid; k < p.K1p; k += G4_THREADS) {
            const bool in = k < a.K;
            cf_a[k] = (in && a.src.a1) ? a.src.a1[k] : 1.f;
            cf_b[k] = ((in && a.src.b1) ? a.src.b1[k] : 0.f) + ((in && a.src.b2) ? a.src.b2[k] : 0.f);
        }
    for (int i = tid; i < 256; i += G4_THREADS) reinterpret_cast<uint16_t*>(ones)[i] = 0x3F80;      // bf16 1.0
    if (p.gram_ones)
        for (int i = tid; i < ATOM_BYTES / 16; i += G4_THREADS) reinterpret_cast<uint4*>(sm + p.off_ones16)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    if (p.mode == 4)             // channels >= K and padding rows of the contracted atoms are never written: they must read as zero
        for (unsigned i = tid; i < (unsigned)p.S * 2u * ATOM_BYTES / 16u; i += G4_THREADS) reinterpret_cast<uint4*>(Asm)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int s = 0; s < G4_MAX_STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); mbar_init(&bars.ready[s], 32 * G4_XF_WARPS); }
        mbar_init(&bars.wbar, 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&bars.acc_full[b], 1); mbar_init(&bars.acc_free[b], G4_EPI_WARPS); }
        for (int b = 0; b < G4_TB; ++b) { mbar_init(&bars.tfull[b], 1); mbar_init(&bars.tempty[b], G4_EPI_WARPS); }
        for (int b = 0; b < G4_MAX_OB; ++b) {
            mbar_init(&bars.out_ready[b], G4_EPI_WARPS); mbar_init(&bars.stat_done[b], 1); mbar_init(&bars.out_free[b], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) tmem_alloc(&tmem_base_s, (uint32_t)p.tmem_cols);
    fence_async_smem();                                  // the ones tile is read by the tensor core (async proxy)
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    // tiles of this CTA: strided (tile = blockIdx.x + i * gridDim.x), or a consecutive run in mode 4
    int n_my = ((int)blockIdx.x < p.n_tiles) ? (p.n_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
    const int tile0 = p.mode == 4 ? (int)blockIdx.x * p.tpc : (int)blockIdx.x, tstep = p.mode == 4 ? 1 : (int)gridDim.x;
    if (p.mode == 4) { n_my = p.n_tiles - tile0; n_my = n_my < 0 ? 0 : (n_my > p.tpc ? p.tpc : n_my); }
    const uint32_t box_bytes = p.mode == 0 ? (uint32_t)ATOM_BYTES : (uint32_t)(p.F * (p.mode == 1 ? p.V : (p.mode == 2 ? p.V + 1 : a.Vin)) * 128);
    const int yy = p.mode == 3 ? (int)blockIdx.y : 0;
    const int natoms = p.mode == 3 ? p.t_n[yy] : p.natoms;
    const bool m4 = p.mode == 4;
    const uint32_t stage_bytes = m4 ? 2u * ATOM_BYTES : (p.mode == 3 ? p.t_stage : (uint32_t)ATOM_BYTES);      // mode 4: [raw | contracted]

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            prefetch_map(&mapA0);
            if (p.natoms > p.natoms1) prefetch_map(&mapA1);
            const uint32_t wbytes = p.mode == 3 ? p.t_wbytes[yy] : p.w_tile_bytes;
            const size_t wsrc = p.mode == 3 ? (yy ? (size_t)p.t_wbytes[0] : 0) : (size_t)blockIdx.y * p.w_tile_bytes;
            mbar_expect_tx(&bars.wbar, wbytes);
            bulk_g2s(Wsm, reinterpret_cast<const unsigned char*>(a.wpack) + wsrc, wbytes, &bars.wbar);
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = tile0 + i * tstep;
                const int smp = p.mode >= 3 ? tile / p.tps : 0, q0 = p.mode >= 3 ? (tile - smp * p.tps) * p.F : 0;
                if (TAILS && p.tn > 0) {
                    // tail operands of this tile: the boxes of the out store, one ring slot per tile
                    const int tb = i % p.TB;
                    mbar_wait(&bars.tempty[tb], (uint32_t)(((i / p.TB) & 1) ^ 1));
                    const int Vo = a.Vin + a.ext_in - a.contract_ext;
                    const uint32_t obox = p.mode == 0 ? (uint32_t)ATOM_BYTES : (uint32_t)(p.F * Vo * 128);
                    mbar_expect_tx(&bars.tfull[tb], (uint32_t)(p.tn * n_oatoms) * obox);
                    for (int t = 0; t < p.tn; ++t)
                        for (int oa = 0; oa < n_oatoms; ++oa) {
                            unsigned char* dst = Tsm + (size_t)tb * p.tail_stage_bytes + (size_t)t * p.out_bytes + (size_t)oa * ATOM_BYTES;
                            const int c0 = n0 + oa * ATOM_CH;
                            if (p.mode == 0) tma_load_2d(dst, &tm.m[t], c0, tile * ATOM_ROWS, &bars.tfull[tb]);
                            else if (p.mode >= 3)
                                for (int f = 0; f < p.F; ++f) tma_load_4d(dst + (size_t)f * p.slot * 128, &tm.m[t], c0, 0, q0 + f, smp, &bars.tfull[tb]);
                            else
                                for (int f = 0; f < p.F; ++f) tma_load_3d(dst + (size_t)f * p.slot * 128, &tm.m[t], c0, 0, tile * p.F + f, &bars.tfull[tb]);
                        }
                }
                for (int ai = 0; ai < natoms; ++ai) {
                    mbar_wait_backoff(&bars.empty[stage], ph ^ 1, m4);
                    mbar_expect_tx(&bars.full[stage], p.mode == 3 ? (uint32_t)(p.t_nfr[yy][ai] * a.Vin * 128) : box_bytes);
                    unsigned char* dst = Asm + (size_t)stage * stage_bytes;
                    if (m4) {
                        for (int f = 0; f < p.F; ++f) tma_load_4d(dst + (size_t)f * p.slot * 128, &mapA0, 0, 0, q0 + f, smp, &bars.full[stage]);
                    } else if (p.mode == 3) {
                        const CUtensorMap* m = p.t_map[yy][ai] ? &mapA1 : &mapA0;
                        for (int f = 0; f < p.t_nfr[yy][ai]; ++f)
                            tma_load_4d(dst + (size_t)f * p.slot * 128, m, p.t_c0[yy][ai], 0, q0 + f + p.t_tsh[yy][ai], smp, &bars.full[stage]);
                    } else {
                        const CUtensorMap* m = ai < p.natoms1 ? &mapA0 : &mapA1;
                        const int c0 = (ai < p.natoms1 ? ai : ai - p.natoms1) * ATOM_CH;
                        if (p.mode == 0) tma_load_2d(dst, m, c0, tile * ATOM_ROWS, &bars.full[stage]);
                        else
                            for (int f = 0; f < p.F; ++f) tma_load_3d(dst + (size_t)f * p.slot * 128, m, c0, 0, tile * p.F + f, &bars.full[stage]);
                    }
                    if (++stage == p.S) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================================================== MMA issuer ==================================================
        if (lane == 0) {
            mbar_wait(&bars.wbar, 0);
            const uint32_t idesc = make_idesc(128, Ntp);
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int buf = i & 1, use = i >> 1;
                if (use > 0) mbar_wait_backoff(&bars.acc_free[buf], (uint32_t)((use - 1) & 1), m4);
                tc_fence_after();
                const uint32_t acc = tmem + (uint32_t)(buf * p.acc_cols);
                int first = 1;
                for (int ai = 0; ai < natoms; ++ai) {
                    mbar_wait_backoff(XF ? &bars.ready[stage] : &bars.full[stage], ph, m4);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(Asm + (size_t)stage * stage_bytes) + (m4 ? (uint32_t)ATOM_BYTES : 0u);
                    if (m4 && p.store_y) {
                        // side output of the fused contraction: the tile the tensor core is about to read, stored by the TMA unit
                        const int tile = tile0 + i * tstep, smp = tile / p.tps, q0 = (tile - smp * p.tps) * p.F;
                        for (int f = 0; f < p.F; ++f)
                            if (q0 + f < p.Tq) tma_store_4d(&mapA1, Asm + (size_t)stage * stage_bytes + ATOM_BYTES + (size_t)f * p.slot * 128, 0, 0, q0 + f, smp);
                        tma_store_commit();
                    }
                    if (p.mode == 3) {
                        // every (branch, tap) op of this window: rows shifted by the tap (whole frame slots: the SWIZZLE_128B phase is
                        // kept), the branch's channels as the K range, its column window of the accumulator as D (op 0 spans the tile)
                        for (int op = p.t_op0[yy][ai]; op < p.t_op0[yy][ai + 1]; ++op) {
                            const uint32_t ar = a0 + (uint32_t)p.o_row[yy][op] * (uint32_t)p.slot * 128u + (uint32_t)p.o_k0[yy][op] * 32u;
                            const uint32_t w0 = smem_u32(Wsm + p.o_woff[yy][op]), idw = make_idesc(128, p.o_nw[yy][op]);
                            const uint32_t dcol = acc + (uint32_t)p.o_ncol[yy][op];
                            for (int ks = 0; ks < p.o_ks[yy][op]; ++ks)
                                umma_f16(dcol, desc_k_sw128(ar + ks * 32u), desc_k_sw128(w0 + ks * 32u), idw, (op | ks) ? 1u : 0u);
                        }
                    } else {
                        const uint32_t w0 = smem_u32(Wsm + (size_t)ai * p.Ntile * 128);
                        for (int ks = 0; ks < p.ksteps[ai]; ++ks) {
                            umma_f16(acc, desc_k_sw128(a0 + ks * 32u), desc_k_sw128(w0 + ks * 32u), idesc, first ? 0u : 1u);
                            first = 0;
                        }
                    }
                    if (m4 && p.store_y) tma_store_wait_read<0>();      // the stage is reused once `empty` fires: the store must have read it
                    umma_commit(&bars.empty[stage]);
                    if (++stage == p.S) { stage = 0; ph ^= 1; }
                }
                umma_commit(&bars.acc_full[buf]);
            }
            if (m4 && p.store_y) tma_store_wait_all<0>();
        }
    } else if (warp < 2 + G4_XF_WARPS) {
        // ================================================ transform warps ===============================================
        if (XF && m4) {
            // ---- fused adjacency contraction (north-star kernel (a), dynamic case): y[t,w,c] = sum_u relu(bn(p))[t,u,c] * adyn[n,u,w,c].
            //      thread = (frame, channel) column of the tile: its V source joints live in registers, the sample's adjacency slice in
            //      shared memory (bf16 [u][w][K]: lanes = consecutive channels, conflict-free), the result goes straight into the
            //      SWIZZLE_128B atom the tensor core multiplies with the `post` weights — Y is never read back from HBM.
            const int t = tid - G4_XF_T0;                 // 0..127
            const int K = a.K, V = p.V;
            const bf16* adj_s = reinterpret_cast<const bf16*>(sm + p.off_adj);
            int stage = 0, cur_smp = -1;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = tile0 + i * tstep, smp = tile / p.tps;
                if (smp != cur_smp) {
                    // new sample: bring its adjacency slice (contiguous [V*V*K] bf16) — the previous tile's readers are past it
                    named_sync(G4_BAR_XF, 32 * G4_XF_WARPS);
                    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.adyn) + (size_t)smp * V * V * K);
                    uint4* dst = reinterpret_cast<uint4*>(sm + p.off_adj);
                    for (int j = t; j < V * V * K / 8; j += 32 * G4_XF_WARPS) dst[j] = src[j];
                    named_sync(G4_BAR_XF, 32 * G4_XF_WARPS);
                    cur_smp = smp;
                }
                mbar_wait(&bars.full[stage], ph);
                const unsigned char* raw = Asm + (size_t)stage * stage_bytes;
                unsigned char* yat = Asm + (size_t)stage * stage_bytes + ATOM_BYTES;
                for (int col = t; col < p.F * K; col += 32 * G4_XF_WARPS) {
                    const int f = col / K, c = col - f * K;
                    const float ka = cf_a[c], kb = cf_b[c];
                    const uint32_t coff = (uint32_t)(c & 7) * 2u;
                    const unsigned char* rcol = raw + coff;
                    unsigned char* ycol = yat + coff;
                    const bf16* ac = adj_s + c;
                    const int row0 = f * p.slot, chunk = c >> 3;
                    const bool relu = a.src.relu != 0;
                    switch (p.cvar) {
                        case 0: xf_contract_col<25, 24>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                        case 1: xf_contract_col<25, 48>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                        case 2: xf_contract_col<17, 24>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                        case 3: xf_contract_col<17, 48>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                        case 4: xf_contract_col<18, 24>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                        default: xf_contract_col<18, 48>(rcol, ycol, ac, row0, chunk, ka, kb, relu); break;
                    }
                }
                fence_async_smem();
                mbar_arrive(&bars.ready[stage]);
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
        } else if (XF) {
            const int t = tid - G4_XF_T0;                 // 0..127
            const int Vr = p.mode == 2 ? p.V + 1 : p.V;   // rows the TMA wrote per frame slot
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                for (int ai = 0; ai < p.natoms; ++ai) {
                    mbar_wait(&bars.full[stage], ph);
                    unsigned char* atom = Asm + (size_t)stage * ATOM_BYTES;
                    const int kbase = (ai < p.natoms1 ? ai : ai - p.natoms1) * ATOM_CH;
                    if (p.act && ai < p.natoms1) {
                        // BN-affine (+ReLU) in place: thread = row; the logical chunk is uniform per step (broadcast coefficient loads,
                        // conflict-free 16-byte data accesses)
                        const bool live = p.mode == 0 ? true : ((t % p.slot) < Vr && t / p.slot < p.F);
                        if (p.xfmap) {
                            // thread = (16-byte chunk column c, row residue r0): the column's 16 coefficients are loaded once per atom and the
                            // thread's 8 rows (r0 + 16 j: same swizzle phase, addresses = base + j * 2048) are independent — 8 loads in flight,
                            // no per-chunk coefficient reloads (ncu source page of the thread-per-row form: ~490 instructions per warp and atom
                            // at ~7 cycles each; the transform warps were busy 80 % of the kernel)
                            const int c = t & 7, r0 = t >> 3;
                            const int k = kbase + c * 8;
                            if (k < a.K) {
                                float ka[8], kb[8];
                                load8f(cf_a + k, ka, 1.f);
                                load8f(cf_b + k, kb, 0.f);
                                const float lo = a.src.relu ? 0.f : -3.0e38f;
                                unsigned char* col = atom + atom_off(r0, c);
                                uint4 qv[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) qv[j] = *reinterpret_cast<const uint4*>(col + j * 2048);
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const int row = r0 + 16 * j;
                                    float x[8];
                                    unpack8(qv[j], x);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) x[e] = fmaxf(fmaf(x[e], ka[e], kb[e]), lo);
                                    if (p.mode == 0 || ((row % p.slot) < Vr && row / p.slot < p.F)) *reinterpret_cast<uint4*>(col + j * 2048) = pack8(x);
                                }
                            }
                        } else if (live) {
#pragma unroll 2
                            for (int c = 0; c < 8; ++c) {
                                const int k = kbase + c * 8;
                                if (k >= a.K) break;
                                uint4* q = reinterpret_cast<uint4*>(atom + atom_off(t, c));
                                float x[8], ka[8], kb[8];
                                unpack8(*q, x);
                                load8f(cf_a + k, ka, 1.f);
                                load8f(cf_b + k, kb, 0.f);
#pragma unroll
                                for (int e = 0; e < 8; ++e) x[e] = fmaf(x[e], ka[e], kb[e]);
                                if (a.src.relu) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
                                }
                                *q = pack8(x);
                            }
                        }
                        if (p.mode != 0) named_sync(G4_BAR_XF, 32 * G4_XF_WARPS);
                    }
                    if (p.mode == 1) {
                        // joint-mean row of every frame: item = (frame, chunk), 4 lanes per item split the joints
                        const int items = p.F * 8 * 4;
                        for (int base = 0; base < items; base += 32 * G4_XF_WARPS) {
                            const int idx = base + t;
                            const bool ok = idx < items;
                            const int part = idx & 3, c = (idx >> 2) & 7, f = idx >> 5;
                            float s8[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) s8[e] = 0.f;
                            if (ok && kbase + c * 8 < a.K)
                                for (int v = part; v < p.V; v += 4) {
                                    float x[8];
                                    unpack8(*reinterpret_cast<const uint4*>(atom + atom_off(f * p.slot + v, c)), x);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) s8[e] += x[e];
                                }
#pragma unroll
                            for (int e = 0; e < 8; ++e) {
                                s8[e] += __shfl_xor_sync(0xffffffffu, s8[e], 1);
                                s8[e] += __shfl_xor_sync(0xffffffffu, s8[e], 2);
                            }
                            if (ok && part == 0) {
                                const float inv = 1.f / (float)p.V;
#pragma unroll
                                for (int e = 0; e < 8; ++e) s8[e] *= inv;
                                *reinterpret_cast<uint4*>(atom + atom_off(f * p.slot + p.V, c)) = pack8(s8);
                            }
                        }
                    } else if (p.mode == 2) {
                        // gradient of the joint mean: every joint row += mean row / V (linear, so it is applied to e and y alike)
                        const int items = p.F * 8 * 4;
                        const float inv = 1.f / (float)p.V;
                        for (int idx = t; idx < items; idx += 32 * G4_XF_WARPS) {
                            const int part = idx & 3, c = (idx >> 2) & 7, f = idx >> 5;
                            if (kbase + c * 8 >= a.K) continue;
                            float g[8];
                            unpack8(*reinterpret_cast<const uint4*>(atom + atom_off(f * p.slot + p.V, c)), g);
                            for (int v = part; v < p.V; v += 4) {
                                uint4* q = reinterpret_cast<uint4*>(atom + atom_off(f * p.slot + v, c));
                                float x[8];
                                unpack8(*q, x);
#pragma unroll
                                for (int e = 0; e < 8; ++e) x[e] = fmaf(g[e], inv, x[e]);
                                *q = pack8(x);
                            }
                        }
                    }
                    fence_async_smem();
                    mbar_arrive(&bars.ready[stage]);
                    if (++stage == p.S) { stage = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp < G4_DRAIN_WARP) {
        // =================================================== epilogue ===================================================
        const int q = warp & 3, half = (warp - 2 - G4_XF_WARPS) >> 2;      // TMEM lane quarter of this warp, column half
        const int r = q * 32 + lane;                      // accumulator row = tile row
        const int nc16 = Ntp >> 4;
        const int cbeg = half ? (nc16 + 1) >> 1 : 0, cend = half ? nc16 : (nc16 + 1) >> 1;
        const int Vout = a.Vin + a.ext_in - a.contract_ext;
        const bf16* addp = reinterpret_cast<const bf16*>(a.add);
        const bf16* add2p = reinterpret_cast<const bf16*>(a.add2);
        const bf16* partp = reinterpret_cast<const bf16*>(a.partner);
        for (int i = 0; i < n_my; ++i) {
            const int tile = tile0 + i * tstep;
            const int buf = i & 1, use = i >> 1, ob = i % p.OB, useo = i / p.OB;
            // ---- this thread's output row
            long long gr = -1;
            int jrow = 0, samp = 0;
            if (p.mode == 0) {
                const long long g = (long long)tile * ATOM_ROWS + r;
                if (g < p.rows_out) gr = g;
                if (TAILS && a.bcast && gr >= 0) { const long long fr = gr / Vout; jrow = (int)(gr - fr * Vout); samp = (int)(fr / a.T_out); }
            } else if (p.mode >= 3) {
                const int smp = tile / p.tps, q = (tile - smp * p.tps) * p.F + r / p.slot, v = r % p.slot;
                if (r / p.slot < p.F && v < Vout && q < p.Tq) { gr = ((long long)smp * p.Tdst + (long long)q * p.qs + p.qp) * Vout + v; jrow = v; samp = smp; }
            } else {
                const int f = r / p.slot, v = r - f * p.slot;
                const long long frame = (long long)tile * p.F + f;
                if (f < p.F && v < Vout && frame < p.n_frames) { gr = frame * Vout + v; jrow = v; samp = (int)(frame / a.T_out); }
            }
            mbar_wait_backoff(&bars.acc_full[buf], (uint32_t)(use & 1), m4);
            tc_fence_after();
            if (i >= p.OB) mbar_wait_backoff(&bars.out_free[ob], (uint32_t)((useo - 1) & 1), m4);     // out / product tiles of `OB` tiles ago have been read
            unsigned char* Ot = Osm + (size_t)ob * p.out_bytes;
            unsigned char* St = Ssm + (size_t)ob * p.out_bytes;
            const bool tst = TAILS && p.tn > 0;
            const int tb = tst ? i % p.TB : 0;
            const unsigned char* Tt = Tsm + (size_t)tb * p.tail_stage_bytes;
            if (tst) mbar_wait(&bars.tfull[tb], (uint32_t)((i / p.TB) & 1));
            const uint32_t acc = tmem + (uint32_t)(buf * p.acc_cols) + ((uint32_t)(q * 32) << 16);
            const bool row_ok = gr >= 0;
            const bool fold2 = !XF && p.mode == 2;       // gradient of the joint mean folded back on the accumulator (lane V of this warp)
            const float inv2 = 1.f / (float)p.V;
            for (int cc = cbeg; cc < ((p.dbg & 4) ? cbeg : cend); ++cc) {
                const int c16 = cc * 16;
                float v[16];
                if (!TAILS) {
                    // ---- lean path: bias, bf16 pack, (square for the statistics), conflict-free 16-byte stores
                    tmem_ld16(acc + (uint32_t)c16, v);
                    if (fold2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = fmaf(__shfl_sync(0xffffffffu, v[e], p.V), inv2, v[e]);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int col = c16 + h * 8;
                        float* vv = v + h * 8;
                        tc::add8(vv, tl_bias + col);
                        if (a.out_f32) {
                            // unrounded fp32 rows straight from the registers (32 bytes per thread and chunk: whole sectors)
                            if (row_ok && n0 + col < a.N) {
                                float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(a.out) + gr * a.ld_out + n0 + col);
                                dst[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                                dst[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
                            }
                            continue;
                        }
                        if (!row_ok) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) vv[e] = 0.f;
                        }
                        const uint32_t off = (uint32_t)(col >> 6) * ATOM_BYTES + atom_off(r, (col & 63) >> 3);
                        *reinterpret_cast<uint4*>(Ot + off) = pack8(vv);
                        if (STATS == 2) {
#pragma unroll
                            for (int e = 0; e < 8; ++e) vv[e] *= vv[e];
                            *reinterpret_cast<uint4*>(St + off) = pack8(vv);
                        }
                    }
                } else {
                    const int c = n0 + c16;
                    const bool live0 = c < a.N, live1 = c + 8 < a.N;
                    uint4 ra[2], ra2[2], rp[2];
                    tc::Act8Raw rm[2];
                    if (row_ok && tst) {
                        // staged tails: this thread's row of the ring slot, at the offsets of the out tile (conflict-free 16-byte reads)
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (!(h ? live1 : live0)) continue;
                            const int col = c16 + h * 8;
                            const unsigned char* tp = Tt + (uint32_t)(col >> 6) * ATOM_BYTES + atom_off(r, (col & 63) >> 3);
                            if (addp) ra[h] = *reinterpret_cast<const uint4*>(tp + (size_t)p.t_add * p.out_bytes);
                            if (add2p) ra2[h] = *reinterpret_cast<const uint4*>(tp + (size_t)p.t_add2 * p.out_bytes);
                            if (partp) rp[h] = *reinterpret_cast<const uint4*>(tp + (size_t)p.t_part * p.out_bytes);
                            if (a.has_mask) {
                                rm[h].a = *reinterpret_cast<const uint4*>(tp + (size_t)p.t_m1 * p.out_bytes);
                                rm[h].b = p.t_m2 >= 0 ? *reinterpret_cast<const uint4*>(tp + (size_t)p.t_m2 * p.out_bytes) : make_uint4(0u, 0u, 0u, 0u);
                            }
                        }
                    } else if (row_ok) {
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (!(h ? live1 : live0)) continue;
                            if (addp) ra[h] = *reinterpret_cast<const uint4*>(addp + gr * a.ld_add + c + h * 8);
                            if (add2p) ra2[h] = *reinterpret_cast<const uint4*>(add2p + gr * a.ld_add2 + c + h * 8);
                            if (partp) rp[h] = *reinterpret_cast<const uint4*>(partp + gr * a.ld_partner + c + h * 8);
                            if (a.has_mask) rm[h] = tc::act8_issue(a.mask, gr, c + h * 8);
                        }
                    }
                    tmem_ld16(acc + (uint32_t)c16, v);
                    if (fold2) {
#pragma unroll
                        for (int e = 0; e < 16; ++e) v[e] = fmaf(__shfl_sync(0xffffffffu, v[e], p.V), inv2, v[e]);
                    }
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int col = c16 + h * 8;                     // column inside the CTA tile
                        float* vv = v + h * 8;
                        uint4 o = make_uint4(0u, 0u, 0u, 0u), pr = make_uint4(0u, 0u, 0u, 0u);
                        if (row_ok && (h ? live1 : live0)) {
                            tc::add8(vv, tl_bias + col);
                            if (addp) { float t8[8]; unpack8(ra[h], t8);
#pragma unroll
                                for (int e = 0; e < 8; ++e) vv[e] += t8[e]; }
                            if (add2p) { float t8[8]; unpack8(ra2[h], t8);
#pragma unroll
                                for (int e = 0; e < 8; ++e) vv[e] += t8[e]; }
                            if (a.bcast) {
                                float t8[8];
                                load8f(a.bcast + ((long long)samp * Vout + jrow) * a.N + c + h * 8, t8, 0.f);
#pragma unroll
                                for (int e = 0; e < 8; ++e) vv[e] = fmaf(t8[e], a.bcast_scale, vv[e]);
                            }
                            if (a.has_mask) {
                                float m8[8];
                                tc::finish_smem(rm[h], a.mask.x2 != nullptr, 0, tl_ma1 + col, tl_mb + col, tl_ma2 + col, m8);
#pragma unroll
                                for (int e = 0; e < 8; ++e) vv[e] = m8[e] > 0.f ? vv[e] : 0.f;
                            }
                            o = pack8(vv);
                            if (STATS == 2) {
                                float pp[8];
                                if (partp) unpack8(rp[h], pp);
#pragma unroll
                                for (int e = 0; e < 8; ++e) vv[e] *= partp ? pp[e] : vv[e];
                                pr = pack8(vv);
                            }
                        }
                        const uint32_t off = (uint32_t)(col >> 6) * ATOM_BYTES + atom_off(r, (col & 63) >> 3);
                        *reinterpret_cast<uint4*>(Ot + off) = o;
                        if (STATS == 2) *reinterpret_cast<uint4*>(St + off) = pr;
                    }
                }
            }
            tc_fence_before();
            fence_async_smem();
            __syncwarp();
            if (lane == 0) {
                mbar_arrive(&bars.acc_free[buf]);
                mbar_arrive(&bars.out_ready[ob]);
                if (tst) mbar_arrive(&bars.tempty[tb]);
            }
        }
        if (STATS != 0 && n_my > 0) {
            // ---- per-channel sums of this CTA (lane = channel): out_free of the last tile = its statistics MMAs are complete
            const int last = n_my - 1;
            for (int j = 0; j < p.OB && j <= last; ++j) mbar_wait(&bars.out_free[(last - j) % p.OB], (uint32_t)(((last - j) / p.OB) & 1));
            tc_fence_after();
            if (half == 0) {
                const bool m64 = Ntp <= 64;
                int ch = -1;
                if (m64) { if (lane < 16) ch = q * 16 + lane; } else ch = q * 32 + lane;
                const uint32_t lanes = (uint32_t)(q * 32) << 16;
                float s1[8], s2v = 0.f;
                tmem_ld8(tmem + (uint32_t)p.sum_col + lanes, s1);
                if (STATS == 2) {
                    float s2[8];
                    tmem_ld8(tmem + (uint32_t)p.stat_col + lanes, s2);
                    s2v = s2[0];
                } else {
                    // diagonal of the Gram matrix: accumulator row = channel (this lane), column = the same channel
                    float g[32];
                    const int cb = m64 ? q * 16 : q * 32;
                    tmem_ld16(tmem + (uint32_t)p.stat_col + (uint32_t)cb + lanes, g);
                    if (!m64) tmem_ld16(tmem + (uint32_t)p.stat_col + (uint32_t)cb + 16u + lanes, g + 16);
#pragma unroll
                    for (int e = 0; e < 32; ++e)
                        if (e == lane && (e < 16 || !m64)) s2v = g[e];
                }
                if (ch >= 0 && ch < Nt) {
                    atomicAdd(a.stat_sum + n0 + ch, (double)s1[0]);
                    atomicAdd(a.stat_sq + n0 + ch, (double)s2v);
                }
            }
        }
    }
    if (warp == G4_DRAIN_WARP && lane == 0) {
        // ==================================================== drain =====================================================
        const int Ms = Ntp <= 64 ? 64 : 128;
        const uint32_t ones_d = smem_u32(ones);
        for (int i = 0; i < n_my; ++i) {
            const int tile = tile0 + i * tstep;
            const int ob = i % p.OB, useo = i / p.OB;
            mbar_wait_backoff(&bars.out_ready[ob], (uint32_t)(useo & 1), m4);
            tc_fence_after();
            const unsigned char* Ot = Osm + (size_t)ob * p.out_bytes;
            for (int oa = 0; oa < ((a.out_f32 || (p.dbg & 2)) ? 0 : n_oatoms); ++oa) {
                const unsigned char* src = Ot + (size_t)oa * ATOM_BYTES;
                if (p.mode == 0) tma_store_2d(&mapO, src, n0 + oa * ATOM_CH, tile * ATOM_ROWS);
                else if (p.mode >= 3) {
                    const int smp = tile / p.tps, q0 = (tile - smp * p.tps) * p.F;
                    for (int f = 0; f < p.F; ++f)
                        if (q0 + f < p.Tq) tma_store_4d(&mapO, src + (size_t)f * p.slot * 128, n0 + oa * ATOM_CH, 0, q0 + f, smp);
                } else
                    for (int f = 0; f < p.F; ++f)
                        if ((long long)tile * p.F + f < p.n_frames) tma_store_3d(&mapO, src + (size_t)f * p.slot * 128, n0 + oa * ATOM_CH, 0, tile * p.F + f);
            }
            tma_store_commit();
            if (STATS != 0 && (p.dbg & 1)) umma_commit(&bars.stat_done[ob]);
            if (STATS != 0 && !(p.dbg & 1)) {
                const uint32_t o0 = smem_u32(Ot), acc0 = i == 0 ? 0u : 1u;
                const uint32_t idesc_1 = idesc_major(Ms, 8, 1, 0);
                if (STATS == 1 && p.gram_ones) {
                    // one MMA sequence: D[channel][0..63] = Gram block, D[channel][64..79] = sum over the rows (B = [tile | ones atom])
                    const uint32_t idesc_go = idesc_major(Ms, 80, 1, 1), lbo = smem_u32(sm + p.off_ones16) - o0;
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(tmem + (uint32_t)p.stat_col, desc_mn_sw128(o0 + ks * 2048u, ATOM_BYTES), desc_mn_sw128(o0 + ks * 2048u, lbo), idesc_go,
                                 (acc0 | (uint32_t)ks) ? 1u : 0u);
                } else {
                for (int ks = 0; ks < 8; ++ks)
                    umma_f16(tmem + (uint32_t)p.sum_col, desc_mn_sw128(o0 + ks * 2048u, ATOM_BYTES), desc_ones(ones_d), idesc_1, (acc0 | (uint32_t)ks) ? 1u : 0u);
                if (STATS == 2) {
                    const uint32_t s0 = smem_u32(Ssm + (size_t)ob * p.out_bytes);
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(tmem + (uint32_t)p.stat_col, desc_mn_sw128(s0 + ks * 2048u, ATOM_BYTES), desc_ones(ones_d), idesc_1, (acc0 | (uint32_t)ks) ? 1u : 0u);
                } else {
                    const uint32_t idesc_g = idesc_major(Ms, Ntp, 1, 1);
                    for (int ks = 0; ks < 8; ++ks)
                        umma_f16(tmem + (uint32_t)p.stat_col, desc_mn_sw128(o0 + ks * 2048u, ATOM_BYTES), desc_mn_sw128(o0 + ks * 2048u, ATOM_BYTES), idesc_g,
                                 (acc0 | (uint32_t)ks) ? 1u : 0u);
                }
                }
                umma_commit(&bars.stat_done[ob]);
            }
            if (p.drain_defer) {
                // retire the PREVIOUS tile: its store has read the out tile (at most this tile's group is still pending) and its statistics
                // MMAs are complete.  The waits of a tile no longer sit between its own issue and the next tile's issue — the drain warp
                // was a serial ~0.6 us per tile (fixed cost of every tile: bench_gemm, 64- vs 128-channel layers).
                if (i > 0) {
                    const int pb = (i - 1) % p.OB, pu = (i - 1) / p.OB;
                    tma_store_wait_read<1>();
                    if (STATS != 0) mbar_wait(&bars.stat_done[pb], (uint32_t)(pu & 1));
                    mbar_arrive(&bars.out_free[pb]);
                }
                if (i == n_my - 1) {
                    tma_store_wait_read<0>();
                    if (STATS != 0) mbar_wait(&bars.stat_done[ob], (uint32_t)(useo & 1));
                    mbar_arrive(&bars.out_free[ob]);
                }
            } else {
                tma_store_wait_read<0>();
                if (STATS != 0) mbar_wait(&bars.stat_done[ob], (uint32_t)(useo & 1));
                mbar_arrive(&bars.out_free[ob]);
            }
        }
        tma_store_wait_all<0>();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// ---- host side -----------------------------------------------------------------------------------------------------
static inline long long tc4_wpack_bytes(int K, int N) {
    const long long kat = 2LL * ((K + ATOM_CH - 1) / ATOM_CH), t64 = (N + 63) / 64;
    return t64 * kat * 64 * 128 + (long long)((N + 63) & ~63) * 4 + 512;
}

// the distinct tail tensors of a call (a tensor that is both mask source and partner is staged once)
struct G4TailSel { const void* ptr[G4_MAX_TAILS]; long long ld[G4_MAX_TAILS]; int n; int idx[5]; bool ok; };
static inline G4TailSel tc4_tail_select(const void* add, long long ld_add, const void* add2, long long ld_add2, const void* partner, long long ld_partner,
                                        int has_mask, const dsg_act_src& mask) {
    G4TailSel s{};
    s.ok = tc4_tail_tma() != 0;
    const void* ptrs[5] = {add, add2, partner, has_mask ? mask.x1 : nullptr, has_mask ? mask.x2 : nullptr};
    const long long lds[5] = {ld_add, ld_add2, ld_partner, mask.ld1, mask.ld2};
    for (int k = 0; k < 5; ++k) {
        s.idx[k] = -1;
        if (!ptrs[k]) continue;
        if (!tma_ptr_ok(ptrs[k], lds[k])) { s.ok = false; continue; }
        for (int j = 0; j < s.n; ++j)
            if (s.ptr[j] == ptrs[k] && s.ld[j] == lds[k]) s.idx[k] = j;
        if (s.idx[k] >= 0) continue;
        if (s.n == G4_MAX_TAILS) { s.ok = false; continue; }
        s.ptr[s.n] = ptrs[k]; s.ld[s.n] = lds[k]; s.idx[k] = s.n++;
    }
    if (s.n == 0) s.ok = false;
    return s;
}
static inline void tc4_plan_tails(G4Plan& p, const G4TailSel& sel, int tn) {
    p.tn = tn; p.TB = G4_TB;
    p.t_add = tn ? sel.idx[0] : -1; p.t_add2 = tn ? sel.idx[1] : -1; p.t_part = tn ? sel.idx[2] : -1;
    p.t_m1 = tn ? sel.idx[3] : -1; p.t_m2 = tn ? sel.idx[4] : -1;
}

static bool tc4_plan(const dsg_conv_gemm_args& a, G4Plan& p, bool fold, const G4TailSel& sel, int tn) {
    p = G4Plan{};
    tc4_plan_tails(p, sel, tn);
    const int Vout = a.Vin + a.ext_in - a.contract_ext;
    p.mode = a.ext_in ? 1 : (a.contract_ext ? 2 : 0);
    p.V = a.ext_in ? a.Vin : (a.contract_ext ? a.Vin - 1 : a.Vin);
    p.n_frames = (long long)a.n_samples * a.T_out;
    p.rows_out = p.n_frames * Vout;
    if (p.mode == 0) {
        p.slot = 0; p.F = 0;
        const long long nt = (p.rows_out + ATOM_ROWS - 1) / ATOM_ROWS;
        if (nt > 0x3fffffff) return false;
        p.n_tiles = (int)nt;
    } else {
        if (p.V + 1 > 32 || p.V < 2) return false;
        p.slot = (p.V + 1 + 7) & ~7;
        p.F = ATOM_ROWS / p.slot;
        const long long nt = (p.n_frames + p.F - 1) / p.F;
        if (nt > 0x3fffffff) return false;
        p.n_tiles = (int)nt;
    }
    const int kat = (a.K + ATOM_CH - 1) / ATOM_CH;
    p.natoms1 = kat;
    p.natoms = (fold && a.src.x2) ? 2 * kat : kat;
    if (p.natoms > G4_MAX_ATOMS) return false;
    for (int i = 0; i < p.natoms; ++i) {
        const int k0 = (i % kat) * ATOM_CH;
        const int live = a.K - k0 < ATOM_CH ? a.K - k0 : ATOM_CH;
        p.ksteps[i] = (live + 15) / 16;
    }
    p.K1p = kat * ATOM_CH;
    p.act = (!fold) ? 1 : 0;                              // ReLU sources (one tensor): CUDA-core prologue in place
    // mode 2 with 32-row frame slots and no ReLU prologue: the fold-back is linear, so it is applied to the ACCUMULATOR rows in the
    // epilogue (a frame = one warp's 32 TMEM lanes: out[v] += out[V] / V is one shuffle per column) and no thread touches the operand
    p.xf = (p.act || p.mode == 1 || (p.mode == 2 && p.slot != 32)) ? 1 : 0;
    p.xfmap = tc4_xf_map();
    p.dbg = tc4_dbg();
    p.stats = a.stat_sum == nullptr ? 0 : (a.partner ? 2 : 1);
    const unsigned cf_bytes = (unsigned)((4 * 128 + 2 * p.K1p) * sizeof(float));
    const unsigned budget = 227u * 1024u - 2048u;         // dynamic shared memory we may ask for (static barriers + alignment slack kept)
    const int cand_nt[2] = {128, 64};
    for (int ci = (a.N <= 64 ? 1 : 0); ci < 2; ++ci) {
        const int Ntile = cand_nt[ci];
        const unsigned wb = (unsigned)p.natoms * Ntile * 128;
        const unsigned ob1 = (unsigned)(Ntile / ATOM_CH) * ATOM_BYTES;
        for (int OB = tc4_ob_max(); OB >= 1; --OB) {
            const unsigned tail_stage = (unsigned)tn * ob1;
            const bool gones = p.stats == 1 && Ntile == 64 && tc4_gram_ones();
            const unsigned fixed = wb + OB * ob1 * (p.stats == 2 ? 2u : 1u) + G4_TB * tail_stage + (gones ? (unsigned)ATOM_BYTES : 0u) + 1024u +
                                   ((cf_bytes + 1023u) & ~1023u) + 1024u;
            if (fixed + 3u * ATOM_BYTES > budget) continue;
            int S = (int)((budget - fixed) / ATOM_BYTES);
            if (S > tc4_s_cap()) S = tc4_s_cap();
            if (OB >= 2 && S < (tc4_s_cap() < 4 ? tc4_s_cap() : 4) && ci == 0) continue;    // prefer a deeper ring over multi-buffered out tiles
            p.Ntile = Ntile; p.S = S; p.OB = OB;
            p.drain_defer = (OB >= 2 && tc4_drain_defer()) ? 1 : 0;      // OB = 1: the epilogue needs the previous tile retired first
            p.off_w = 0;
            p.off_a = wb;                                  // multiples of 1024 throughout
            p.off_out = p.off_a + (unsigned)S * ATOM_BYTES;
            p.out_bytes = ob1;
            p.off_stat = p.off_out + OB * ob1;
            p.off_tail = p.off_stat + (p.stats == 2 ? OB * ob1 : 0u);
            p.tail_stage_bytes = tail_stage;
            p.gram_ones = gones ? 1 : 0;
            p.off_ones16 = p.off_tail + G4_TB * tail_stage;                // behind every out tile: the operand's atom stride is positive
            p.off_ones = p.off_ones16 + (gones ? (unsigned)ATOM_BYTES : 0u);
            p.off_cf = p.off_ones + 1024u;
            p.smem_total = p.off_cf + ((cf_bytes + 1023u) & ~1023u) + 1024u;
            p.w_tile_bytes = wb;
            p.acc_cols = Ntile <= 32 ? 32 : (Ntile <= 64 ? 64 : 128);
            p.stat_col = 2 * p.acc_cols;                                   // Gram: acc_cols columns; product sums: 8
            p.sum_col = p.stat_col + (gones ? 64 : (p.stats == 1 ? p.acc_cols : 8));      // [tile | ones]: sum x = column 64 of the Gram block
            const int need = p.stats ? p.sum_col + (gones ? 16 : 8) : 2 * p.acc_cols;
            p.tmem_cols = 32;
            while (p.tmem_cols < need) p.tmem_cols <<= 1;
            return true;
        }
    }
    return false;
}

// ---- fused adjacency contraction + 1x1 convolution (G4Plan mode 4)
static const char* launch_conv_gemm_tc4_fused(const dsg_conv_gemm_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (a.dtype != DSG_BF16 || a.taps != 1 || a.t_mul != 1 || a.t_div != 1 || a.tap_off != 0 || a.T_in != a.T_out) return nullptr;
    if (a.K % 8 != 0 || a.K > ATOM_CH || a.N % 8 != 0 || a.N < 16 || a.N > 128 || a.ext_in || a.contract_ext || a.src.x2 || a.src.a2 || a.out_f32) return nullptr;
    if (a.Vin < 2 || a.Vin > 32 || (uintptr_t)a.adyn % 16 != 0 || !a.wpack || (uintptr_t)a.wpack % 128 != 0) return nullptr;
    if (!tma_ptr_ok(a.src.x1, a.src.ld1) || !tma_ptr_ok(a.out, a.ld_out) || (a.y_out && !tma_ptr_ok(a.y_out, a.ld_y))) return nullptr;
    auto al16 = [](const void* q, long long ld) { return q == nullptr || ((uintptr_t)q % 16 == 0 && ld % 8 == 0); };
    if (!(al16(a.add, a.ld_add) && al16(a.add2, a.ld_add2) && al16(a.partner, a.ld_partner) && (!a.has_mask || act8_ok(a.mask)))) return nullptr;
    if (a.n_samples <= 0 || a.T_out <= 0) { *handled = true; return nullptr; }
    if (!encode_fn()) return nullptr;
    G4Plan p{};
    p.mode = 4;
    p.V = a.Vin;
    p.slot = (a.Vin + 7) & ~7;
    p.F = ATOM_ROWS / p.slot;
    p.qs = 1; p.qp = 0; p.Tdst = a.T_out; p.Tq = a.T_out;
    p.tps = (a.T_out + p.F - 1) / p.F;
    const long long nt = (long long)a.n_samples * p.tps;
    if (nt > 0x3fffffff) return nullptr;
    p.n_tiles = (int)nt;
    p.n_frames = (long long)a.n_samples * a.T_out;
    p.rows_out = p.n_frames * a.Vin;
    p.natoms = p.natoms1 = 1;
    p.ksteps[0] = (a.K + 15) / 16;
    p.K1p = ATOM_CH;
    p.act = 1; p.xf = 1;
    p.stats = a.stat_sum == nullptr ? 0 : (a.partner ? 2 : 1);
    p.store_y = a.y_out ? 1 : 0;
    {   // the contraction is compiled for the joint counts of the three layouts and the 3R of R = 8 / 16 (gcn_ratio = 0.125)
        const int vi = a.Vin == 25 ? 0 : (a.Vin == 17 ? 1 : (a.Vin == 18 ? 2 : -1)), ki = a.K == 24 ? 0 : (a.K == 48 ? 1 : -1);
        if (vi < 0 || ki < 0) return nullptr;
        p.cvar = vi * 2 + ki;
    }
    p.Ntile = a.N <= 64 ? 64 : 128;
    p.adj_bytes = ((unsigned)(a.Vin * a.Vin * a.K * 2) + 1023u) & ~1023u;
    const unsigned wb = (unsigned)p.Ntile * 128;
    const unsigned ob1 = (unsigned)(p.Ntile / ATOM_CH) * ATOM_BYTES;
    const unsigned cf_bytes = (unsigned)((4 * 128 + 2 * p.K1p) * sizeof(float));
    const unsigned budget = 227u * 1024u - 2048u;
    bool fit = false;
    for (int OB = 2; OB >= 1 && !fit; --OB) {
        const unsigned fixed = wb + OB * ob1 * (p.stats == 2 ? 2u : 1u) + 1024u + ((cf_bytes + 1023u) & ~1023u) + p.adj_bytes + 1024u;
        if (fixed + 2u * 2u * ATOM_BYTES > budget) continue;
        int S = (int)((budget - fixed) / (2u * ATOM_BYTES));
        if (S > 4) S = 4;
        p.S = S; p.OB = OB;
        p.drain_defer = (OB >= 2 && tc4_drain_defer()) ? 1 : 0;
        p.off_w = 0;
        p.off_a = wb;
        p.off_out = p.off_a + (unsigned)S * 2u * ATOM_BYTES;
        p.out_bytes = ob1;
        p.off_stat = p.off_out + OB * ob1;
        p.off_ones = p.off_stat + (p.stats == 2 ? OB * ob1 : 0u);
        p.off_cf = p.off_ones + 1024u;
        p.off_adj = p.off_cf + ((cf_bytes + 1023u) & ~1023u);
        p.smem_total = p.off_adj + p.adj_bytes + 1024u;
        fit = true;
    }
    if (!fit) return nullptr;
    p.w_tile_bytes = wb;
    p.acc_cols = p.Ntile <= 64 ? 64 : 128;
    p.stat_col = 2 * p.acc_cols;
    p.sum_col = p.stat_col + (p.stats == 1 ? p.acc_cols : 8);
    const int need = p.stats ? p.sum_col + 8 : 2 * p.acc_cols;
    p.tmem_cols = 32;
    while (p.tmem_cols < need) p.tmem_cols <<= 1;
    if ((long long)(wb + (size_t)a.N * 4 + 256) > tc4_wpack_bytes(a.K, a.N)) return nullptr;
    CUtensorMap mA0, mY, mO;
    bool ok = make_map_4d(&mA0, a.src.x1, a.n_samples, a.T_in, a.Vin, a.K, a.src.ld1, 0, 1);
    mY = mA0;
    if (ok && a.y_out) ok = make_map_4d(&mY, a.y_out, a.n_samples, a.T_out, a.Vin, a.K, a.ld_y, 0, 1);
    ok = ok && make_map_4d(&mO, a.out, a.n_samples, a.T_out, a.Vin, a.N, a.ld_out, 0, 1);
    if (!ok) return nullptr;
    float* cbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.wpack) + (((size_t)wb + 255) & ~(size_t)255));
    tc4_wpack_kernel<<<dim3(1, 2), dim3(256), 0, st>>>(a.W, a.ws_n, a.ws_k, a.K, a.N, 1, 1, p.Ntile, nullptr, nullptr, nullptr, nullptr, a.bias,
                                                         reinterpret_cast<unsigned char*>(a.wpack), cbias);
    if (const char* e = dsg_launch_error()) return e;
    int gx = num_sms();
    if (gx > p.n_tiles) gx = p.n_tiles;
    p.tpc = (p.n_tiles + gx - 1) / gx;
    gx = (p.n_tiles + p.tpc - 1) / p.tpc;
    const bool tails = a.add || a.add2 || a.bcast || a.has_mask || a.partner;
    if (p.stats == 2 && !tails) return "conv_gemm (fused contraction): statistics with a partner need the partner";
#define DSG_T4F_LAUNCH(TL_, ST_)                                                                                                   \
    do {                                                                                                                          \
        cudaFuncSetAttribute(tc4_gemm_kernel<true, TL_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_total);    \
        tc4_gemm_kernel<true, TL_, ST_><<<dim3((unsigned)gx, 1), dim3(G4_THREADS), p.smem_total, st>>>(mA0, mY, mO, a, p, cbias, G4TailMaps{}); \
    } while (0)
    if (!tails && p.stats == 0) DSG_T4F_LAUNCH(false, 0);
    else if (!tails && p.stats == 1) DSG_T4F_LAUNCH(false, 1);
    else if (tails && p.stats == 0) DSG_T4F_LAUNCH(true, 0);
    else if (tails && p.stats == 1) DSG_T4F_LAUNCH(true, 1);
    else DSG_T4F_LAUNCH(true, 2);
#undef DSG_T4F_LAUNCH
    *handled = true;
    return dsg_launch_error();
}

static const char* launch_conv_gemm_tc4(const dsg_conv_gemm_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (a.adyn) return tc4_enabled() ? launch_conv_gemm_tc4_fused(a, st, handled) : nullptr;
    if (!tc4_enabled() || a.dtype != DSG_BF16 || a.taps != 1 || a.t_mul != 1 || a.t_div != 1 || a.tap_off != 0) return nullptr;
    if (a.K % 8 != 0 || a.N % 8 != 0 || a.N < 16 || a.T_in != a.T_out) return nullptr;
    if (a.ext_in && a.contract_ext) return nullptr;
    if (!a.wpack || (uintptr_t)a.wpack % 128 != 0) return nullptr;
    if (!tma_ptr_ok(a.src.x1, a.src.ld1) || (a.src.x2 && !tma_ptr_ok(a.src.x2, a.src.ld2))) return nullptr;
    if (a.out_f32 ? ((uintptr_t)a.out % 16 != 0 || a.ld_out % 4 != 0 || a.add || a.add2 || a.bcast || a.has_mask || a.partner || a.stat_sum)
                  : !tma_ptr_ok(a.out, a.ld_out))
        return nullptr;
    const bool fold = !a.src.relu;                        // affine / two-tensor BN-backward prologue folds into the weights
    if (!fold && a.src.x2) return nullptr;                // ReLU over two tensors: older engines
    auto al16 = [](const void* p, long long ld) { return p == nullptr || ((uintptr_t)p % 16 == 0 && ld % 8 == 0); };
    if (!(al16(a.add, a.ld_add) && al16(a.add2, a.ld_add2) && al16(a.partner, a.ld_partner) && (!a.has_mask || act8_ok(a.mask)) &&
          (!a.bcast || (uintptr_t)a.bcast % 16 == 0)))
        return nullptr;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0) { *handled = true; return nullptr; }
    G4Plan p;
    const G4TailSel sel = a.out_f32 ? G4TailSel{} : tc4_tail_select(a.add, a.ld_add, a.add2, a.ld_add2, a.partner, a.ld_partner, a.has_mask, a.mask);
    if (!(sel.ok && tc4_plan(a, p, fold, sel, sel.n)) && !tc4_plan(a, p, fold, sel, 0)) return nullptr;
    if (!encode_fn()) return nullptr;
    CUtensorMap mA0, mA1, mO;
    G4TailMaps tmaps;
    const long long rows_in = n_frames * a.Vin;
    bool ok;
    if (p.mode == 0) {
        ok = make_map_2d(&mA0, a.src.x1, rows_in, a.K, a.src.ld1, ATOM_ROWS);
        mA1 = mA0;
        if (ok && p.natoms > p.natoms1) ok = make_map_2d(&mA1, a.src.x2, rows_in, a.K, a.src.ld2, ATOM_ROWS);
        if (a.out_f32) mO = mA0;                               // never dereferenced: rows go out through plain stores
        else ok = ok && make_map_2d(&mO, a.out, p.rows_out, a.N, a.ld_out, ATOM_ROWS);
        for (int t = 0; t < p.tn && ok; ++t) ok = make_map_2d(&tmaps.m[t], sel.ptr[t], p.rows_out, a.N, sel.ld[t], ATOM_ROWS);
    } else {
        const int rin = a.Vin, rout = a.Vin + a.ext_in - a.contract_ext;
        ok = make_map_3d(&mA0, a.src.x1, n_frames, rin, a.K, a.src.ld1, rin, 1);
        mA1 = mA0;
        if (ok && p.natoms > p.natoms1) ok = make_map_3d(&mA1, a.src.x2, n_frames, rin, a.K, a.src.ld2, rin, 1);
        if (a.out_f32) mO = mA0;
        else ok = ok && make_map_3d(&mO, a.out, n_frames, rout, a.N, a.ld_out, rout, 1);
        for (int t = 0; t < p.tn && ok; ++t) ok = make_map_3d(&tmaps.m[t], sel.ptr[t], n_frames, rout, a.N, sel.ld[t], rout, 1);
    }
    if (!ok) return nullptr;
    const unsigned gy = (unsigned)((a.N + p.Ntile - 1) / p.Ntile);
    const size_t wtot = (size_t)gy * p.w_tile_bytes;
    if ((long long)(wtot + (size_t)a.N * 4 + 256) > tc4_wpack_bytes(a.K, a.N)) return nullptr;
    float* cbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.wpack) + ((wtot + 255) & ~(size_t)255));
    const float* s1 = fold ? a.src.a1 : nullptr;
    const float* s2 = fold ? a.src.a2 : nullptr;
    const float* c1 = fold ? a.src.b1 : nullptr;
    const float* c2 = fold ? a.src.b2 : nullptr;
    tc4_wpack_kernel<<<dim3(gy, (unsigned)p.natoms + 1), dim3(256), 0, st>>>(a.W, a.ws_n, a.ws_k, a.K, a.N, p.natoms1, p.natoms, p.Ntile, s1, s2, c1, c2,
                                                                            a.bias, reinterpret_cast<unsigned char*>(a.wpack), cbias,
                                                                            p.mode == 2 ? 1.f + 1.f / (float)p.V : 1.f);
    if (const char* e = dsg_launch_error()) return e;
    int gx = num_sms() / (int)gy;
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    const bool tails = a.add || a.add2 || a.bcast || a.has_mask || a.partner;
    const int variant = (p.xf ? 6 : 0) + (tails ? 3 : 0) + p.stats;
#define DSG_T4_LAUNCH(XF_, TL_, ST_)                                                                                              \
    do {                                                                                                                          \
        cudaFuncSetAttribute(tc4_gemm_kernel<XF_, TL_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_total);     \
        tc4_gemm_kernel<XF_, TL_, ST_><<<dim3((unsigned)gx, gy), dim3(G4_THREADS), p.smem_total, st>>>(mA0, mA1, mO, a, p, cbias, tmaps); \
    } while (0)
    switch (variant) {
        case 0: DSG_T4_LAUNCH(false, false, 0); break;
        case 1: DSG_T4_LAUNCH(false, false, 1); break;
        case 3: DSG_T4_LAUNCH(false, true, 0); break;
        case 4: DSG_T4_LAUNCH(false, true, 1); break;
        case 5: DSG_T4_LAUNCH(false, true, 2); break;
        case 6: DSG_T4_LAUNCH(true, false, 0); break;
        case 7: DSG_T4_LAUNCH(true, false, 1); break;
        case 9: DSG_T4_LAUNCH(true, true, 0); break;
        case 10: DSG_T4_LAUNCH(true, true, 1); break;
        case 11: DSG_T4_LAUNCH(true, true, 2); break;
        default: return "tc4: statistics with a partner need a tail operand";
    }
#undef DSG_T4_LAUNCH
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc4
}  // namespace dsg
#endif
