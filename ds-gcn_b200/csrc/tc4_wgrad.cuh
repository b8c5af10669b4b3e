// tc4 weight gradient: TMA-fed, warp-specialised, persistent tcgen05 engine of dsg_conv_wgrad for the 1x1 convolutions
// (taps == 1) — north-star kernel (d), the backward twin of tc4_gemm.cuh.
//
//   dW[n,k] = sum_rows dy[row,n] * x[row,k],   dy = ca[n]*e + cb[n]*y + cc[n]   (BatchNorm-backward form, dsg_act_src of B)
//
// The reduction runs over the ROWS of the activation tiles, so both operands are MN-major: the very atoms the TMA unit
// writes ([128 rows x 64 channels], SWIZZLE_128B) are read by the tensor core with the reduction along the row axis —
// nothing is transposed or re-staged by a thread.  The BatchNorm-backward combination is linear, so it is applied AFTER
// the reduction, to the accumulators, instead of to every element:
//
//   dW = diag(ca) (e^T x)  +  diag(cb) (y^T x)  +  cc (1^T x)          db = ca*(1^T e) + cb*(1^T y) + cc*rows
//
// e^T x and y^T x are two TMEM accumulators ([<=128 dy channels] x [<=128 / 256 x channels]); the column sums 1^T e, 1^T y,
// 1^T x come from "ones" MMAs (N = 8) over the same atoms.  e, y and x go from HBM to the tensor core untouched.
//
//   warp 0        TMA producer (one elected thread): per 128-row tile the x atoms of this CTA's input-channel tile and the
//                 e / y atoms of its output-channel tile, S-stage mbarrier ring (up to 192 KB in flight per SM)
//   warp 1        MMA issuer: 8 K-steps of 16 rows per tile and accumulator; tcgen05.commit frees the stage
//   warps 2-5     transform warps, only when x is not a plain tensor: BatchNorm-affine + ReLU in place (x of the `transform`
//                 conv is relu(bn(feat)), tcn.py:393-395) and the joint-mean row of dgmstcn (tcn.py:409)
//   warps 6-9     final epilogue, once per CTA: accumulators -> ca/cb/cc combination -> red.global.add.v4.f32 into dW / db
#pragma once
#include "tc4_common.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc4 {

constexpr int W4_XF_WARPS = 4, W4_EPI_WARPS = 4;
constexpr int W4_THREADS = 32 * (2 + W4_XF_WARPS + W4_EPI_WARPS);        // 320
constexpr int W4_XF_T0 = 64, W4_EPI_T0 = 64 + 32 * W4_XF_WARPS;
constexpr int W4_MAX_STAGES = 6;
constexpr int W4_BAR_XF = 3, W4_BAR_EPI = 4;

struct W4Plan {
    int mode;                    // 0: plain rows (2-D maps, dense 128-row tiles); 1: frame slots (3-D maps, F frames x `slot` rows)
    int Vx, Vb, slot, F;         // rows per frame of x / of dy, rows per frame slot (8-aligned), frames per tile
    int n_tiles, ksteps;         // 128-row tiles; 16-row MMA steps per tile
    long long rows_b, n_frames;
    int Kt_max, ka_max, na_max, two;   // x channels per CTA, x atoms, dy atoms per CTA, dy = f(e, y)
    int S;
    unsigned stage_bytes, off_ones, off_xs, smem_total;
    int xf_act, xf_mean, need_xs, do_bias;
    int xfmap;                   // affine prologue mapping: 1 = thread owns a 16-byte chunk column (as tc4_gemm.cuh)
    int colY, colS, tmem_cols;   // TMEM columns: e^T x at 0, y^T x at colY, sums at colS (+0 e, +8 y, +16 / +24 x)
};

struct W4Bars {
    uint64_t full[W4_MAX_STAGES], empty[W4_MAX_STAGES], ready[W4_MAX_STAGES];
    uint64_t done;
};

__global__ void __launch_bounds__(W4_THREADS, 1)
tc4_wgrad_kernel(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapE, const __grid_constant__ CUtensorMap mapY,
                 const dsg_conv_wgrad_args a, const W4Plan p) {
    DSG_DYN_SMEM(smem_raw);
    __shared__ W4Bars bars;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    unsigned char* sm = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    unsigned char* ring = sm;
    unsigned char* ones = sm + p.off_ones;
    float* xs_s = reinterpret_cast<float*>(sm + p.off_xs);          // [256] column sums of x, then [256] a1 | [256] b of the x prologue
    float* cf_a = xs_s + 256;
    float* cf_b = cf_a + 256;

    const int n0 = blockIdx.y * 128, k0 = blockIdx.z * p.Kt_max;
    const int Nt = a.N - n0 < 128 ? a.N - n0 : 128;
    const int na = (Nt + ATOM_CH - 1) / ATOM_CH;
    const int M = na == 1 ? 64 : 128;
    const int Kt = a.K - k0 < p.Kt_max ? a.K - k0 : p.Kt_max;
    const int ka = (Kt + ATOM_CH - 1) / ATOM_CH;
    const int Ktp = (Kt + 15) & ~15;
    const int nxs = (Ktp + 127) / 128;
    const bool bias_cta = p.do_bias && blockIdx.z == 0;

    // ---- one-time setup
    if (p.xf_act)
        for (int k = tid; k < 256; k += W4_THREADS) {
            const int ch = k0 + k;
            const bool in = k < Kt;
            cf_a[k] = (in && a.A.a1) ? a.A.a1[ch] : 1.f;
            cf_b[k] = ((in && a.A.b1) ? a.A.b1[ch] : 0.f) + ((in && a.A.b2) ? a.A.b2[ch] : 0.f);
        }
    for (int i = tid; i < 256; i += W4_THREADS) reinterpret_cast<uint16_t*>(ones)[i] = 0x3F80;      // bf16 1.0
    if (p.mode == 1) {
        // padding rows of the frame slots are never written by the TMA unit: they must read as zero (reduction over rows)
        const unsigned n16 = (unsigned)p.S * p.stage_bytes / 16u;
        for (unsigned i = tid; i < n16; i += W4_THREADS) reinterpret_cast<uint4*>(ring)[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (tid == 0) {
        for (int s = 0; s < W4_MAX_STAGES; ++s) { mbar_init(&bars.full[s], 1); mbar_init(&bars.empty[s], 1); mbar_init(&bars.ready[s], 32 * W4_XF_WARPS); }
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
    const bool xf = p.xf_act || p.xf_mean;
    const unsigned off_e = (unsigned)p.ka_max * ATOM_BYTES, off_y = off_e + (unsigned)p.na_max * ATOM_BYTES;

    if (warp == 0) {
        // ================================================= TMA producer =================================================
        if (lane == 0) {
            prefetch_map(&mapX);
            prefetch_map(&mapE);
            if (p.two) prefetch_map(&mapY);
            const uint32_t bx = p.mode == 0 ? (uint32_t)ATOM_BYTES : (uint32_t)(p.F * p.Vx * 128);
            const uint32_t bb = p.mode == 0 ? (uint32_t)ATOM_BYTES : (uint32_t)(p.F * p.Vb * 128);
            const uint32_t tx = (uint32_t)ka * bx + (uint32_t)(na * (1 + p.two)) * bb;
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                mbar_wait(&bars.empty[stage], ph ^ 1);
                mbar_expect_tx(&bars.full[stage], tx);
                unsigned char* st = ring + (size_t)stage * p.stage_bytes;
                for (int ai = 0; ai < ka + na * (1 + p.two); ++ai) {
                    const CUtensorMap* m;
                    unsigned char* dst;
                    int c0;
                    if (ai < ka) { m = &mapX; dst = st + (size_t)ai * ATOM_BYTES; c0 = k0 + ai * ATOM_CH; }
                    else if (ai < ka + na) { m = &mapE; dst = st + off_e + (size_t)(ai - ka) * ATOM_BYTES; c0 = n0 + (ai - ka) * ATOM_CH; }
                    else { m = &mapY; dst = st + off_y + (size_t)(ai - ka - na) * ATOM_BYTES; c0 = n0 + (ai - ka - na) * ATOM_CH; }
                    if (p.mode == 0) tma_load_2d(dst, m, c0, tile * ATOM_ROWS, &bars.full[stage]);
                    else
                        for (int f = 0; f < p.F; ++f) tma_load_3d(dst + (size_t)f * p.slot * 128, m, c0, 0, tile * p.F + f, &bars.full[stage]);
                }
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ================================================== MMA issuer ==================================================
        if (lane == 0 && n_my > 0) {
            const uint32_t idesc_w = idesc_major(M, Ktp, 1, 1);
            const uint32_t idesc_s = idesc_major(M, 8, 1, 0);
            const uint32_t ones_d = smem_u32(ones);
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                mbar_wait(xf ? &bars.ready[stage] : &bars.full[stage], ph);
                tc_fence_after();
                const uint32_t x0 = smem_u32(ring + (size_t)stage * p.stage_bytes), e0 = x0 + off_e, y0 = x0 + off_y;
                for (int ks = 0; ks < p.ksteps; ++ks) {
                    const uint32_t acc = (i | ks) ? 1u : 0u;
                    const uint64_t xd = desc_mn_sw128(x0 + ks * 2048u, ATOM_BYTES);
                    const uint64_t ed = desc_mn_sw128(e0 + ks * 2048u, ATOM_BYTES);
                    umma_f16(tmem, ed, xd, idesc_w, acc);
                    if (p.two) {
                        const uint64_t yd = desc_mn_sw128(y0 + ks * 2048u, ATOM_BYTES);
                        umma_f16(tmem + (uint32_t)p.colY, yd, xd, idesc_w, acc);
                        if (bias_cta) umma_f16(tmem + (uint32_t)p.colS + 8u, yd, desc_ones(ones_d), idesc_s, acc);
                    }
                    if (bias_cta) umma_f16(tmem + (uint32_t)p.colS, ed, desc_ones(ones_d), idesc_s, acc);
                    if (p.need_xs)
                        for (int j = 0; j < nxs; ++j) {
                            const int Mx = Ktp - j * 128 <= 64 ? 64 : 128;
                            umma_f16(tmem + (uint32_t)p.colS + 16u + 8u * j, desc_mn_sw128(x0 + (uint32_t)j * 2u * ATOM_BYTES + ks * 2048u, ATOM_BYTES),
                                     desc_ones(ones_d), idesc_major(Mx, 8, 1, 0), acc);
                        }
                }
                umma_commit(&bars.empty[stage]);
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
            umma_commit(&bars.done);
        }
    } else if (warp < 2 + W4_XF_WARPS) {
        // ================================================ transform warps ===============================================
        if (xf) {
            const int t = tid - W4_XF_T0;                 // 0..127 = tile row
            const long long rows_x = p.mode == 0 ? p.rows_b : 0;
            int stage = 0;
            uint32_t ph = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                mbar_wait(&bars.full[stage], ph);
                unsigned char* st = ring + (size_t)stage * p.stage_bytes;
                if (p.xf_act) {
                    bool live;
                    if (p.mode == 0) live = (long long)tile * ATOM_ROWS + t < rows_x;
                    else {
                        const int f = t / p.slot, v = t - f * p.slot;
                        live = f < p.F && v < p.Vx && (long long)tile * p.F + f < p.n_frames;
                    }
                    if (p.xfmap) {
                        // thread = (16-byte chunk column, row residue): coefficients in registers, 8 independent rows per thread
                        const int c = t & 7, r0 = t >> 3;
                        const float lo = a.A.relu ? 0.f : -3.0e38f;
                        for (int ai = 0; ai < ka; ++ai) {
                            const int k = ai * ATOM_CH + c * 8;
                            if (k >= Kt) break;
                            float ca8[8], cb8[8];
                            load8f(cf_a + k, ca8, 1.f);
                            load8f(cf_b + k, cb8, 0.f);
                            unsigned char* col = st + (size_t)ai * ATOM_BYTES + atom_off(r0, c);
                            uint4 qv[8];
#pragma unroll
                            for (int j = 0; j < 8; ++j) qv[j] = *reinterpret_cast<const uint4*>(col + j * 2048);
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                const int row = r0 + 16 * j;
                                bool lv;
                                if (p.mode == 0) lv = (long long)tile * ATOM_ROWS + row < rows_x;
                                else {
                                    const int f = row / p.slot, v = row - f * p.slot;
                                    lv = f < p.F && v < p.Vx && (long long)tile * p.F + f < p.n_frames;
                                }
                                float x[8];
                                unpack8(qv[j], x);
#pragma unroll
                                for (int e = 0; e < 8; ++e) x[e] = fmaxf(fmaf(x[e], ca8[e], cb8[e]), lo);
                                if (lv) *reinterpret_cast<uint4*>(col + j * 2048) = pack8(x);
                            }
                        }
                    } else if (live)
                        for (int ai = 0; ai < ka; ++ai) {
                            unsigned char* atom = st + (size_t)ai * ATOM_BYTES;
#pragma unroll 2
                            for (int c = 0; c < 8; ++c) {
                                const int k = ai * ATOM_CH + c * 8;
                                if (k >= Kt) break;
                                uint4* q = reinterpret_cast<uint4*>(atom + atom_off(t, c));
                                float x[8], ca8[8], cb8[8];
                                unpack8(*q, x);
                                load8f(cf_a + k, ca8, 1.f);
                                load8f(cf_b + k, cb8, 0.f);
#pragma unroll
                                for (int e = 0; e < 8; ++e) x[e] = fmaf(x[e], ca8[e], cb8[e]);
                                if (a.A.relu) {
#pragma unroll
                                    for (int e = 0; e < 8; ++e) x[e] = fmaxf(x[e], 0.f);
                                }
                                *q = pack8(x);
                            }
                        }
                    if (p.xf_mean) named_sync(W4_BAR_XF, 32 * W4_XF_WARPS);
                }
                if (p.xf_mean) {
                    // joint-mean row of every frame (tcn.py:409): item = (atom, frame, chunk), 4 lanes per item split the joints
                    const int items = ka * p.F * 8 * 4;
                    const float inv = 1.f / (float)p.Vx;
                    for (int base = 0; base < items; base += 32 * W4_XF_WARPS) {
                        const int idx = base + t;
                        const bool ok = idx < items;
                        const int part = idx & 3, c = (idx >> 2) & 7, fa = idx >> 5, f = fa % p.F, ai = fa / p.F;
                        unsigned char* atom = st + (size_t)(ok ? ai : 0) * ATOM_BYTES;
                        float s8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) s8[e] = 0.f;
                        if (ok && ai * ATOM_CH + c * 8 < Kt)
                            for (int v = part; v < p.Vx; v += 4) {
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
#pragma unroll
                            for (int e = 0; e < 8; ++e) s8[e] *= inv;
                            *reinterpret_cast<uint4*>(atom + atom_off(f * p.slot + p.Vx, c)) = pack8(s8);
                        }
                    }
                }
                fence_async_smem();
                mbar_arrive(&bars.ready[stage]);
                if (++stage == p.S) { stage = 0; ph ^= 1; }
            }
        }
    } else {
        // ================================================ final epilogue ================================================
        if (n_my > 0) {
            const int q = warp & 3;                       // TMEM lane quarter this warp may read
            const uint32_t lanes = (uint32_t)(q * 32) << 16;
            mbar_wait(&bars.done, 0);
            tc_fence_after();
            if (p.need_xs) {
                for (int j = 0; j < nxs; ++j) {
                    const int Mx = Ktp - j * 128 <= 64 ? 64 : 128;
                    float t8[8];
                    tmem_ld8(tmem + (uint32_t)p.colS + 16u + 8u * j + lanes, t8);
                    const int ch = Mx == 128 ? q * 32 + lane : (lane < 16 ? q * 16 + lane : -1);
                    if (ch >= 0) xs_s[j * 128 + ch] = t8[0];
                }
                named_sync(W4_BAR_EPI, 32 * W4_EPI_WARPS);
            }
            const int nch = M == 128 ? q * 32 + lane : (lane < 16 ? q * 16 + lane : -1);
            const bool valid = nch >= 0 && nch < Nt;
            const int n = n0 + (valid ? nch : 0);
            const float ca = a.B.a1 ? a.B.a1[n] : 1.f;
            const float cb = p.two ? (a.B.a2 ? a.B.a2[n] : 1.f) : 0.f;
            const float cc = (a.B.b1 ? a.B.b1[n] : 0.f) + (a.B.b2 ? a.B.b2[n] : 0.f);
            const bool vecW = a.ws_k == 1 && a.ws_n % 4 == 0 && (uintptr_t)a.dW % 16 == 0;
            for (int g = 0; g < (Ktp >> 4); ++g) {
                float ve[16], vy[16];
                tmem_ld16(tmem + (uint32_t)(g * 16) + lanes, ve);
                if (p.two) tmem_ld16(tmem + (uint32_t)p.colY + (uint32_t)(g * 16) + lanes, vy);
                if (!valid) continue;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float v = ca * ve[j];
                    if (p.two) v = fmaf(cb, vy[j], v);
                    if (p.need_xs) v = fmaf(cc, xs_s[g * 16 + j], v);
                    ve[j] = v;
                }
                float* dst = a.dW + (long long)n * a.ws_n + (long long)(k0 + g * 16) * a.ws_k;
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    if (g * 16 + j >= Kt) break;
                    if (vecW && g * 16 + j + 4 <= Kt) tc::red_add_v4(dst + j, ve[j], ve[j + 1], ve[j + 2], ve[j + 3]);
                    else
                        for (int e = 0; e < 4; ++e)
                            if (g * 16 + j + e < Kt) atomicAdd(dst + (long long)(j + e) * a.ws_k, ve[j + e]);
                }
            }
            if (bias_cta) {
                float se[8], sy[8];
                tmem_ld8(tmem + (uint32_t)p.colS + lanes, se);
                sy[0] = 0.f;
                if (p.two) tmem_ld8(tmem + (uint32_t)p.colS + 8u + lanes, sy);
                if (valid) {
                    // rows this CTA reduced over (the tail tile is partial)
                    const bool has_last = ((p.n_tiles - 1 - (int)blockIdx.x) % (int)gridDim.x) == 0;
                    float rows;
                    if (p.mode == 0) rows = (float)((long long)n_my * ATOM_ROWS - (has_last ? (long long)p.n_tiles * ATOM_ROWS - p.rows_b : 0));
                    else rows = (float)(((long long)n_my * p.F - (has_last ? (long long)p.n_tiles * p.F - p.n_frames : 0)) * p.Vb);
                    atomicAdd(a.db + n, fmaf(ca, se[0], fmaf(cb, sy[0], cc * rows)));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, (uint32_t)p.tmem_cols);
}

// 3-D view (C, rows_per_frame, frames) whose frame pitch is `t_mul` frames (strided block residual, dgstgcn.py:56-59)
static inline bool make_map_3d_strided(CUtensorMap* m, const void* base, long long frames, int rpf, int C, long long ld, int t_mul) {
    EncodeTiledFn enc = encode_fn();
    if (!enc || frames <= 0 || C <= 0 || rpf <= 0) return false;
    cuuint64_t dims[3] = {(cuuint64_t)C, (cuuint64_t)rpf, (cuuint64_t)frames};
    cuuint64_t strides[2] = {(cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)rpf * (cuuint64_t)t_mul};
    cuuint32_t box[3] = {(cuuint32_t)ATOM_CH, (cuuint32_t)rpf, 1};
    cuuint32_t es[3] = {1, 1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static const char* launch_conv_wgrad_tc4(const dsg_conv_wgrad_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (!tc4_enabled() || a.dtype != DSG_BF16 || a.taps != 1 || a.t_div != 1 || a.tap_off != 0) return nullptr;
    if (a.K % 8 != 0 || a.N % 8 != 0 || a.K < 8 || a.N < 8 || a.t_mul < 1) return nullptr;
    if (a.A.x2 || a.A.a2 || a.B.relu) return nullptr;
    if (!tma_ptr_ok(a.A.x1, a.A.ld1) || !tma_ptr_ok(a.B.x1, a.B.ld1) || (a.B.x2 && !tma_ptr_ok(a.B.x2, a.B.ld2))) return nullptr;
    if ((long long)a.T_out * a.t_mul != a.T_in && !(a.t_mul == 1 && a.T_in == a.T_out)) return nullptr;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0) { *handled = true; return nullptr; }
    if (!encode_fn()) return nullptr;
    W4Plan p{};
    p.mode = (a.ext_in || a.t_mul != 1) ? 1 : 0;
    p.Vx = a.Vin;
    p.Vb = a.Vin + a.ext_in;
    p.n_frames = n_frames;
    p.rows_b = n_frames * p.Vb;
    if (p.mode == 0) {
        p.slot = 0; p.F = 0;
        const long long nt = (p.rows_b + ATOM_ROWS - 1) / ATOM_ROWS;
        if (nt > 0x3fffffff) return nullptr;
        p.n_tiles = (int)nt;
        p.ksteps = 8;
    } else {
        if (p.Vb > 32 || p.Vx < 1) return nullptr;
        p.slot = (p.Vb + 7) & ~7;
        p.F = ATOM_ROWS / p.slot;
        const long long nt = (n_frames + p.F - 1) / p.F;
        if (nt > 0x3fffffff) return nullptr;
        p.n_tiles = (int)nt;
        p.ksteps = (p.F * p.slot + 15) / 16;
    }
    p.two = a.B.x2 ? 1 : 0;
    p.Kt_max = p.two ? 128 : 256;
    if (((a.K + ATOM_CH - 1) & ~(ATOM_CH - 1)) < p.Kt_max) p.Kt_max = (a.K + ATOM_CH - 1) & ~(ATOM_CH - 1);
    p.ka_max = p.Kt_max / ATOM_CH;
    p.na_max = a.N > ATOM_CH ? 2 : 1;
    p.xf_act = (a.A.a1 || a.A.b1 || a.A.b2 || a.A.relu) ? 1 : 0;
    p.xfmap = tc4_xf_map();
    p.xf_mean = a.ext_in ? 1 : 0;
    p.need_xs = (a.B.b1 || a.B.b2) ? 1 : 0;
    p.do_bias = a.db ? 1 : 0;
    p.stage_bytes = (unsigned)(p.ka_max + p.na_max * (1 + p.two)) * ATOM_BYTES;
    const unsigned fixed = 1024u /* ones */ + 3u * 1024u /* xs + prologue coefficients */ + 1024u /* alignment slack */;
    const unsigned budget = 227u * 1024u - 2048u;
    int S = (int)((budget - fixed) / p.stage_bytes);
    if (S > W4_MAX_STAGES) S = W4_MAX_STAGES;
    if (S < 2) return nullptr;
    p.S = S;
    p.off_ones = (unsigned)S * p.stage_bytes;
    p.off_xs = p.off_ones + 1024u;
    p.smem_total = p.off_xs + 3u * 1024u + 1024u;
    p.colY = p.Kt_max;
    p.colS = p.Kt_max * (1 + p.two);
    const int need = p.colS + 32;
    p.tmem_cols = 32;
    while (p.tmem_cols < need) p.tmem_cols <<= 1;
    if (p.tmem_cols > 512) return nullptr;

    CUtensorMap mX, mE, mY;
    bool ok;
    if (p.mode == 0) {
        ok = make_map_2d(&mX, a.A.x1, p.rows_b, a.K, a.A.ld1, ATOM_ROWS) && make_map_2d(&mE, a.B.x1, p.rows_b, a.N, a.B.ld1, ATOM_ROWS);
        mY = mE;
        if (ok && p.two) ok = make_map_2d(&mY, a.B.x2, p.rows_b, a.N, a.B.ld2, ATOM_ROWS);
    } else {
        ok = make_map_3d_strided(&mX, a.A.x1, n_frames, p.Vx, a.K, a.A.ld1, a.t_mul) && make_map_3d_strided(&mE, a.B.x1, n_frames, p.Vb, a.N, a.B.ld1, 1);
        mY = mE;
        if (ok && p.two) ok = make_map_3d_strided(&mY, a.B.x2, n_frames, p.Vb, a.N, a.B.ld2, 1);
    }
    if (!ok) return nullptr;
    const unsigned gy = (unsigned)((a.N + 127) / 128), gz = (unsigned)((a.K + p.Kt_max - 1) / p.Kt_max);
    int gx = num_sms() / (int)(gy * gz);
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    cudaFuncSetAttribute(tc4_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_total);
    tc4_wgrad_kernel<<<dim3((unsigned)gx, gy, gz), dim3(W4_THREADS), p.smem_total, st>>>(mX, mE, mY, a, p);
    *handled = true;
    return dsg_launch_error();
}

}  // namespace tc4
}  // namespace dsg
#endif
