// Fused branch stage of the multi-scale temporal unit (mstcn / dgmstcn) on tcgen05 — reference tcn.py:383-396,
// 407-420.  Implicit GEMM: for one sample and MS_TO = 4 output frames the post-BN-ReLU branch tile (with its temporal
// halo) is staged ONCE in shared memory as a K-major no-swizzle UMMA operand whose rows are (frame, joint) with every
// frame padded to 32 rows; tap dt of the dilated (3 x 1) convolution is then just a *shifted descriptor* into that
// tile (frame shifts are multiples of 32 rows = 4 core-matrix groups), so nothing is re-staged per tap.  With a
// temporal stride s the frames are staged de-interleaved in s planes (t = s*q + p) so every tap still addresses a
// contiguous run of frames.  All conv branches accumulate into disjoint TMEM column ranges of one 128-lane
// accumulator; the max-pool and pass-through branches run on the CUDA cores; the CTA assembles complete `feat`
// rows in shared memory (local + global * add_coeff) and writes them with 16-byte stores, accumulating the
// BatchNorm statistics of transform.0 on the way out.
#pragma once
#include "conv_gemm_tc.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc {

constexpr int MS_TO = 4;            // output (fwd) / input (bwd) frames per CTA: M = 128 = 4 frames x 32 padded rows
constexpr int MS_THREADS = 256;

DSG_HD int floordiv(int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); }
DSG_HD int posmod(int a, int b) { int m = a % b; return m < 0 ? m + b : m; }

// A conv branch owns channels [lo,hi).  Its UMMA operands use the *8-aligned channel window* [lo8, lo8 + 8*nchw) that
// covers it, so staging is whole 16-byte chunks; the few foreign channels inside the window meet zero weight rows.
struct MsBranchGeom {
    int w, lo8, off, nchw, Kp, nch, d, qmin, Fq, col;   // width, window start, lo-lo8, window chunks, padded K (=N), chunks, dilation, ...
    int nsh;                                            // log2(nch) when nch is a power of two, else -1
};
// geometry of every branch, computed once on the host and passed by value (no per-thread recomputation)
struct MsGeomPack { MsBranchGeom g[8]; };

DSG_HD void ms_window(int lo, int hi, int& lo8, int& nchw, int& Kp) {
    lo8 = lo & ~7;
    nchw = ((hi + 7) >> 3) - (lo >> 3);
    Kp = (nchw * 8 + 15) & ~15;
}

DSG_HD MsBranchGeom ms_geom(const dsg_ms_temporal_args& a, int j, int s) {
    MsBranchGeom g;
    g.w = a.br[j].hi - a.br[j].lo;
    ms_window(a.br[j].lo, a.br[j].hi, g.lo8, g.nchw, g.Kp);
    g.off = a.br[j].lo - g.lo8;
    g.nch = g.Kp >> 3;
    g.d = a.br[j].dilation;
    g.qmin = floordiv(-g.d, s);
    g.Fq = MS_TO + floordiv(g.d, s) - g.qmin;
    g.col = 0;
    for (int i = 0; i < j; ++i)
        if (a.br[i].kind == 0) { int l8, nw, kp; ms_window(a.br[i].lo, a.br[i].hi, l8, nw, kp); g.col += kp; }
    g.nsh = -1;
    for (int b = 0; b < 8; ++b) if ((1 << b) == g.nch) g.nsh = b;
    return g;
}
static MsGeomPack ms_geom_pack(const dsg_ms_temporal_args& a) {
    MsGeomPack p{};
    for (int j = 0; j < a.n_branches && j < 8; ++j)
        if (a.br[j].kind == 0) p.g[j] = ms_geom(a, j, a.stride);
    return p;
}
// item -> (row, chunk) without an integer division when the chunk count is a power of two
DSG_D void ms_split(const MsBranchGeom& g, int it, int& row, int& kc) {
    if (g.nsh >= 0) { row = it >> g.nsh; kc = it & (g.nch - 1); }
    else { row = it / g.nch; kc = it - row * g.nch; }
}

// byte offset of branch j's packed weight tiles inside wpack: [orientation 0: n=co,k=ci | 1: n=ci,k=co][tap][Kp x Kp] bf16
DSG_HD long long ms_wpack_off(const dsg_ms_temporal_args& a, int j) {
    long long o = 0;
    for (int i = 0; i < j; ++i)
        if (a.br[i].kind == 0) { int l8, nw, kp; ms_window(a.br[i].lo, a.br[i].hi, l8, nw, kp); o += 6LL * kp * kp * 2; }
    return o;
}

// one CTA per (conv branch, orientation): zero-padded bf16 tiles in the K-major no-swizzle UMMA layout
__global__ void ms_wpack_kernel(dsg_ms_temporal_args a) {
    int j = -1, cnt = 0;
    for (int i = 0; i < a.n_branches; ++i)
        if (a.br[i].kind == 0) { if (cnt == (int)blockIdx.x) j = i; ++cnt; }
    if (j < 0) return;
    const int orient = blockIdx.y;
    int lo8, nchw, Kp;
    ms_window(a.br[j].lo, a.br[j].hi, lo8, nchw, Kp);
    const int w = a.br[j].hi - a.br[j].lo, off = a.br[j].lo - lo8, nch = Kp >> 3;
    unsigned char* dst = reinterpret_cast<unsigned char*>(a.wpack) + ms_wpack_off(a, j) + (long long)orient * 3 * Kp * Kp * 2;
    for (int idx = threadIdx.x; idx < 3 * Kp * nch; idx += blockDim.x) {       // one 16-byte chunk (8 k) per item
        const int kc = idx % nch, n = (idx / nch) % Kp, dt = idx / (nch * Kp);
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int k = kc * 8 + e;
            const int co = (orient == 0 ? n : k) - off, ci = (orient == 0 ? k : n) - off;
            v[e] = (co >= 0 && co < w && ci >= 0 && ci < w) ? a.br[j].W[((long long)co * w + ci) * 3 + dt] : 0.f;
        }
        *reinterpret_cast<uint4*>(dst + dt * Kp * Kp * 2 + op_off(n, kc, nch)) = pack8(v);
    }
}

// copy branch j's three packed tiles (orientation o) into shared memory
DSG_D void ms_load_w(const dsg_ms_temporal_args& a, int j, int orient, int Kp, unsigned char* Wt) {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned char*>(a.wpack) + ms_wpack_off(a, j) + (long long)orient * 3 * Kp * Kp * 2);
    uint4* dst = reinterpret_cast<uint4*>(Wt);
    for (int i = threadIdx.x; i < 3 * Kp * Kp * 2 / 16; i += (int)blockDim.x) dst[i] = src[i];
}

constexpr int MS_CMAX = 512;        // channels the staged coefficient arrays hold

// BN coefficients of the branch pre-activations, staged once per CTA
DSG_D void ms_stage_b(const dsg_ms_temporal_args& a, float* cfa, float* cfb) {
    for (int c = threadIdx.x; c < a.C; c += (int)blockDim.x) { cfa[c] = a.b.a1[c]; cfb[c] = a.b.b1[c]; }
}
// coefficients of the dfeat source (a1, b1+b2, a2)
DSG_D void ms_stage_d(const dsg_ms_temporal_args& a, float* d1, float* db, float* d2) {
    for (int c = threadIdx.x; c < a.C; c += (int)blockDim.x) {
        d1[c] = a.dfeat.a1 ? a.dfeat.a1[c] : 1.f;
        db[c] = (a.dfeat.b1 ? a.dfeat.b1[c] : 0.f) + (a.dfeat.b2 ? a.dfeat.b2[c] : 0.f);
        d2[c] = a.dfeat.a2 ? a.dfeat.a2[c] : 1.f;
    }
}
// 8 channels of dfeat at row r with staged coefficients
DSG_D void ms_dfeat8(const dsg_ms_temporal_args& a, long long r, int c8, const float* d1, const float* db, const float* d2, float* v) {
    float x[8], k1[8], kb[8], k2[8];
    const uint4 r1 = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.dfeat.x1) + r * a.dfeat.ld1 + c8);
    uint4 r2 = make_uint4(0u, 0u, 0u, 0u);
    if (a.dfeat.x2) r2 = *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(a.dfeat.x2) + r * a.dfeat.ld2 + c8);
    load8f(d1 + c8, k1, 1.f);                                  // 16-byte shared loads (c8 % 8 == 0, arrays 16-byte aligned)
    load8f(db + c8, kb, 0.f);
    unpack8(r1, x);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], k1[e], kb[e]);
    if (a.dfeat.x2) {
        load8f(d2 + c8, k2, 1.f);
        unpack8(r2, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], k2[e], v[e]);
    }
}

// ---- batched staging: 4 items per thread in flight (loads first, then prologue + 16-byte shared store) -------------
// H tile: relu(bn(B)) of input frames t = s*(tq0 + qi) + p, for planes p in [0,s) and qi in [0,Fq); 32 rows per frame.
DSG_D void ms_stage_H(const dsg_ms_temporal_args& a, const MsBranchGeom& g, int n, int tq0, int s, int Vp, unsigned char* Ht,
                      const float* cfa, const float* cfb) {
    const bf16* Bx = reinterpret_cast<const bf16*>(a.b.x1);
    const int total = s * g.Fq * 32 * g.nch;
    for (int it0 = threadIdx.x; it0 < total; it0 += (int)blockDim.x * 4) {
        uint4 raw[4];
        int off[4], c8s[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int it = it0 + b * (int)blockDim.x;
            off[b] = -1; c8s[b] = -1;
            if (it < total) {
                int kc, row;
                ms_split(g, it, row, kc);
                const int v = row & 31, fi = row >> 5;
                off[b] = (int)op_off(row, kc, g.nch);
                if (v < Vp && kc < g.nchw) {
                    int p = 0, qi = fi;
                    while (qi >= g.Fq) { qi -= g.Fq; ++p; }
                    const int t = s * (tq0 + qi) + p;
                    if (t >= 0 && t < a.T_in) {
                        c8s[b] = g.lo8 + kc * 8;
                        raw[b] = *reinterpret_cast<const uint4*>(Bx + (((long long)n * a.T_in + t) * Vp + v) * a.b.ld1 + c8s[b]);
                    }
                }
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (off[b] < 0) continue;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (c8s[b] >= 0) {
                float x[8], ka[8], kb[8];
                unpack8(raw[b], x);
                load8f(cfa + c8s[b], ka, 1.f);
                load8f(cfb + c8s[b], kb, 0.f);
#pragma unroll
                for (int e = 0; e < 8; ++e) x[e] = fmaxf(fmaf(x[e], ka[e], kb[e]), 0.f);
                pk = pack8(x);
            }
            *reinterpret_cast<uint4*>(Ht + off[b]) = pk;
        }
    }
}
// dO tile: dfeat of output frames tpo0 + qi (qi in [0,nfr)), joint rows only (padding rows and the joint-mean row zero)
DSG_D void ms_stage_dO(const dsg_ms_temporal_args& a, const MsBranchGeom& g, int n, int tpo0, int nfr, int V, unsigned char* Dt,
                       const float* d1, const float* db, const float* d2) {
    const bf16* X1 = reinterpret_cast<const bf16*>(a.dfeat.x1);
    const bf16* X2 = reinterpret_cast<const bf16*>(a.dfeat.x2);
    const int total = nfr * 32 * g.nch;
    for (int it0 = threadIdx.x; it0 < total; it0 += (int)blockDim.x * 4) {
        uint4 r1[4], r2[4];
        int off[4], c8s[4];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int it = it0 + b * (int)blockDim.x;
            off[b] = -1; c8s[b] = -1;
            if (it < total) {
                int kc, row;
                ms_split(g, it, row, kc);
                const int v = row & 31, qi = row >> 5;
                off[b] = (int)op_off(row, kc, g.nch);
                const int tpo = tpo0 + qi;
                if (v < V && kc < g.nchw && tpo >= 0 && tpo < a.T_out) {
                    c8s[b] = g.lo8 + kc * 8;
                    const long long r = ((long long)n * a.T_out + tpo) * V + v;
                    r1[b] = *reinterpret_cast<const uint4*>(X1 + r * a.dfeat.ld1 + c8s[b]);
                    if (X2) r2[b] = *reinterpret_cast<const uint4*>(X2 + r * a.dfeat.ld2 + c8s[b]);
                }
            }
        }
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            if (off[b] < 0) continue;
            uint4 pk = make_uint4(0u, 0u, 0u, 0u);
            if (c8s[b] >= 0) {
                float x[8], v[8], k1[8], kb[8];
                unpack8(r1[b], x);
                load8f(d1 + c8s[b], k1, 1.f);
                load8f(db + c8s[b], kb, 0.f);
#pragma unroll
                for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], k1[e], kb[e]);
                if (X2) {
                    float k2[8];
                    unpack8(r2[b], x);
                    load8f(d2 + c8s[b], k2, 1.f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) v[e] = fmaf(x[e], k2[e], v[e]);
                }
                pk = pack8(v);
            }
            *reinterpret_cast<uint4*>(Dt + off[b]) = pk;
        }
    }
}

// NT = 256 (up to 4 CTAs per SM) or 512 (wide layers whose tiles leave room for one CTA per SM only: twice the warps)
template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 4 : 1) ms_temporal_fwd_kernel(dsg_ms_temporal_args a, MsGeomPack gp, int h_bytes, int w_bytes, int tmem_cols) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = a.V, Vp = a.V + a.has_ext, s = a.stride, C = a.C;
    const int n = blockIdx.y, tp0 = blockIdx.x * MS_TO;
    unsigned char* Ht = smem;                                // staged branch tile (K-major no-swizzle, 32 rows / frame)
    unsigned char* Wt = smem + h_bytes;                      // 3 taps x [Kp x Kp]
    bf16* feat_s = reinterpret_cast<bf16*>(smem + h_bytes + w_bytes);      // [MS_TO*V][C]
    const bf16* Bx = reinterpret_cast<const bf16*>(a.b.x1);
    __shared__ __align__(16) float cfa[MS_CMAX], cfb[MS_CMAX], bias_s[MS_CMAX], addc_s[32];
    const int CS = C + 8;                                    // feat_s row pitch: 16-byte aligned rows, bank-staggered
    ms_stage_b(a, cfa, cfb);
    for (int j = 0; j < a.n_branches; ++j)
        if (a.br[j].kind == 0)
            for (int k = tid; k < a.br[j].hi - a.br[j].lo; k += NT) bias_s[a.br[j].lo + k] = a.br[j].bias[k];
    if (tid < 32) addc_s[tid] = (a.has_ext && tid < V) ? a.add_coeff[tid] : 0.f;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    uint32_t phase = 0;
    int issued = 0;
    for (int j = 0; j < a.n_branches; ++j) {
        if (a.br[j].kind != 0) continue;
        const MsBranchGeom& g = gp.g[j];
        if (issued) mbar_wait(&mbar, phase ^ 1);             // the previous branch's MMAs are done with Ht / Wt
        // ---- stage relu(bn(B)) over the branch's channel window (with the temporal halo)
        ms_stage_H(a, g, n, tp0 + g.qmin, s, Vp, Ht, cfa, cfb);
        ms_load_w(a, j, 0, g.Kp, Wt);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t idesc = make_idesc(128, g.Kp);
            const uint32_t sbo = (uint32_t)g.nch * 128u;
            const uint32_t h0 = smem_u32(Ht), w0 = smem_u32(Wt);
            for (int dt = 0; dt < 3; ++dt) {
                const int o = (dt - 1) * g.d;
                const int p = posmod(o, s);
                const int qoff = (o - p) / s - g.qmin;                       // >= 0
                const uint32_t abase = h0 + (uint32_t)((p * g.Fq + qoff) * 4) * sbo;   // 4 row groups per frame
                const uint32_t bbase = w0 + (uint32_t)(dt * g.Kp * g.Kp * 2);
                for (int ks = 0; ks < (g.Kp >> 4); ++ks)
                    umma_f16(tmem_d + (uint32_t)g.col, make_desc(abase + ks * 256u, 128u, sbo), make_desc(bbase + ks * 256u, 128u, sbo), idesc,
                             (dt == 0 && ks == 0) ? 0u : 1u);
            }
            umma_commit(&mbar);
        }
        issued = 1;
        phase ^= 1;
    }
    // ---- max-pool / pass-through branches on the CUDA cores while the last MMAs drain: warp = (frame, half of the joints),
    //      lanes walk the channels of the range; the joint-mean column is evaluated first and stays in a register
    {
        constexpr int NP = NT / 128;                         // warps per frame
        const int fl = warp & 3, half = warp >> 2;           // frame, share of the joints
        const int tp = tp0 + fl;
        const int v0 = half * V / NP, v1 = (half + 1) * V / NP;
        for (int j = 0; j < a.n_branches; ++j) {
            const int kind = a.br[j].kind;
            if (kind == 0 || tp >= a.T_out) continue;
            const int lo = a.br[j].lo, w = a.br[j].hi - lo;
            const int tc = tp * s;                                           // centre input frame
            const bool has_m = tc - 1 >= 0, has_p = tc + 1 < a.T_in;
            for (int c0 = 0; c0 < w; c0 += 32) {
                if (c0 + lane >= w) continue;
                const int c = lo + c0 + lane;
                const float ca = cfa[c], cb = cfb[c];
                const bf16* col = Bx + ((long long)n * a.T_in + tc) * Vp * a.b.ld1 + c;     // (frame tc, column 0, channel c)
                const long long fstep = (long long)Vp * a.b.ld1;
                float glob = 0.f;
                for (int vv = a.has_ext ? -1 : v0; vv < v1; vv = (vv < 0 ? v0 : vv + 1)) {
                    const int cv = vv < 0 ? V : vv;                          // -1: the joint-mean column
                    const bf16* q = col + (long long)cv * a.b.ld1;
                    float val = fmaf(__bfloat162float(q[0]), ca, cb);
                    if (kind == 1) {
                        val = fmaxf(val, 0.f);
                        if (has_m) val = fmaxf(val, fmaf(__bfloat162float(q[-fstep]), ca, cb));
                        if (has_p) val = fmaxf(val, fmaf(__bfloat162float(q[fstep]), ca, cb));
                    }
                    if (vv < 0) {
                        glob = val;
                        if (half == 0) a.oglob[((long long)n * a.T_out + tp) * C + c] = val;
                        continue;
                    }
                    if (a.has_ext) val = fmaf(glob, addc_s[vv], val);
                    feat_s[(fl * V + vv) * CS + c] = __float2bfloat16(val);
                }
            }
        }
    }
    if (issued) {
        mbar_wait(&mbar, phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        // ---- conv branches: TMEM -> registers; warp w reads frame (w & 3); lanes are joints (lane V = joint-mean column)
        const int fl = warp & 3, half = warp >> 2;
        const int tp = tp0 + fl;
        const float addv = (a.has_ext && lane < V) ? addc_s[lane] : 0.f;
        const bool row_live = lane < V && tp < a.T_out, glob_live = lane == V && tp < a.T_out;
        bf16* frow = feat_s + (fl * V + lane) * CS;
        float* grow = a.oglob + ((long long)n * a.T_out + tp) * C;
        int gcount = 0;
        for (int j = 0; j < a.n_branches; ++j) {
            if (a.br[j].kind != 0) continue;
            const MsBranchGeom& g = gp.g[j];
            const int blo = a.br[j].lo, bhi = a.br[j].hi;
            for (int c16 = 0; c16 < g.Kp; c16 += 16, ++gcount) {
                if ((gcount % (NT / 128)) != half) continue;
                float v[16];
                tmem_ld16(tmem_d + ((uint32_t)(fl * 32) << 16) + (uint32_t)(g.col + c16), v);
                const int ch0 = g.lo8 + c16;                                 // absolute channel of window column c16
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    const int ch = ch0 + e;
                    if (ch < blo || ch >= bhi) continue;                     // warp-uniform
                    float val = v[e] + bias_s[ch];
                    if (a.has_ext) {
                        const float glob = __shfl_sync(0xffffffffu, val, V);
                        if (glob_live) grow[ch] = val;
                        val = fmaf(glob, addv, val);
                    }
                    if (row_live) frow[ch] = __float2bfloat16(val);
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
    // ---- write-out: complete rows, 16-byte stores, BatchNorm statistics of transform.0
    {
        const int nchunks = C >> 3, lanes = NT / nchunks;
        const int cc = tid % nchunks, rl = tid / nchunks;
        float s1[8], s2[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
        bf16* feat = reinterpret_cast<bf16*>(a.feat);
        for (int r = rl; r < MS_TO * V; r += lanes) {
            const int fl = r / V, vv = r - fl * V;
            const int tp = tp0 + fl;
            if (tp >= a.T_out) break;
            const uint4 u = *reinterpret_cast<const uint4*>(feat_s + r * CS + cc * 8);
            *reinterpret_cast<uint4*>(feat + (((long long)n * a.T_out + tp) * V + vv) * a.ld_feat + cc * 8) = u;
            if (a.stat_sum) {
                float x[8];
                unpack8(u, x);
#pragma unroll
                for (int e = 0; e < 8; ++e) { s1[e] += x[e]; s2[e] += x[e] * x[e]; }
            }
        }
        if (a.stat_sum) {
            float* red = reinterpret_cast<float*>(smem);     // [2][lanes][C]  (operand tiles are dead)
#pragma unroll
            for (int e = 0; e < 8; ++e) { red[(0 * lanes + rl) * C + cc * 8 + e] = s1[e]; red[(1 * lanes + rl) * C + cc * 8 + e] = s2[e]; }
            __syncthreads();
            for (int c = tid; c < C; c += NT) {
                float t1 = 0.f, t2 = 0.f;
                for (int l = 0; l < lanes; ++l) { t1 += red[(0 * lanes + l) * C + c]; t2 += red[(1 * lanes + l) * C + c]; }
                atomicAdd(a.stat_sum + c, (double)t1);
                atomicAdd(a.stat_sq + c, (double)t2);
            }
        }
    }
}

struct MsHostGeom { int h_bytes, w_bytes, tmem_cols, feat_bytes; bool ok; };

static MsHostGeom ms_host_geom(const dsg_ms_temporal_args& a, int out_rows_per_frame) {
    MsHostGeom h{0, 0, 0, 0, true};
    int Vp = a.V + a.has_ext, s = a.stride, cols = 0;
    if (a.C > MS_CMAX || Vp > 32 || a.n_branches > 8 || a.n_branches < 1 || s < 1 || a.C % 8 != 0 || (MS_THREADS % (a.C / 8)) != 0 || a.C / 8 > MS_THREADS) h.ok = false;
    for (int j = 0; j < a.n_branches && h.ok; ++j) {
        if (a.br[j].kind != 0) continue;
        int w = a.br[j].hi - a.br[j].lo, d = a.br[j].dilation, lo8, nchw, Kp;
        ms_window(a.br[j].lo, a.br[j].hi, lo8, nchw, Kp);
        if (w < 1 || Kp > 128 || d < 1 || !a.br[j].W || !a.br[j].bias) { h.ok = false; break; }
        int qmin = -((d + s - 1) / s), qmax = d / s;
        int Fq = MS_TO + qmax - qmin;
        int hb = s * Fq * 32 * Kp * 2, wb = 3 * Kp * Kp * 2;
        if (hb > h.h_bytes) h.h_bytes = hb;
        if (wb > h.w_bytes) h.w_bytes = wb;
        cols += Kp;
    }
    if (cols > 512) h.ok = false;
    h.tmem_cols = 32;
    while (h.tmem_cols < cols) h.tmem_cols <<= 1;
    h.feat_bytes = MS_TO * out_rows_per_frame * a.C * 2;
    // the statistics scratch re-uses the operand region: [2][lanes][C] floats
    int red = 2 * (512 / (a.C / 8 > 0 ? a.C / 8 : 1)) * a.C * 4;      // sized for the 512-thread variant
    if (h.h_bytes + h.w_bytes < red) h.h_bytes = red - h.w_bytes > 0 ? red - h.w_bytes : h.h_bytes;
    h.h_bytes = (h.h_bytes + 127) & ~127;
    h.w_bytes = (h.w_bytes + 127) & ~127;
    if ((size_t)h.h_bytes + h.w_bytes + h.feat_bytes > 200 * 1024) h.ok = false;
    return h;
}

static long long ms_wpack_bytes(const dsg_ms_temporal_args& a) { return ms_wpack_off(a, a.n_branches); }

static const char* ms_launch_wpack(const dsg_ms_temporal_args& a, dsg_stream_t st) {
    int nconv = 0;
    for (int j = 0; j < a.n_branches; ++j) nconv += a.br[j].kind == 0;
    if (nconv == 0) return nullptr;
    if (!a.wpack || (uintptr_t)a.wpack % 16 != 0) return "ms_temporal: wpack workspace missing or misaligned";
    ms_wpack_kernel<<<dim3(nconv, 2), dim3(256), 0, st>>>(a);
    return dsg_launch_error();
}

static bool ms_args_ok(const dsg_ms_temporal_args& a) {
    if (!a.b.x1 || a.b.x2 || !a.b.a1 || !a.b.b1 || a.b.a2 || a.b.b2) return false;
    if ((uintptr_t)a.b.x1 % 16 != 0 || a.b.ld1 % 8 != 0) return false;
    return true;
}

static const char* launch_ms_temporal_fwd(const dsg_ms_temporal_args& a, dsg_stream_t st) {
    MsHostGeom h = ms_host_geom(a, a.V);
    if (!h.ok || !ms_args_ok(a)) return "ms_temporal_fwd: unsupported shape (use the per-branch path)";
    if ((uintptr_t)a.feat % 16 != 0 || a.ld_feat % 8 != 0) return "ms_temporal_fwd: feat must be 16-byte aligned";
    if (a.n_samples <= 0 || a.T_out <= 0) return nullptr;
    size_t smem = (size_t)h.h_bytes + h.w_bytes + (size_t)MS_TO * a.V * (a.C + 8) * 2;
    if (smem > 200 * 1024) return "ms_temporal_fwd: shared memory budget exceeded";
    if (const char* e = ms_launch_wpack(a, st)) return e;
    dim3 grid((a.T_out + MS_TO - 1) / MS_TO, a.n_samples);
    if (smem > 100 * 1024) {                               // one CTA per SM anyway: give it 16 warps
        cudaFuncSetAttribute(ms_temporal_fwd_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms_temporal_fwd_kernel<512><<<grid, dim3(512), smem, st>>>(a, ms_geom_pack(a), h.h_bytes, h.w_bytes, h.tmem_cols);
    } else {
        cudaFuncSetAttribute(ms_temporal_fwd_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms_temporal_fwd_kernel<256><<<grid, dim3(256), smem, st>>>(a, ms_geom_pack(a), h.h_bytes, h.w_bytes, h.tmem_cols);
    }
    return dsg_launch_error();
}

// ------------------------------------------------------------------------------------------------------------
// Backward (data).  One CTA = one sample, one temporal plane p_in (t = s*q + p_in) and MS_TO input frames of it.
// For a conv branch, tap dt (offset o = (dt-1)*d) reaches input frame t from output frame t' = (t - o)/s when
// s | (p_in - o); for a fixed tap those t' are a contiguous run, so the staged dO tile (output-frame rows, 32 per
// frame, joint-mean row = sum_v dfeat[v]*add_coeff[v]) is again addressed by shifted descriptors and the B operand
// is W[:, :, dt]^T.  dH lands in TMEM; max-pool / pass-through gradients are routed on the CUDA cores; the write-out
// pass applies the ReLU mask from the stored pre-activations and accumulates the BatchNorm-backward sums.
struct MsBwdTaps { int n, sh[3], dt[3], shmin, Fq; };

DSG_D MsBwdTaps ms_bwd_taps(int d, int s, int p_in) {
    MsBwdTaps r;
    r.n = 0; r.shmin = 1 << 20;
    int shmax = -(1 << 20);
    for (int dt = 0; dt < 3; ++dt) {
        const int o = (dt - 1) * d;
        if (posmod(p_in - o, s) != 0) continue;
        const int sh = (p_in - o) / s;
        r.sh[r.n] = sh; r.dt[r.n] = dt; ++r.n;
        if (sh < r.shmin) r.shmin = sh;
        if (sh > shmax) shmax = sh;
    }
    r.Fq = r.n ? MS_TO + shmax - r.shmin : 0;
    return r;
}

template <int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? 3 : 1) ms_temporal_bwd_data_kernel(dsg_ms_temporal_args a, MsGeomPack gp, int h_bytes, int w_bytes, int tmem_cols,
                                                                          int mp_lo, int mp_hi) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_dadd[32];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = a.V, Vp = a.V + a.has_ext, s = a.stride, C = a.C;
    const int n = blockIdx.y, p_in = blockIdx.z, q0 = blockIdx.x * MS_TO;
    unsigned char* Dt = smem;                                 // staged dO tile
    unsigned char* Wt = smem + h_bytes;
    const int CS = C + 8;                                     // E_s row pitch: 16-byte aligned rows, bank-staggered
    bf16* E_s = reinterpret_cast<bf16*>(smem + h_bytes + w_bytes);          // [MS_TO*Vp][CS]
    const int mpw = mp_hi - mp_lo;                            // channels of the max/pass ranges (contiguous span)
    float* dg_s = reinterpret_cast<float*>(smem + h_bytes + w_bytes + MS_TO * Vp * CS * 2);   // [6][mpw]
    const bf16* Bx = reinterpret_cast<const bf16*>(a.b.x1);
    __shared__ __align__(16) float cfa[MS_CMAX], cfb[MS_CMAX], dc1[MS_CMAX], dcb[MS_CMAX], dc2[MS_CMAX];
    __shared__ unsigned char kind_s[MS_CMAX];                // branch kind per channel (0 conv, 1 max, 2 pass, 3 none)
    __shared__ MsBwdTaps taps_s[8];                          // per conv branch: the taps that reach this plane (once per CTA)
    ms_stage_b(a, cfa, cfb);
    ms_stage_d(a, dc1, dcb, dc2);
    for (int j = 0; j < a.n_branches; ++j)
        for (int c = a.br[j].lo + tid; c < a.br[j].hi; c += NT) kind_s[c] = (unsigned char)a.br[j].kind;
    if (tid < a.n_branches && a.br[tid].kind == 0) taps_s[tid] = ms_bwd_taps(gp.g[tid].d, s, p_in);

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols);
    if (tid < 32) s_dadd[tid] = 0.f;
    for (int i = tid * 8; i < MS_TO * Vp * CS; i += NT * 8) *reinterpret_cast<uint4*>(E_s + i) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    uint32_t phase = 0;
    int issued = 0;
    for (int j = 0; j < a.n_branches; ++j) {
        if (a.br[j].kind != 0) continue;
        const MsBranchGeom& g = gp.g[j];
        const MsBwdTaps& tp = taps_s[j];
        if (tp.n == 0) continue;
        if (issued) mbar_wait(&mbar, phase ^ 1);
        // ---- stage dO over the channel window (joint rows; padding rows and the joint-mean row start as zeros)
        ms_stage_dO(a, g, n, q0 + tp.shmin, tp.Fq, V, Dt, dc1, dcb, dc2);
        ms_load_w(a, j, 1, g.Kp, Wt);                                   // B operand [n = ci][k = co]
        if (a.has_ext) {
            __syncthreads();
            for (int it = tid; it < tp.Fq * g.nchw; it += NT) {   // joint-mean row: sum_v dO[v] * add_coeff[v], 8 channels per item
                const int kc = it % g.nchw, qi = it / g.nchw;
                float sacc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) sacc[e] = 0.f;
                for (int v = 0; v < V; ++v) {
                    float d[8];
                    unpack8(*reinterpret_cast<const uint4*>(Dt + op_off(qi * 32 + v, kc, g.nch)), d);
                    const float wv = a.add_coeff[v];
#pragma unroll
                    for (int e = 0; e < 8; ++e) sacc[e] = fmaf(d[e], wv, sacc[e]);
                }
                *reinterpret_cast<uint4*>(Dt + op_off(qi * 32 + V, kc, g.nch)) = pack8(sacc);
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (tid == 0) {
            const uint32_t idesc = make_idesc(128, g.Kp);
            const uint32_t sbo = (uint32_t)g.nch * 128u;
            const uint32_t h0 = smem_u32(Dt), w0 = smem_u32(Wt);
            for (int ti = 0; ti < tp.n; ++ti) {
                const uint32_t abase = h0 + (uint32_t)((tp.sh[ti] - tp.shmin) * 4) * sbo;
                const uint32_t bbase = w0 + (uint32_t)(tp.dt[ti] * g.Kp * g.Kp * 2);
                for (int ks = 0; ks < (g.Kp >> 4); ++ks)
                    umma_f16(tmem_d + (uint32_t)g.col, make_desc(abase + ks * 256u, 128u, sbo), make_desc(bbase + ks * 256u, 128u, sbo), idesc,
                             (ti == 0 && ks == 0) ? 0u : 1u);
            }
            umma_commit(&mbar);
        }
        issued |= (1 << j);
        phase ^= 1;
    }
    // ---- max-pool / pass-through gradients on the CUDA cores, 8 channels (one 16-byte chunk) per thread-item
    if (mpw > 0) {
        const int t_first = s * q0 + p_in;
        const int tp_lo = floordiv(t_first - 1, s);
        const int ac0 = mp_lo >> 3, nac = ((mp_hi + 7) >> 3) - ac0;
        if (a.has_ext) {
            for (int it = tid; it < 6 * nac; it += NT) {         // dO of the joint-mean column for the frames in reach
                const int ac = it % nac, fi = it / nac;
                const int tpo = tp_lo + fi, c8 = (ac0 + ac) * 8;
                float sacc[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) sacc[e] = 0.f;
                if (tpo >= 0 && tpo < a.T_out)
                    for (int v = 0; v < V; ++v) {
                        float d[8];
                        ms_dfeat8(a, ((long long)n * a.T_out + tpo) * V + v, c8, dc1, dcb, dc2, d);
                        const float w = a.add_coeff[v];
#pragma unroll
                        for (int e = 0; e < 8; ++e) sacc[e] = fmaf(d[e], w, sacc[e]);
                    }
#pragma unroll
                for (int e = 0; e < 8; ++e)
                    if (c8 + e >= mp_lo && c8 + e < mp_hi) dg_s[fi * mpw + c8 + e - mp_lo] = sacc[e];
            }
            __syncthreads();
        }
        for (int it = tid; it < MS_TO * Vp * nac; it += NT) {
            const int ac = it % nac, rest = it / nac;
            const int vv = rest % Vp, i = rest / Vp;
            const int t = s * (q0 + i) + p_in, c8 = (ac0 + ac) * 8;
            if (t >= a.T_in) continue;
            int kind[8];
            bool any_max = false;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                kind[e] = kind_s[c8 + e];
                any_max |= kind[e] == 1;
            }
            const float* ca = cfa + c8;
            const float* cb = cfb + c8;
            float h[5][8];                                               // relu(bn(B)) at frames t-2 .. t+2 (-1: out of range)
#pragma unroll
            for (int dd = 0; dd < 5; ++dd) {
                const int t2 = t + dd - 2;
                if (any_max && t2 >= 0 && t2 < a.T_in) {
                    float x[8];
                    unpack8(*reinterpret_cast<const uint4*>(Bx + (((long long)n * a.T_in + t2) * Vp + vv) * a.b.ld1 + c8), x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[dd][e] = fmaxf(fmaf(x[e], ca[e], cb[e]), 0.f);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) h[dd][e] = -1.f;
                }
            }
            float eout[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) eout[e] = 0.f;
#pragma unroll
            for (int dt = -1; dt <= 1; ++dt) {                           // windows t' with s*t' + dt == t
                const int num = t - dt;
                if (num < 0) continue;
                int tpo = num;
                if (s != 1) {                                            // stride 1 (eight layers of ten) divides nothing
                    if (num % s != 0) continue;
                    tpo = num / s;
                }
                if (tpo >= a.T_out) continue;
                float d[8];
                if (vv < V) ms_dfeat8(a, ((long long)n * a.T_out + tpo) * V + vv, c8, dc1, dcb, dc2, d);
                else {
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[e] = (c8 + e >= mp_lo && c8 + e < mp_hi) ? dg_s[(tpo - tp_lo) * mpw + c8 + e - mp_lo] : 0.f;
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (kind[e] == 2) { if (dt == 0) eout[e] = d[e]; continue; }
                    if (kind[e] != 1) continue;
                    // window tpo covers frames t-dt-1 .. t-dt+1 = h[1-dt .. 3-dt]; first maximum wins (ATen max_pool2d)
                    float m = -3.0e38f;
                    int am = -2;
#pragma unroll
                    for (int d2 = -1; d2 <= 1; ++d2) {
                        const float hv = h[2 - dt + d2][e];
                        if (hv >= 0.f && hv > m) { m = hv; am = d2; }
                    }
                    if (am == dt && h[2][e] > 0.f) eout[e] += d[e];
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (kind[e] == 1 || kind[e] == 2) E_s[(i * Vp + vv) * CS + c8 + e] = __float2bfloat16(eout[e]);
        }
    }
    // ---- dadd_coeff for the output frames this CTA owns (plane 0: t' = q0 + i): thread = (8-channel chunk, joint)
    if (a.has_ext && p_in == 0) {
        const int nchunks = C >> 3;
        const bool pow2 = (nchunks & (nchunks - 1)) == 0 && nchunks <= 32;      // a joint's chunks = one aligned lane group
        const int total = nchunks * V;
        for (int it0 = 0; it0 < total; it0 += NT) {                     // every lane iterates: warp shuffles below
            const int it = it0 + tid;
            const bool valid = it < total;
            const int cc = valid ? it % nchunks : 0, v = valid ? it / nchunks : 0;
            float part = 0.f;
            if (valid) {
                for (int i = 0; i < MS_TO; ++i) {
                    const int tpo = q0 + i;
                    if (tpo >= a.T_out) break;
                    float d[8], og[8];
                    ms_dfeat8(a, ((long long)n * a.T_out + tpo) * V + v, cc * 8, dc1, dcb, dc2, d);
                    load8f(a.oglob + ((long long)n * a.T_out + tpo) * C + cc * 8, og, 0.f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) part = fmaf(d[e], og[e], part);
                }
            }
            if (pow2) {
                for (int o = nchunks >> 1; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
                if (valid && cc == 0) s_dadd[v] += part;                        // one writer per joint
            } else if (valid) atomicAdd(&s_dadd[v], part);
        }
    }
    if (issued) {
        mbar_wait(&mbar, phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int i = warp & 3, half = warp >> 2;
        int gcount = 0;
        for (int j = 0; j < a.n_branches; ++j) {
            if (!(issued & (1 << j))) continue;
            const MsBranchGeom& g = gp.g[j];
            for (int c16 = 0; c16 < g.Kp; c16 += 16, ++gcount) {
                if ((gcount % (NT / 128)) != half) continue;
                float v[16];
                tmem_ld16(tmem_d + ((uint32_t)(i * 32) << 16) + (uint32_t)(g.col + c16), v);
                if (lane < Vp) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int k = c16 + e - g.off;
                        if (k >= 0 && k < g.w) E_s[(i * Vp + lane) * CS + a.br[j].lo + k] = __float2bfloat16(v[e]);
                    }
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
    if (a.has_ext && p_in == 0 && tid < V) atomicAdd(a.dadd_coeff + tid, s_dadd[tid]);
    // ---- write-out: ReLU mask from the stored pre-activations (not on the pass range), BN-backward sums, 16-byte stores
    {
        const int nchunks = C >> 3, lanes = NT / nchunks;
        const int cc = tid % nchunks, rl = tid / nchunks;
        const int c0 = cc * 8;
        float s1[8], s2[8], msk[8];
        const float* ca = cfa + c0;
        const float* cb = cfb + c0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            s1[e] = s2[e] = 0.f;
            msk[e] = kind_s[c0 + e] == 2 ? 0.f : 1.f;            // pass range: no ReLU
        }
        bf16* E = reinterpret_cast<bf16*>(a.e);
        for (int r = rl; r < MS_TO * Vp; r += lanes) {
            const int i = r / Vp, vv = r - i * Vp;
            const int t = s * (q0 + i) + p_in;
            if (t >= a.T_in) break;
            const long long grow = ((long long)n * a.T_in + t) * Vp + vv;
            float x[8], braw[8];
            unpack8(*reinterpret_cast<const uint4*>(E_s + r * CS + c0), x);
            unpack8(*reinterpret_cast<const uint4*>(Bx + grow * a.b.ld1 + c0), braw);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (msk[e] != 0.f && !(fmaf(braw[e], ca[e], cb[e]) > 0.f)) x[e] = 0.f;
                s1[e] += x[e];
                s2[e] += x[e] * braw[e];
            }
            *reinterpret_cast<uint4*>(E + grow * a.ld_e + c0) = pack8(x);
        }
        if (a.e_sum) {
            float* red = reinterpret_cast<float*>(smem);
#pragma unroll
            for (int e = 0; e < 8; ++e) { red[(0 * lanes + rl) * C + c0 + e] = s1[e]; red[(1 * lanes + rl) * C + c0 + e] = s2[e]; }
            __syncthreads();
            for (int c = tid; c < C; c += NT) {
                float t1 = 0.f, t2 = 0.f;
                for (int l = 0; l < lanes; ++l) { t1 += red[(0 * lanes + l) * C + c]; t2 += red[(1 * lanes + l) * C + c]; }
                atomicAdd(a.e_sum + c, (double)t1);
                atomicAdd(a.e_sq + c, (double)t2);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Backward (weights).  Persistent CTAs, tile-outer: a CTA walks its share of (sample, 4 output frames) tiles; for every
// tile it stages relu(bn(B)) with halo (as forward) and dO (4 frames) for ALL conv branches of the current group at
// once — every byte of B and dfeat is read once per tile, all the loads of a tile are in flight together, and there
// is one barrier / MMA round per tile — then accumulates dW_j[:, :, dt] (+)= dO_j^T * H_j,shift(dt) for every branch
// and tap into disjoint TMEM column ranges.  Both operands are read as MN-major (the reduction runs over rows): the
// staged tiles are byte-identical to the K-major ones, only the descriptor strides swap (SBO = 128 B between
// 8-channel groups, LBO = 128 B * chunks between 8-row groups).  Branch groups exist because TMEM has 512 columns
// (3 * Kp per branch, at most 256 per CTA so two CTAs share an SM) and shared memory is finite: 64-channel layers run
// one group, 128-channel layers two, 256-channel layers one branch per group.
struct MsBwPlan {
    int ngroups, gstart[9];            // group g = conv branches with ordinal in [gstart[g], gstart[g+1])
    int hoff[8], doff[8], col[8];      // per branch j: byte offsets of its H / dO tiles, first TMEM column
};

__global__ void __launch_bounds__(MS_THREADS) ms_temporal_bwd_weight_kernel(dsg_ms_temporal_args a, MsGeomPack gp, MsBwPlan pl, int tmem_cols) {
    DSG_DYN_SMEM(smem);
    __shared__ uint64_t mbar;
    __shared__ uint32_t tmem_base_s;
    __shared__ float s_db[8][128];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int V = a.V, Vp = a.V + a.has_ext, s = a.stride;
    __shared__ __align__(16) float cfa[MS_CMAX], cfb[MS_CMAX], dc1[MS_CMAX], dcb[MS_CMAX], dc2[MS_CMAX], addc_s[32];
    ms_stage_b(a, cfa, cfb);
    ms_stage_d(a, dc1, dcb, dc2);
    if (tid < 32) addc_s[tid] = (a.has_ext && tid < V) ? a.add_coeff[tid] : 0.f;
    const int chunks_t = (a.T_out + MS_TO - 1) / MS_TO;
    const int n_tiles = a.n_samples * chunks_t;

    if (tid == 0) {
        mbar_init(&mbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, (uint32_t)tmem_cols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_d = tmem_base_s;

    // conv branches in order
    int cj[8], ncj = 0;
    for (int j = 0; j < a.n_branches; ++j)
        if (a.br[j].kind == 0) cj[ncj++] = j;

    uint32_t phase = 0;
    int pending = 0;
    for (int grp = 0; grp < pl.ngroups; ++grp) {
        const int b0 = pl.gstart[grp], b1 = pl.gstart[grp + 1];
        for (int i = tid; i < 8 * 128; i += MS_THREADS) (&s_db[0][0])[i] = 0.f;
        // this thread's (branch, frame, chunk) item of the joint-mean / bias pass: fixed for the whole group
        int my_j = -1, my_qi = 0, my_kc = 0;
        {
            int base = 0;
            for (int bi = b0; bi < b1; ++bi) {
                const int j = cj[bi], cnt = MS_TO * gp.g[j].nchw;
                if (my_j < 0 && tid >= base && tid < base + cnt) { my_j = j; my_kc = (tid - base) % gp.g[j].nchw; my_qi = (tid - base) / gp.g[j].nchw; }
                base += cnt;
            }
        }
        float dbacc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) dbacc[e] = 0.f;
        int first = 1;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
            const int n = tile / chunks_t, tp0 = (tile - n * chunks_t) * MS_TO;
            if (pending) { mbar_wait(&mbar, phase ^ 1); pending = 0; }
            for (int bi = b0; bi < b1; ++bi) {
                const int j = cj[bi];
                const MsBranchGeom& g = gp.g[j];
                ms_stage_H(a, g, n, tp0 + g.qmin, s, Vp, smem + pl.hoff[j], cfa, cfb);
                ms_stage_dO(a, g, n, tp0, MS_TO, V, smem + pl.doff[j], dc1, dcb, dc2);
            }
            __syncthreads();
            if (my_j >= 0) {                                                  // joint-mean row of dO; bias gradient
                const MsBranchGeom& g = gp.g[my_j];
                unsigned char* Dt = smem + pl.doff[my_j];
                float sacc[8], tot[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) sacc[e] = tot[e] = 0.f;
                for (int v = 0; v < V; ++v) {
                    float d[8];
                    unpack8(*reinterpret_cast<const uint4*>(Dt + op_off(my_qi * 32 + v, my_kc, g.nch)), d);
                    const float wv = addc_s[v];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { sacc[e] = fmaf(d[e], wv, sacc[e]); tot[e] += d[e]; }
                }
                if (a.has_ext) *reinterpret_cast<uint4*>(Dt + op_off(my_qi * 32 + V, my_kc, g.nch)) = pack8(sacc);
#pragma unroll
                for (int e = 0; e < 8; ++e) dbacc[e] += tot[e] + (a.has_ext ? sacc[e] : 0.f);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncthreads();
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (tid == 0) {
                for (int bi = b0; bi < b1; ++bi) {
                    const int j = cj[bi];
                    const MsBranchGeom& g = gp.g[j];
                    const uint32_t idesc = make_idesc_mn(128, g.Kp);
                    const uint32_t lbo = (uint32_t)g.nch * 128u;                   // between 8-row groups
                    const uint32_t h0 = smem_u32(smem + pl.hoff[j]), d0 = smem_u32(smem + pl.doff[j]);
                    for (int dt = 0; dt < 3; ++dt) {
                        const int o = (dt - 1) * g.d;
                        const int p = posmod(o, s);
                        const int qoff = (o - p) / s - g.qmin;
                        const uint32_t hbase = h0 + (uint32_t)((p * g.Fq + qoff) * 4) * lbo;
                        for (int ks = 0; ks < 8; ++ks)                              // 128 rows = 8 x K16
                            umma_f16(tmem_d + (uint32_t)(pl.col[j] + dt * g.Kp), make_desc(d0 + ks * 2u * lbo, lbo, 128u),
                                     make_desc(hbase + ks * 2u * lbo, lbo, 128u), idesc, (first && ks == 0) ? 0u : 1u);
                    }
                }
                umma_commit(&mbar);
            }
            first = 0;
            pending = 1;
            phase ^= 1;
        }
        if (pending) { mbar_wait(&mbar, phase ^ 1); pending = 0; }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (my_j >= 0) {
#pragma unroll
            for (int e = 0; e < 8; ++e) atomicAdd(&s_db[my_j][my_kc * 8 + e], dbacc[e]);
        }
        __syncthreads();
        if (!first) {
            // D rows = output channel co (lanes), columns = (tap, ci)
            const int lq = warp & 3, half = warp >> 2;
            const int co = lq * 32 + lane;
            int gcount = 0;
            for (int bi = b0; bi < b1; ++bi) {
                const int j = cj[bi];
                const MsBranchGeom& g = gp.g[j];
                for (int dt = 0; dt < 3; ++dt)
                    for (int c16 = 0; c16 < g.Kp; c16 += 16, ++gcount) {
                        if ((gcount & 1) != half) continue;
                        if (lq * 32 >= g.Kp) continue;                              // warp-uniform: no live rows in this lane quarter
                        float v[16];
                        tmem_ld16(tmem_d + ((uint32_t)(lq * 32) << 16) + (uint32_t)(pl.col[j] + dt * g.Kp + c16), v);
                        const int cor = co - g.off;                               // window row -> output channel of the branch
                        if (cor >= 0 && cor < g.w) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int ci = c16 + e - g.off;
                                if (ci >= 0 && ci < g.w) atomicAdd(a.br[j].dW + ((long long)cor * g.w + ci) * 3 + dt, v[e]);
                            }
                        }
                    }
                if (tid < g.w && a.br[j].db) atomicAdd(a.br[j].db + tid, s_db[j][tid + g.off]);
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (warp == 0) tmem_dealloc(tmem_d, (uint32_t)tmem_cols);
}

static const char* launch_ms_temporal_bwd_data(const dsg_ms_temporal_args& a, dsg_stream_t st) {
    const int Vp = a.V + a.has_ext;
    MsHostGeom h = ms_host_geom(a, Vp);
    if (!h.ok || !ms_args_ok(a)) return "ms_temporal_bwd_data: unsupported shape (use the per-branch path)";
    if (!act8_ok(a.dfeat) || a.dfeat.relu) return "ms_temporal_bwd_data: dfeat must be 16-byte aligned";
    if ((uintptr_t)a.e % 16 != 0 || a.ld_e % 8 != 0) return "ms_temporal_bwd_data: e must be 16-byte aligned";
    if (a.n_samples <= 0 || a.T_in <= 0) return nullptr;
    int mp_lo = 1 << 30, mp_hi = 0;
    for (int j = 0; j < a.n_branches; ++j)
        if (a.br[j].kind != 0) { if (a.br[j].lo < mp_lo) mp_lo = a.br[j].lo; if (a.br[j].hi > mp_hi) mp_hi = a.br[j].hi; }
    if (mp_hi <= mp_lo) { mp_lo = 0; mp_hi = 0; }
    size_t smem = (size_t)h.h_bytes + h.w_bytes + (size_t)MS_TO * Vp * (a.C + 8) * 2 + (size_t)6 * (mp_hi - mp_lo) * 4 + 16;
    if (smem > 200 * 1024) return "ms_temporal_bwd_data: shared memory budget exceeded";
    if (const char* e = ms_launch_wpack(a, st)) return e;

    const int frames_per_plane = (a.T_in + a.stride - 1) / a.stride;
    dim3 grid((frames_per_plane + MS_TO - 1) / MS_TO, a.n_samples, a.stride);
    if (smem > 100 * 1024) {                               // one CTA per SM anyway: give it 16 warps
        cudaFuncSetAttribute(ms_temporal_bwd_data_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms_temporal_bwd_data_kernel<512><<<grid, dim3(512), smem, st>>>(a, ms_geom_pack(a), h.h_bytes, h.w_bytes, h.tmem_cols, mp_lo, mp_hi);
    } else {
        cudaFuncSetAttribute(ms_temporal_bwd_data_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        ms_temporal_bwd_data_kernel<256><<<grid, dim3(256), smem, st>>>(a, ms_geom_pack(a), h.h_bytes, h.w_bytes, h.tmem_cols, mp_lo, mp_hi);
    }
    return dsg_launch_error();
}

static const char* launch_ms_temporal_bwd_weight(const dsg_ms_temporal_args& a, dsg_stream_t st) {
    MsHostGeom h = ms_host_geom(a, a.V);
    if (!h.ok || !ms_args_ok(a)) return "ms_temporal_bwd_weight: unsupported shape (use the per-branch path)";
    if (!act8_ok(a.dfeat) || a.dfeat.relu) return "ms_temporal_bwd_weight: dfeat must be 16-byte aligned";
    if (a.n_samples <= 0 || a.T_out <= 0) return nullptr;
    const MsGeomPack gp = ms_geom_pack(a);
    int cj[8], ncj = 0;
    for (int j = 0; j < a.n_branches; ++j)
        if (a.br[j].kind == 0) {
            if (!a.br[j].dW) return "ms_temporal_bwd_weight: dW missing";
            cj[ncj++] = j;
        }
    if (ncj == 0) return nullptr;
    // greedy grouping: a group fits 256 TMEM columns (3 * Kp per branch; two CTAs per SM stay resident and other
    // tcgen05 kernels can still allocate), ~96 KB of operand tiles and 256 bias items
    MsBwPlan pl{};
    const int smem_budget = 96 * 1024;
    int maxcols = 0, maxbytes = 0, cols = 0, bytes = 0, items = 0;
    pl.ngroups = 0;
    pl.gstart[0] = 0;
    for (int bi = 0; bi < ncj; ++bi) {
        const int j = cj[bi];
        const MsBranchGeom& g = gp.g[j];
        const int hb = (a.stride * g.Fq * 32 * g.Kp * 2 + 127) & ~127;
        // the A operand is read as M = 128 channel rows: reserve 16 channel groups per 8-row group even when Kp < 128
        const int db = (MS_TO * 32 * g.Kp * 2 + 16 * 128 + 127) & ~127;
        const int it = MS_TO * g.nchw;
        if (3 * g.Kp > 512 || it > MS_THREADS) return "ms_temporal_bwd_weight: TMEM budget exceeded";
        if (bi > pl.gstart[pl.ngroups] && (cols + 3 * g.Kp > 256 || bytes + hb + db > smem_budget || items + it > MS_THREADS)) {
            pl.gstart[++pl.ngroups] = bi;
            cols = bytes = items = 0;
        }
        pl.col[j] = cols;
        pl.hoff[j] = bytes;
        pl.doff[j] = bytes + hb;
        cols += 3 * g.Kp;
        bytes += hb + db;
        items += it;
        if (cols > maxcols) maxcols = cols;
        if (bytes > maxbytes) maxbytes = bytes;
    }
    pl.gstart[++pl.ngroups] = ncj;
    int tcols = 32;
    while (tcols < maxcols) tcols <<= 1;
    size_t smem = (size_t)maxbytes + 2048;
    if (smem > 200 * 1024) return "ms_temporal_bwd_weight: shared memory budget exceeded";
    cudaFuncSetAttribute(ms_temporal_bwd_weight_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    int n_tiles = a.n_samples * ((a.T_out + MS_TO - 1) / MS_TO);
    int per_sm = (int)((200 * 1024) / (smem + 16 * 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 2) per_sm = 2;                          // register-limited
    if (per_sm * tcols > 512) per_sm = 512 / tcols;
    int grid = n_tiles < per_sm * dsg_num_sms() ? n_tiles : per_sm * dsg_num_sms();
    ms_temporal_bwd_weight_kernel<<<dim3(grid), dim3(MS_THREADS), smem, st>>>(a, gp, pl, tcols);
    return dsg_launch_error();
}

}  // namespace tc
}  // namespace dsg
#endif
