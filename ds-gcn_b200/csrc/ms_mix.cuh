// Streaming halves of the multi-scale temporal unit around dsg_ms_conv (tcn.py:383-396, 407-420): everything that is not a
// convolution.  16-byte (8-channel) vector kernels, thread = (frame, channel chunk), every row of a frame walked by the same
// thread so the joint-mean ("global") row is a register, several independent loads in flight per thread.
//   forward   : feat[n,t',v,:] = O_v + O_V * add_coeff[v] with O = conv outputs (read) | 3x1 max-pool of relu(bn(B)) | bn(B) at s*t';
//               oglob = O_V (fp32, saved for backward); BatchNorm statistics of transform.0
//   backward 1: dO[n,t',v,:] = dfeat_v, dO[n,t',V,:] = sum_v dfeat_v * add_coeff[v]; dadd_coeff[v] += sum dfeat_v . oglob
//   backward 2: e of the max-pool range (arg-max routing, first maximum wins as ATen max_pool2d, ReLU mask) and of the pass range,
//               BN-backward sums of the max range; read-modify-write of chunks shared with the conv range
#pragma once
#include "dsg_common.h"

namespace dsg {

constexpr int MX_THREADS = 256;
constexpr int MX_U = 5;              // rows in flight per thread

struct MxKinds { unsigned conv, mx, pass; };     // bit e: channel c8 + e is of that kind
DSG_D MxKinds mx_kinds(const dsg_ms_combine_args& a, int c8) {
    MxKinds k{0u, 0u, 0u};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        const int c = c8 + e;
        if (c >= a.conv_lo && c < a.conv_hi) k.conv |= 1u << e;
        else if (c >= a.max_lo && c < a.max_hi) k.mx |= 1u << e;
        else if (c >= a.pass_lo && c < a.pass_hi) k.pass |= 1u << e;
    }
    return k;
}
DSG_D void mx_coefs(const dsg_ms_combine_args& a, int c8, float* ka, float* kb) {
    load8f(a.b.a1 ? a.b.a1 + c8 : nullptr, ka, 1.f);
    load8f(a.b.b1 ? a.b.b1 + c8 : nullptr, kb, 0.f);
    if (a.b.b2) { float t[8]; load8f(a.b.b2 + c8, t, 0.f);
#pragma unroll
        for (int e = 0; e < 8; ++e) kb[e] += t[e]; }
}
DSG_D uint4 mx_ld(const void* base, long long row, long long ld, int c8) {
    return *reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + row * ld + c8);
}

struct MxRaw { uint4 o, bm, bc, bp; };
DSG_D MxRaw mx_issue(const dsg_ms_combine_args& a, const MxKinds& k, long long fo, long long fi, int j, int Vp, bool has_m, bool has_p, int c8) {
    MxRaw r;
    r.o = r.bm = r.bc = r.bp = make_uint4(0u, 0u, 0u, 0u);
    if (k.conv) r.o = mx_ld(a.o, fo * Vp + j, a.ld_o, c8);
    if (k.mx | k.pass) r.bc = mx_ld(a.b.x1, fi * Vp + j, a.b.ld1, c8);
    if (k.mx) {
        if (has_m) r.bm = mx_ld(a.b.x1, (fi - 1) * Vp + j, a.b.ld1, c8);
        if (has_p) r.bp = mx_ld(a.b.x1, (fi + 1) * Vp + j, a.b.ld1, c8);
    }
    return r;
}
DSG_D void mx_finish(const MxRaw& r, const MxKinds& k, bool has_m, bool has_p, const float* ka, const float* kb, float* out) {
    float o[8], hc[8];
    unpack8(r.o, o);
    unpack8(r.bc, hc);
#pragma unroll
    for (int e = 0; e < 8; ++e) hc[e] = fmaf(hc[e], ka[e], kb[e]);
#pragma unroll
    for (int e = 0; e < 8; ++e) out[e] = (k.conv >> e & 1u) ? o[e] : ((k.pass >> e & 1u) ? hc[e] : 0.f);
    if (k.mx) {
        float m[8], t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) m[e] = fmaxf(hc[e], 0.f);
        if (has_m) { unpack8(r.bm, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], fmaf(t[e], ka[e], kb[e])); }
        if (has_p) { unpack8(r.bp, t);
#pragma unroll
            for (int e = 0; e < 8; ++e) m[e] = fmaxf(m[e], fmaf(t[e], ka[e], kb[e])); }
#pragma unroll
        for (int e = 0; e < 8; ++e) if (k.mx >> e & 1u) out[e] = m[e];
    }
}

template <int U, int OCC>
__global__ void __launch_bounds__(MX_THREADS, OCC) ms_mix_fwd_kernel(dsg_ms_combine_args a, int frames_per_cta) {
    DSG_SHARED float s_red[2][2048];                     // [lanes][C] with lanes * C == 2048
    DSG_SHARED float addc_s[32];
    const int tid = threadIdx.x, nch = a.C >> 3, lanes = MX_THREADS / nch;
    const int cc = tid % nch, fl = tid / nch, c8 = cc * 8;
    const int V = a.V, Vp = a.V + a.has_ext, s = a.stride;
    if (tid < 32) addc_s[tid] = (a.has_ext && tid < V) ? a.add_coeff[tid] : 0.f;
    __syncthreads();
    const MxKinds k = mx_kinds(a, c8);
    float ka[8], kb[8], s1[8], s2[8];
    mx_coefs(a, c8, ka, kb);
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    long long fend = (long long)(blockIdx.x + 1) * frames_per_cta;
    if (fend > n_frames) fend = n_frames;
    bf16* feat = reinterpret_cast<bf16*>(a.feat);
    if (fl < lanes)
        for (long long f = (long long)blockIdx.x * frames_per_cta + fl; f < fend; f += lanes) {
            const int n = (int)(f / a.T_out), tp = (int)(f - (long long)n * a.T_out), tc = tp * s;
            const long long fi = (long long)n * a.T_in + tc;
            const bool has_m = tc - 1 >= 0, has_p = tc + 1 < a.T_in;
            float glob[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) glob[e] = 0.f;
            if (a.has_ext) {
                mx_finish(mx_issue(a, k, f, fi, V, Vp, has_m, has_p, c8), k, has_m, has_p, ka, kb, glob);
                float4* og = reinterpret_cast<float4*>(a.oglob + f * a.C + c8);
                og[0] = make_float4(glob[0], glob[1], glob[2], glob[3]);
                og[1] = make_float4(glob[4], glob[5], glob[6], glob[7]);
            }
            for (int v0 = 0; v0 < V; v0 += U) {
                MxRaw raw[U];
#pragma unroll
                for (int u = 0; u < U; ++u)
                    if (v0 + u < V) raw[u] = mx_issue(a, k, f, fi, v0 + u, Vp, has_m, has_p, c8);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (v0 + u >= V) break;
                    float val[8];
                    mx_finish(raw[u], k, has_m, has_p, ka, kb, val);
                    const float ac = addc_s[v0 + u];
#pragma unroll
                    for (int e = 0; e < 8; ++e) val[e] = fmaf(glob[e], ac, val[e]);
                    const uint4 pk = pack8(val);
                    if (a.stat_sum) {
                        unpack8(pk, val);                         // statistics of the stored (rounded) values
#pragma unroll
                        for (int e = 0; e < 8; ++e) { s1[e] += val[e]; s2[e] += val[e] * val[e]; }
                    }
                    *reinterpret_cast<uint4*>(feat + (f * V + v0 + u) * a.ld_feat + c8) = pk;
                }
            }
        }
    if (a.stat_sum) {
        if (fl < lanes) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { s_red[0][fl * a.C + c8 + e] = s1[e]; s_red[1][fl * a.C + c8 + e] = s2[e]; }
        }
        __syncthreads();
        for (int c = tid; c < a.C; c += MX_THREADS) {
            float t1 = 0.f, t2 = 0.f;
            for (int l = 0; l < lanes; ++l) { t1 += s_red[0][l * a.C + c]; t2 += s_red[1][l * a.C + c]; }
            atomicAdd(a.stat_sum + c, (double)t1);
            atomicAdd(a.stat_sq + c, (double)t2);
        }
    }
}

// dfeat (two-tensor BatchNorm-backward form) -> 8 channels
struct MxD { float a1[8], b[8], a2[8]; };
DSG_D void mx_dcoefs(const dsg_ms_combine_args& a, int c8, MxD& d) {
    float t[8];
    load8f(a.dfeat.a1 ? a.dfeat.a1 + c8 : nullptr, d.a1, 1.f);
    load8f(a.dfeat.b1 ? a.dfeat.b1 + c8 : nullptr, d.b, 0.f);
    load8f(a.dfeat.b2 ? a.dfeat.b2 + c8 : nullptr, t, 0.f);
#pragma unroll
    for (int e = 0; e < 8; ++e) d.b[e] += t[e];
    load8f(a.dfeat.a2 ? a.dfeat.a2 + c8 : nullptr, d.a2, 1.f);
}

__global__ void __launch_bounds__(MX_THREADS) ms_mix_bwd_o_kernel(dsg_ms_combine_args a, int frames_per_cta) {
    DSG_SHARED float s_dadd[32], addc_s[32];
    const int tid = threadIdx.x, nch = a.C >> 3, lanes = MX_THREADS / nch;
    const int cc = tid % nch, fl = tid / nch, c8 = cc * 8;
    const int V = a.V, Vp = a.V + a.has_ext;
    if (tid < 32) { s_dadd[tid] = 0.f; addc_s[tid] = (a.has_ext && tid < V) ? a.add_coeff[tid] : 0.f; }
    __syncthreads();
    MxD dc;
    mx_dcoefs(a, c8, dc);
    const long long n_frames = (long long)a.n_samples * a.T_out;
    long long fend = (long long)(blockIdx.x + 1) * frames_per_cta;
    if (fend > n_frames) fend = n_frames;
    bf16* dO = reinterpret_cast<bf16*>(a.d_o);
    // the loop bounds are CTA-uniform (warp shuffles inside)
    for (long long f0 = (long long)blockIdx.x * frames_per_cta; f0 < fend; f0 += lanes) {
        const long long f = f0 + fl;
        const bool ok = fl < lanes && f < fend;
        float og[8], gsum[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) { og[e] = 0.f; gsum[e] = 0.f; }
        if (ok && a.has_ext) load8f(a.oglob + f * a.C + c8, og, 0.f);
        for (int v0 = 0; v0 < V; v0 += MX_U) {
            uint4 r1[MX_U], r2[MX_U];
#pragma unroll
            for (int u = 0; u < MX_U; ++u) {
                r1[u] = r2[u] = make_uint4(0u, 0u, 0u, 0u);
                if (ok && v0 + u < V) {
                    r1[u] = mx_ld(a.dfeat.x1, f * V + v0 + u, a.dfeat.ld1, c8);
                    if (a.dfeat.x2) r2[u] = mx_ld(a.dfeat.x2, f * V + v0 + u, a.dfeat.ld2, c8);
                }
            }
#pragma unroll
            for (int u = 0; u < MX_U; ++u) {
                if (v0 + u >= V) break;
                float d[8], x[8];
                unpack8(r1[u], x);
#pragma unroll
                for (int e = 0; e < 8; ++e) d[e] = fmaf(x[e], dc.a1[e], dc.b[e]);
                if (a.dfeat.x2) {
                    unpack8(r2[u], x);
#pragma unroll
                    for (int e = 0; e < 8; ++e) d[e] = fmaf(x[e], dc.a2[e], d[e]);
                }
                float part = 0.f;
                if (ok) {
                    const uint4 pk = pack8(d);
                    *reinterpret_cast<uint4*>(dO + (f * Vp + v0 + u) * a.ld_do + c8) = pk;
                    const float ac = addc_s[v0 + u];
#pragma unroll
                    for (int e = 0; e < 8; ++e) { gsum[e] = fmaf(d[e], ac, gsum[e]); part = fmaf(d[e], og[e], part); }
                }
                if (a.has_ext) {
                    part = warp_sum(part);
                    if ((tid & 31) == 0) atomicAdd(&s_dadd[v0 + u], part);
                }
            }
        }
        if (ok && a.has_ext) *reinterpret_cast<uint4*>(dO + (f * Vp + V) * a.ld_do + c8) = pack8(gsum);
    }
    if (a.has_ext) {
        __syncthreads();
        if (tid < V) atomicAdd(a.dadd_coeff + tid, s_dadd[tid]);
    }
}

// e of the max / pass ranges.  Thread = (sample, joint row, 8-channel chunk of the span [ac0*8, (ac0+nac)*8), run of MX_TT input
// frames): the thread slides along t, so relu(bn(B)) at t-2 .. t+2 is a register ring (one new 16-byte load per frame instead of
// five) and all index arithmetic is done once per run.
constexpr int MX_TT = 10;
__global__ void __launch_bounds__(MX_THREADS, 2) ms_mix_bwd_e_kernel(dsg_ms_combine_args a, int ac0, int nac, int nseg) {
    DSG_SHARED float s_red[2][1024];
    const int tid = threadIdx.x;
    const int Vp = a.V + a.has_ext, s = a.stride;
    // item = ((n * nseg + seg) * Vp + j) * nacp + ac : chunks of a row are adjacent lanes (nacp = nac rounded up to a power of two,
    // so the lanes of a warp that own the same chunk are a fixed stride apart: shuffle reduction of the statistics), rows next
    int nacp = 1;
    while (nacp < nac) nacp <<= 1;
    const long long n_items = (long long)a.n_samples * nseg * Vp * nacp;
    const long long item = (long long)blockIdx.x * MX_THREADS + tid;
    const int ac = (int)(item & (nacp - 1));
    const bool live = item < n_items && ac < nac;
    const long long r1 = item / nacp;
    const int j = live ? (int)(r1 % Vp) : 0;
    const long long r2 = r1 / Vp;
    const int seg = live ? (int)(r2 % nseg) : 0, n = live ? (int)(r2 / nseg) : 0;
    const int c8 = (ac0 + ac) * 8;
    const MxKinds k = mx_kinds(a, c8);
    float ka[8], kb[8], s1[8], s2[8];
    mx_coefs(a, c8, ka, kb);
#pragma unroll
    for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
    bf16* E = reinterpret_cast<bf16*>(a.e);
    const bool rmw = (k.mx | k.pass) != 0xffu;           // chunk shared with channels this pass does not own (conv range)
    if (live && (k.mx | k.pass)) {
        const int t_beg = seg * MX_TT, t_end = t_beg + MX_TT < a.T_in ? t_beg + MX_TT : a.T_in;
        const bf16* Bp = reinterpret_cast<const bf16*>(a.b.x1) + ((long long)n * a.T_in * Vp + j) * a.b.ld1 + c8;     // frame 0 of this column
        const long long bstep = (long long)Vp * a.b.ld1;
        const bf16* Dp = reinterpret_cast<const bf16*>(a.d_o) + ((long long)n * a.T_out * Vp + j) * a.ld_do + c8;
        const long long dstep = (long long)Vp * a.ld_do;
        bf16* Ep = E + ((long long)n * a.T_in * Vp + j) * a.ld_e + c8;
        const long long estep = (long long)Vp * a.ld_e;
        // ring of the RAW 16-byte chunks of B at frames t-2 .. t+2 (20 registers; relu(bn(.)) is re-evaluated per use, element by
        // element, so few values are live at a time and two CTAs fit an SM); rin bit i: frame t-2+i is inside the sample
        uint4 rb[5];
        unsigned rin = 0u;
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const int t2 = t_beg - 2 + i;
            const bool in = k.mx && t2 >= 0 && t2 < a.T_in;
            rb[i] = in ? *reinterpret_cast<const uint4*>(Bp + t2 * bstep) : make_uint4(0u, 0u, 0u, 0u);
            rin |= in ? (1u << i) : 0u;
        }
        // gradient windows of the first frame
        uint4 dq[3];
        unsigned din = 0u;
#pragma unroll
        for (int w = 0; w < 3; ++w) {                    // window t' with s*t' + dt == t, dt = w - 1
            const int num = t_beg - (w - 1);
            const bool in = (w == 1 || k.mx) && num >= 0 && num % s == 0 && num / s < a.T_out;
            dq[w] = in ? *reinterpret_cast<const uint4*>(Dp + (num / s) * dstep) : make_uint4(0u, 0u, 0u, 0u);
            din |= in ? (1u << w) : 0u;
        }
        uint4 old = rmw ? *reinterpret_cast<const uint4*>(Ep + t_beg * estep) : make_uint4(0u, 0u, 0u, 0u);
        for (int t = t_beg; t < t_end; ++t) {
            // ---- issue the loads of the NEXT frame before touching this one (software pipeline: two frames in flight)
            const bool more = t + 1 < t_end;
            const int tn = t + 3;
            const bool nin = k.mx && more && tn < a.T_in;
            const uint4 nraw = nin ? *reinterpret_cast<const uint4*>(Bp + tn * bstep) : make_uint4(0u, 0u, 0u, 0u);
            uint4 ndq[3];
            unsigned ndin = 0u;
#pragma unroll
            for (int w = 0; w < 3; ++w) {
                const int num = t + 1 - (w - 1);
                const bool in = more && (w == 1 || k.mx) && num >= 0 && num % s == 0 && num / s < a.T_out;
                ndq[w] = in ? *reinterpret_cast<const uint4*>(Dp + (num / s) * dstep) : make_uint4(0u, 0u, 0u, 0u);
                ndin |= in ? (1u << w) : 0u;
            }
            const uint4 nold = (rmw && more) ? *reinterpret_cast<const uint4*>(Ep + (t + 1) * estep) : make_uint4(0u, 0u, 0u, 0u);
            // ---- this frame
            uint4 pk;
            float bsum[8];
            if (k.mx == 0u) {
                // pass-only chunk (and conv / foreign channels kept): e = dO of the window that maps onto t, a word-wise select
                const uint4 d1 = (din >> 1 & 1u) ? dq[1] : make_uint4(0u, 0u, 0u, 0u);
                const uint32_t* dw_ = &d1.x;
                const uint32_t* ow_ = &old.x;
                uint32_t o4[4];
#pragma unroll
                for (int wd = 0; wd < 4; ++wd) {
                    const uint32_t mlo = (k.pass >> (2 * wd) & 1u) ? 0x0000ffffu : 0u, mhi = (k.pass >> (2 * wd + 1) & 1u) ? 0xffff0000u : 0u;
                    const uint32_t m = mlo | mhi;
                    o4[wd] = (dw_[wd] & m) | (ow_[wd] & ~m);
                }
                pk = make_uint4(o4[0], o4[1], o4[2], o4[3]);
#pragma unroll
                for (int e = 0; e < 8; ++e) bsum[e] = 0.f;
            } else {
                const uint32_t* rbw[5] = {&rb[0].x, &rb[1].x, &rb[2].x, &rb[3].x, &rb[4].x};
                const uint32_t* dqw[3] = {&dq[0].x, &dq[1].x, &dq[2].x};
                const uint32_t* oldw = &old.x;
                uint32_t outw[4];
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int wd = e >> 1;
                    const bool hi16 = e & 1;
                    auto f16 = [&](uint32_t u) { return __uint_as_float(hi16 ? (u & 0xffff0000u) : (u << 16)); };
                    float ev = f16(oldw[wd]);
                    const float braw = f16(rbw[2][wd]);
                    if (k.pass >> e & 1u) ev = (din >> 1 & 1u) ? f16(dqw[1][wd]) : 0.f;
                    else if (k.mx >> e & 1u) {
                        // h[i] = relu(bn(B)) at frame t-2+i, -1 outside the sample.  The gradient of window t' (= t - dt) reaches t iff t
                        // is that window's FIRST maximum (ATen max_pool2d): strictly above the earlier frames, not below the later ones
                        float h[5];
#pragma unroll
                        for (int i = 0; i < 5; ++i) h[i] = (rin >> i & 1u) ? fmaxf(fmaf(f16(rbw[i][wd]), ka[e], kb[e]), 0.f) : -1.f;
                        const bool pos = h[2] > 0.f;
                        const bool w_m1 = pos && h[2] >= h[3] && h[2] >= h[4];      // dt = -1: window frames t, t+1, t+2
                        const bool w_0 = pos && h[2] > h[1] && h[2] >= h[3];        // dt =  0: t-1, t, t+1
                        const bool w_p1 = pos && h[2] > h[0] && h[2] > h[1];        // dt = +1: t-2, t-1, t
                        ev = 0.f;
                        if (w_m1 && (din & 1u)) ev += f16(dqw[0][wd]);
                        if (w_0 && (din >> 1 & 1u)) ev += f16(dqw[1][wd]);
                        if (w_p1 && (din >> 2 & 1u)) ev += f16(dqw[2][wd]);
                    }
                    bsum[e] = braw;
                    // round to bf16 exactly as pack8 does, two elements per word
                    if (!hi16) outw[wd] = __float_as_uint(ev);
                    else outw[wd] = dsg_pack_bf16x2(__uint_as_float(outw[wd]), ev);
                }
                pk = make_uint4(outw[0], outw[1], outw[2], outw[3]);
            }
            *reinterpret_cast<uint4*>(Ep + t * estep) = pk;
            if (a.e_sum && k.mx) {
                float x[8];
                unpack8(pk, x);
#pragma unroll
                for (int e = 0; e < 8; ++e) { s1[e] += x[e]; s2[e] += x[e] * bsum[e]; }
            }
            // ---- slide
            rb[0] = rb[1]; rb[1] = rb[2]; rb[2] = rb[3]; rb[3] = rb[4]; rb[4] = nraw;
            rin = (rin >> 1) | (nin ? 16u : 0u);
            dq[0] = ndq[0]; dq[1] = ndq[1]; dq[2] = ndq[2];
            din = ndin;
            old = nold;
        }
    }
    if (a.e_sum) {
        // lanes ac, ac + nacp, ... of a warp own the same chunk: xor-shuffle over the lane bits above nacp, then one shared-memory
        // pass over the warps and one atomic per (chunk, channel) and CTA; max range only
        const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            for (int o = 16; o >= nacp; o >>= 1) {
                s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], o);
                s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], o);
            }
        }
        float* red = &s_red[0][0];                        // [warp][nacp <= 16][16] floats = 8 * 16 * 16 <= 2 * MX_THREADS * 4
        if (nacp <= 16 && lane < nacp) {
#pragma unroll
            for (int e = 0; e < 8; ++e) { red[(warp * nacp + lane) * 16 + e] = s1[e]; red[(warp * nacp + lane) * 16 + 8 + e] = s2[e]; }
        }
        __syncthreads();
        for (int i = tid; i < nac * 16; i += MX_THREADS) {
            const int c = i >> 4, q = i & 15, e = q & 7;
            if (!(mx_kinds(a, (ac0 + c) * 8).mx >> e & 1u)) continue;
            float t = 0.f;
            for (int w = 0; w < MX_THREADS / 32; ++w) t += red[(w * nacp + c) * 16 + q];
            atomicAdd((q < 8 ? a.e_sum : a.e_sq) + (ac0 + c) * 8 + e, (double)t);
        }
    }
}

static inline bool ms_mix_ok(const dsg_ms_combine_args& a) {
    if (a.dtype != DSG_BF16 || a.C % 8 != 0 || a.C > 256 || a.C < 8 || MX_THREADS % (a.C / 8) != 0 || a.V + a.has_ext > 32) return false;
    if (a.b.x2 || a.b.a2 || !act8_ok(a.b)) return false;
    return true;
}
static inline bool al16(const void* p, long long ld) { return p != nullptr && (uintptr_t)p % 16 == 0 && ld % 8 == 0; }

static const char* launch_ms_mix_fwd(const dsg_ms_combine_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (!ms_mix_ok(a) || !al16(a.feat, a.ld_feat) || (a.conv_hi > a.conv_lo && !al16(a.o, a.ld_o))) return nullptr;
    if (a.has_ext && (!a.oglob || (uintptr_t)a.oglob % 16 != 0 || !a.add_coeff)) return nullptr;
    const long long n_frames = (long long)a.n_samples * a.T_out;
    if (n_frames <= 0) { *handled = true; return nullptr; }
    const int lanes = MX_THREADS / (a.C / 8);
    int fpc = lanes * 2;
    // rows in flight per thread x CTAs per SM: <2, 2> (128 registers, 16 warps per SM) 1.01 ms per step, <5, 1> (224 registers) 1.12 ms,
    // <3, 2> (spills) 1.39 ms — 128 clips, B200
    static const int var = [] { const char* e = getenv("DSG_MIX_FWD_VAR"); return (e && e[0]) ? atoi(e) : 2; }();
    const dim3 grid((unsigned)((n_frames + fpc - 1) / fpc));
    if (var == 1) dsg_launch(ms_mix_fwd_kernel<3, 2>, grid, dim3(MX_THREADS), 0, st, a, fpc);
    else if (var == 2) dsg_launch(ms_mix_fwd_kernel<2, 2>, grid, dim3(MX_THREADS), 0, st, a, fpc);
    else dsg_launch(ms_mix_fwd_kernel<5, 1>, grid, dim3(MX_THREADS), 0, st, a, fpc);
    *handled = true;
    return dsg_launch_error();
}

static const char* launch_ms_mix_bwd(const dsg_ms_combine_args& a, int parts, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (!a.d_o_full || !ms_mix_ok(a) || !act8_ok(a.dfeat) || !al16(a.d_o, a.ld_do) || a.ld_do < a.C || !al16(a.e, a.ld_e)) return nullptr;
    if (a.has_ext && (!a.oglob || (uintptr_t)a.oglob % 16 != 0 || !a.add_coeff || !a.dadd_coeff)) return nullptr;
    const long long n_out = (long long)a.n_samples * a.T_out;
    if (n_out <= 0) { *handled = true; return nullptr; }
    int lo = 1 << 30, hi = 0;
    if (a.max_hi > a.max_lo) { lo = a.max_lo < lo ? a.max_lo : lo; hi = a.max_hi > hi ? a.max_hi : hi; }
    if (a.pass_hi > a.pass_lo) { lo = a.pass_lo < lo ? a.pass_lo : lo; hi = a.pass_hi > hi ? a.pass_hi : hi; }
    const int ac0 = hi > lo ? lo >> 3 : 0, nac = hi > lo ? ((hi + 7) >> 3) - ac0 : 0;
    int nacp = 1;
    while (nacp < nac) nacp <<= 1;
    if (nacp > 16) return nullptr;                         // before anything is launched: the scalar kernels take the whole call
    if (parts & 1) {
        const int lanes = MX_THREADS / (a.C / 8);
        const int fpc = lanes * 2;
        dsg_launch(ms_mix_bwd_o_kernel, dim3((unsigned)((n_out + fpc - 1) / fpc)), dim3(MX_THREADS), 0, st, a, fpc);
        if (const char* e = dsg_launch_error()) return e;
    }
    if ((parts & 2) && nac > 0) {
        const int nseg = (a.T_in + MX_TT - 1) / MX_TT;
        const long long n_items = (long long)a.n_samples * nseg * (a.V + a.has_ext) * nacp;
        dsg_launch(ms_mix_bwd_e_kernel, dim3((unsigned)((n_items + MX_THREADS - 1) / MX_THREADS)), dim3(MX_THREADS), 0, st, a, ac0, nac, nseg);
        if (const char* e = dsg_launch_error()) return e;
    }
    *handled = true;
    return nullptr;
}

}  // namespace dsg
