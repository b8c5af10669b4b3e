// tc4 temporal mode: host side of dsg_ms_conv — the dilated (3 x 1) convolutions of mstcn / dgmstcn (tcn.py:383-391) and their
// data gradient on the TMA-fed engine of tc4_gemm.cuh (G4Plan mode 3).
//
// Implicit GEMM without staging: a tile is F frames (32-row slots) of one sample; for every (branch, tap) the producer issues
// 4-D TMA loads of the branch's 64-channel window at frame + shift — frames outside the sample come back as zeros, which IS the
// convolution's zero padding — and the MMA issuer multiplies the atom into the branch's column window of the accumulator with
// that tap's [window x 64] weight tile.  A temporal stride turns into two parity-plane tensor maps (forward) or one launch per
// destination parity plane (data gradient).  North-star kernel (c): "temporal convolutions as implicit-GEMM kernels".
#pragma once
#include "tc4_gemm.cuh"

#ifndef DSG_EMU
namespace dsg {
namespace tc4 {

struct TcAtom { int br, tap, map, tsh; };

// weight tiles of one launch: per column tile y, per atom: [nw rows (output channel = n0 + ncol + row)] x [64 k (input channel = c0 + k)]
// bf16 in the K-major SWIZZLE_128B layout; element = W[co, ci, tap] (forward) or W[k-side, n-side, tap] (data gradient)
struct TcPackJob { const float* W; int lo, hi, tap, c0, nabs0, nw; unsigned off; };
struct TcPackJobs { TcPackJob j[2 * G4_MAX_OPS]; int n, transposed, span_lo, N; const float* bias[8]; int blo[8], bhi[8], nb; };

__global__ void __launch_bounds__(256) tc4_tconv_wpack_kernel(TcPackJobs jobs, unsigned char* out, float* cbias) {
    if ((int)blockIdx.x == jobs.n) {                      // folded bias per absolute output column (0 for the data gradient)
        for (int c = threadIdx.x; c < jobs.N; c += 256) {
            float v = 0.f;
            if (!jobs.transposed)
                for (int b = 0; b < jobs.nb; ++b)
                    if (jobs.span_lo + c >= jobs.blo[b] && jobs.span_lo + c < jobs.bhi[b] && jobs.bias[b]) v = jobs.bias[b][jobs.span_lo + c - jobs.blo[b]];
            cbias[c] = v;
        }
        return;
    }
    const TcPackJob& J = jobs.j[blockIdx.x];
    const int w = J.hi - J.lo;
    unsigned char* dst = out + J.off;
    for (int idx = threadIdx.x; idx < J.nw * 8; idx += 256) {
        const int ch = idx & 7, nl = idx >> 3;
        const int nabs = J.nabs0 + nl;                      // absolute output channel of this row
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int kabs = J.c0 + ch * 8 + e;
            float x = 0.f;
            if (nabs >= J.lo && nabs < J.hi && kabs >= J.lo && kabs < J.hi) {
                const int co = jobs.transposed ? kabs - J.lo : nabs - J.lo, ci = jobs.transposed ? nabs - J.lo : kabs - J.lo;
                x = J.W[((long long)co * w + ci) * 3 + J.tap];
            }
            v[e] = x;
        }
        *reinterpret_cast<uint4*>(dst + atom_off(nl, ch)) = pack8(v);
    }
}

static inline long long tc4_tconv_wpack_bytes(const dsg_ms_conv_args& a) {
    (void)a;
    return 2LL * G4_MAX_OPS * 64 * 128 + 2LL * 128 * 128 + 256 * 4 + 2048;      // worst case: every op a 64-row tile + the full-width op 0 per column tile, + folded bias
}

struct TcOp { int b, tap, map, tsh, c0, k0, ks, ncol, nw; };

// one launch: destination plane (qs, qp) of `out`.  `windows`: load every 64-channel source column once per tile with its temporal
// halo and let all (branch, tap) ops of the column read it at row offsets (narrow layers: four branches share one column, so the
// tile's source frames cross the L2 -> shared-memory path 12 times instead of 48); else one F-frame atom per (branch, tap).
static const char* tconv_launch(const dsg_ms_conv_args& a, int qs, int qp, bool windows, dsg_stream_t st, bool* handled) {
    *handled = false;
    const int nb = a.n_branches, s = a.stride;
    const int span_lo = a.br[0].lo, span_hi = a.br[nb - 1].hi, N = span_hi - span_lo;
    const int Tsrc = a.transposed ? a.T_out : a.T_in, Tdst = a.transposed ? a.T_in : a.T_out;
    G4Plan p{};
    p.mode = 3;
    p.V = a.Vr;
    p.slot = (a.Vr + 7) & ~7;
    p.F = ATOM_ROWS / p.slot;
    p.qs = qs; p.qp = qp; p.Tdst = Tdst;
    p.Tq = Tdst > qp ? (Tdst - qp + qs - 1) / qs : 0;
    if (p.Tq <= 0) { *handled = true; return nullptr; }
    p.tps = (p.Tq + p.F - 1) / p.F;
    const long long nt = (long long)a.n_samples * p.tps;
    if (nt > 0x3fffffff) return nullptr;
    p.n_tiles = (int)nt;
    p.n_frames = (long long)a.n_samples * Tdst;
    p.rows_out = p.n_frames * a.Vr;
    p.Ntile = N <= 64 ? 64 : 128;
    const int gy = (N + p.Ntile - 1) / p.Ntile;
    if (gy > 2) return nullptr;
    p.stats = a.stat_sum == nullptr ? 0 : (a.partner ? 2 : 1);
    bool use_map1 = false;
    TcPackJobs jobs{};
    jobs.transposed = a.transposed; jobs.span_lo = span_lo; jobs.N = N; jobs.nb = nb;
    for (int b = 0; b < nb; ++b) { jobs.bias[b] = a.br[b].bias; jobs.blo[b] = a.br[b].lo; jobs.bhi[b] = a.br[b].hi; }
    unsigned woff_total = 0;
    int max_nfr = p.F;
    for (int y = 0; y < gy; ++y) {
        const int n0 = span_lo + y * p.Ntile, n1 = (n0 + p.Ntile < span_hi) ? n0 + p.Ntile : span_hi;
        const int Ntp = ((n1 - n0) + 15) & ~15;
        // ---- ops: (branch, tap, source column)
        TcOp ops[G4_MAX_OPS];
        int nops = 0;
        for (int b = 0; b < nb; ++b) {
            const int lo = a.br[b].lo, hi = a.br[b].hi, d = a.br[b].dilation;
            if (hi <= n0 || lo >= n1) continue;
            const int lo8 = lo & ~7;
            if (!windows && hi - lo8 > ATOM_CH) return nullptr;
            for (int tap = 0; tap < 3; ++tap) {
                const int o = (tap - 1) * d;
                int map = 0, tsh = 0;
                if (!a.transposed) {                       // src frame = s*t' + o = s*(t' + (o - par)/s) + par
                    const int par = ((o % s) + s) % s;
                    if (par > 1) return nullptr;
                    map = par; tsh = (o - par) / s;
                } else {                                   // src frame = (t - o)/s with t = qs*q + qp
                    if (qs != s) return nullptr;
                    if ((((qp - o) % s) + s) % s != 0) continue;
                    tsh = (qp - o) / s;
                }
                use_map1 |= map == 1;
                const int c_lo = lo > n0 ? lo : n0, c_hi = hi < n1 ? hi : n1;
                const int ncol = (c_lo - n0) & ~15, nw = (((c_hi - n0) + 15) & ~15) - ncol;
                if (windows) {
                    for (int c0 = lo & ~(ATOM_CH - 1); c0 < hi; c0 += ATOM_CH) {      // 64-aligned source columns the branch touches
                        const int klo = lo > c0 ? lo : c0, khi = hi < c0 + ATOM_CH ? hi : c0 + ATOM_CH;
                        if (nops >= G4_MAX_OPS) return nullptr;
                        const int k0 = (klo - c0) / 16;
                        ops[nops++] = TcOp{b, tap, map, tsh, c0, k0, (khi - c0 + 15) / 16 - k0, ncol, nw};
                    }
                } else {
                    if (nops >= G4_MAX_OPS) return nullptr;
                    ops[nops++] = TcOp{b, tap, map, tsh, lo8, 0, (hi - lo8 + 15) / 16, ncol, nw};
                }
            }
        }
        if (nops == 0) ops[nops++] = TcOp{-1, 0, 0, 0, span_lo & ~7, 0, 1, 0, Ntp};      // no tap reaches this plane: one all-zero op
        // ---- atoms: windows = ops grouped by (map, column); else one atom per op
        int natoms = 0, nplaced = 0;
        bool placed[G4_MAX_OPS] = {false};
        unsigned woff = 0;
        for (int i = 0; i < nops; ++i) {
            if (placed[i]) continue;
            if (natoms >= G4_MAX_ATOMS) return nullptr;
            int tmin = ops[i].tsh, tmax = ops[i].tsh;
            if (windows)
                for (int j = i + 1; j < nops; ++j)
                    if (ops[j].map == ops[i].map && ops[j].c0 == ops[i].c0) { tmin = ops[j].tsh < tmin ? ops[j].tsh : tmin; tmax = ops[j].tsh > tmax ? ops[j].tsh : tmax; }
            p.t_map[y][natoms] = (short)ops[i].map; p.t_c0[y][natoms] = ops[i].c0; p.t_tsh[y][natoms] = (short)tmin;
            p.t_nfr[y][natoms] = (short)(p.F + tmax - tmin);
            if (p.t_nfr[y][natoms] > max_nfr) max_nfr = p.t_nfr[y][natoms];
            p.t_op0[y][natoms] = (short)nplaced;
            for (int j = i; j < nops; ++j) {
                if (placed[j] || (j != i && !(windows && ops[j].map == ops[i].map && ops[j].c0 == ops[i].c0))) continue;
                placed[j] = true;
                TcOp o = ops[j];
                if (nplaced == 0) { o.ncol = 0; o.nw = Ntp; }                      // op 0 initialises the whole accumulator tile
                p.o_row[y][nplaced] = (short)(o.tsh - tmin); p.o_k0[y][nplaced] = (short)o.k0; p.o_ks[y][nplaced] = (short)o.ks;
                p.o_ncol[y][nplaced] = (short)o.ncol; p.o_nw[y][nplaced] = (short)o.nw; p.o_woff[y][nplaced] = woff;
                TcPackJob& J = jobs.j[jobs.n++];
                if (o.b >= 0) { J.W = a.br[o.b].W; J.lo = a.br[o.b].lo; J.hi = a.br[o.b].hi; }
                else { J.W = nullptr; J.lo = 0; J.hi = 0; }
                J.tap = o.tap; J.c0 = o.c0 + o.k0 * 16; J.nabs0 = n0 + o.ncol; J.nw = o.nw; J.off = woff_total + woff;
                woff += (unsigned)o.nw * 128u;
                ++nplaced;
            }
            ++natoms;
        }
        p.t_op0[y][natoms] = (short)nplaced;
        p.t_n[y] = natoms;
        p.t_wbytes[y] = (woff + 1023u) & ~1023u;
        woff_total += p.t_wbytes[y];
    }
    p.t_stage = (unsigned)max_nfr * (unsigned)p.slot * 128u;
    p.t_stage = (p.t_stage + 1023u) & ~1023u;
    p.natoms1 = 0;
    p.natoms = use_map1 ? 1 : 0;
    const unsigned wb = p.t_wbytes[0] > p.t_wbytes[1] ? p.t_wbytes[0] : p.t_wbytes[1];
    const unsigned ob1 = (unsigned)(p.Ntile / ATOM_CH) * ATOM_BYTES;
    const unsigned cf_bytes = (unsigned)(4 * 128 * sizeof(float));
    const unsigned budget = 227u * 1024u - 2048u;
    bool fit = false;
    // tail operands at the destination rows (column 0 = channel span_lo): staged through the TMA ring when they fit (tc4_gemm.cuh)
    dsg_act_src mk = a.mask;
    if (a.has_mask) {
        mk.x1 = reinterpret_cast<const bf16*>(a.mask.x1) + span_lo;
        if (mk.x2) mk.x2 = reinterpret_cast<const bf16*>(a.mask.x2) + span_lo;
    }
    const void* partp = a.partner ? reinterpret_cast<const bf16*>(a.partner) + span_lo : nullptr;
    const G4TailSel sel = tc4_tail_select(nullptr, 0, nullptr, 0, partp, a.ld_partner, a.has_mask, mk);
    for (int pass = sel.ok ? 0 : 1; pass < 2 && !fit; ++pass)
    for (int OB = 2; OB >= 1 && !fit; --OB) {
        const int tn = pass == 0 ? sel.n : 0;
        const unsigned tail_stage = (unsigned)tn * ob1;
        const unsigned fixed = wb + OB * ob1 * (p.stats == 2 ? 2u : 1u) + G4_TB * tail_stage + 1024u + ((cf_bytes + 1023u) & ~1023u) + 1024u;
        const unsigned min_stages = windows ? 2u : 3u;
        if (fixed + min_stages * p.t_stage > budget) continue;
        int S = (int)((budget - fixed) / p.t_stage);
        if (S > tc4_s_cap()) S = tc4_s_cap();
        if (OB == 2 && S < (windows ? 3 : 4) && tc4_s_cap() >= 4) continue;
        p.S = S; p.OB = OB;
        p.drain_defer = (OB >= 2 && tc4_drain_defer()) ? 1 : 0;
        p.off_w = 0;
        p.off_a = wb;
        p.off_out = p.off_a + (unsigned)S * p.t_stage;
        p.out_bytes = ob1;
        p.off_stat = p.off_out + OB * ob1;
        tc4_plan_tails(p, sel, tn);
        p.off_tail = p.off_stat + (p.stats == 2 ? OB * ob1 : 0u);
        p.tail_stage_bytes = tail_stage;
        p.off_ones = p.off_tail + G4_TB * tail_stage;
        p.off_cf = p.off_ones + 1024u;
        p.smem_total = p.off_cf + ((cf_bytes + 1023u) & ~1023u) + 1024u;
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
    if ((long long)woff_total + N * 4 + 256 > tc4_tconv_wpack_bytes(a)) return nullptr;

    // ---- tensor maps
    CUtensorMap mA0, mA1, mO;
    const int src_planes = a.transposed ? 1 : s;
    bool ok = make_map_4d(&mA0, a.src, a.n_samples, Tsrc, a.Vr, span_hi, a.ld_src, 0, src_planes);
    mA1 = mA0;
    if (ok && use_map1) ok = make_map_4d(&mA1, a.src, a.n_samples, Tsrc, a.Vr, span_hi, a.ld_src, 1, src_planes);
    const bf16* outp = reinterpret_cast<const bf16*>(a.out) + span_lo;
    ok = ok && make_map_4d(&mO, outp, a.n_samples, Tdst, a.Vr, N, a.ld_out, qp, qs);
    G4TailMaps tmaps;
    for (int t = 0; t < p.tn && ok; ++t) ok = make_map_4d(&tmaps.m[t], sel.ptr[t], a.n_samples, Tdst, a.Vr, N, sel.ld[t], qp, qs);
    if (!ok) return nullptr;

    // ---- the engine's argument block (column 0 = channel span_lo)
    dsg_conv_gemm_args g{};
    g.dtype = DSG_BF16;
    g.K = ATOM_CH; g.N = N;
    g.taps = 1; g.t_mul = 1; g.t_div = 1;
    g.n_samples = a.n_samples; g.T_in = Tdst; g.T_out = Tdst; g.Vin = a.Vr;
    g.out = const_cast<bf16*>(outp); g.ld_out = a.ld_out;
    g.wpack = a.wpack;
    if (a.has_mask) {
        g.has_mask = 1;
        g.mask = a.mask;
        g.mask.x1 = reinterpret_cast<const bf16*>(a.mask.x1) + span_lo;
        if (g.mask.x2) g.mask.x2 = reinterpret_cast<const bf16*>(a.mask.x2) + span_lo;
        if (g.mask.a1) g.mask.a1 += span_lo;
        if (g.mask.b1) g.mask.b1 += span_lo;
        if (g.mask.a2) g.mask.a2 += span_lo;
        if (g.mask.b2) g.mask.b2 += span_lo;
    }
    if (a.partner) { g.partner = reinterpret_cast<const bf16*>(a.partner) + span_lo; g.ld_partner = a.ld_partner; }
    if (a.stat_sum) { g.stat_sum = a.stat_sum + span_lo; g.stat_sq = a.stat_sq + span_lo; }
    float* cbias = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(a.wpack) + (((size_t)woff_total + 255) & ~(size_t)255));
    tc4_tconv_wpack_kernel<<<dim3((unsigned)jobs.n + 1), dim3(256), 0, st>>>(jobs, reinterpret_cast<unsigned char*>(a.wpack), cbias);
    if (const char* e = dsg_launch_error()) return e;
    int gx = num_sms() / gy;
    if (gx < 1) gx = 1;
    if (gx > p.n_tiles) gx = p.n_tiles;
    const bool tails = g.has_mask || g.partner;
    if (p.stats == 2 && !tails) return "ms_conv: statistics with a partner need the partner";
#define DSG_TC_LAUNCH(TL_, ST_)                                                                                                      \
    do {                                                                                                                             \
        cudaFuncSetAttribute(tc4_gemm_kernel<false, TL_, ST_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_total);      \
        tc4_gemm_kernel<false, TL_, ST_><<<dim3((unsigned)gx, (unsigned)gy), dim3(G4_THREADS), p.smem_total, st>>>(mA0, mA1, mO, g, p, cbias, tmaps); \
    } while (0)
    if (!tails && p.stats == 0) DSG_TC_LAUNCH(false, 0);
    else if (!tails && p.stats == 1) DSG_TC_LAUNCH(false, 1);
    else if (tails && p.stats == 0) DSG_TC_LAUNCH(true, 0);
    else if (tails && p.stats == 1) DSG_TC_LAUNCH(true, 1);
    else DSG_TC_LAUNCH(true, 2);
#undef DSG_TC_LAUNCH
    *handled = true;
    return dsg_launch_error();
}

static const char* launch_ms_conv_tc4(const dsg_ms_conv_args& a, dsg_stream_t st, bool* handled) {
    *handled = false;
    if (!tc4_enabled() || a.n_branches < 1 || a.n_branches > 8 || a.stride < 1 || a.stride > 2 || a.Vr < 1 || a.Vr > 32) return nullptr;
    if (a.br[0].lo % 8 != 0 || !tma_ptr_ok(a.src, a.ld_src) || !tma_ptr_ok(a.out, a.ld_out) || !a.wpack || (uintptr_t)a.wpack % 128 != 0) return nullptr;
    for (int b = 0; b < a.n_branches; ++b) {
        if (a.br[b].kind != 0 || a.br[b].hi <= a.br[b].lo || a.br[b].dilation < 1 || !a.br[b].W) return nullptr;
        if (b > 0 && a.br[b].lo != a.br[b - 1].hi) return nullptr;
    }
    if (a.has_mask && !act8_ok(a.mask)) return nullptr;
    if (a.partner && ((uintptr_t)a.partner % 16 != 0 || a.ld_partner % 8 != 0)) return nullptr;
    if ((a.stat_sum == nullptr) != (a.stat_sq == nullptr)) return "ms_conv: stat_sum and stat_sq go together";
    if (a.n_samples <= 0 || a.T_in <= 0 || a.T_out <= 0) { *handled = true; return nullptr; }
    if (!encode_fn()) return nullptr;
    static const bool win_on = [] { const char* e = getenv("DSG_MS_WINDOWS"); return !(e && e[0] == '0'); }();
    auto launch = [&](const dsg_ms_conv_args& aa, int qs, int qp, bool* h) -> const char* {
        // shared windows when they fit the shared-memory budget (narrow layers), else one atom per (branch, tap)
        if (win_on) {
            const char* e = tconv_launch(aa, qs, qp, true, st, h);
            if (e || *h) return e;
        }
        return tconv_launch(aa, qs, qp, false, st, h);
    };
    if (!a.transposed || a.stride == 1) return launch(a, 1, 0, handled);
    // data gradient through a temporal stride: one launch per parity plane of the destination (all of them or none)
    bool h0 = false, h1 = false;
    const char* e = launch(a, a.stride, 0, &h0);
    if (e || !h0) return e;
    dsg_ms_conv_args a1 = a;                               // the second plane packs its own weight tiles behind the first plane's
    a1.wpack = reinterpret_cast<unsigned char*>(a.wpack) + tc4_tconv_wpack_bytes(a);
    e = launch(a1, a.stride, 1, &h1);
    if (e) return e;
    if (!h1) return "ms_conv: the second parity plane was declined after the first was launched";
    *handled = true;
    return nullptr;
}

}  // namespace tc4
}  // namespace dsg
#endif
