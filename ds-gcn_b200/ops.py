"""Tensor-level wrappers of the C ABI (one Python function per entry point).

Activations are passed as 2-D (or N-D, flattened) tensors whose last dimension is the
channel axis (stride 1) and whose rows have a uniform pitch `ld` — i.e. channels-last
storage of the reference's logical [n, C, t, v] tensors, possibly a channel slice of a
wider buffer.
"""
import ctypes as C

import torch

from . import _lib as L


def _ld(t):
    """row pitch (elements) of an activation tensor [..., C]"""
    assert t.stride(-1) == 1 or t.shape[-1] == 1, "channel axis must be contiguous"
    if t.dim() == 1:
        return t.shape[0]
    ld = t.stride(-2)
    # every leading dim must be consistent with a single row pitch
    exp = ld
    for d in range(t.dim() - 2, -1, -1):
        assert t.shape[d] == 1 or t.stride(d) == exp, f"activation is not row-uniform: {t.shape} {t.stride()}"
        exp *= t.shape[d]
    return ld


def _nrows(t):
    n = 1
    for s in t.shape[:-1]:
        n *= s
    return n


class Act:
    """value = f(a1*x1 + b1 + a2*x2 + b2)   (include/dsgcn_b200.h: dsg_act_src)"""

    def __init__(self, x1, a1=None, b1=None, x2=None, a2=None, b2=None, relu=False):
        self.x1, self.a1, self.b1, self.x2, self.a2, self.b2, self.relu = x1, a1, b1, x2, a2, b2, relu
        if x2 is not None:
            assert x2.dtype == x1.dtype and x2.shape == x1.shape, (x1.shape, x2.shape)
        for c in (a1, b1, a2, b2):
            if c is not None:
                assert c.dtype == torch.float32 and c.is_contiguous() and c.numel() == x1.shape[-1]

    def struct(self):
        s = L.ActSrc()
        s.x1 = L.ptr(self.x1)
        s.ld1 = _ld(self.x1)
        s.x2 = L.ptr(self.x2)
        s.ld2 = _ld(self.x2) if self.x2 is not None else 0
        s.a1, s.b1, s.a2, s.b2 = L.ptr(self.a1), L.ptr(self.b1), L.ptr(self.a2), L.ptr(self.b2)
        s.relu = int(self.relu)
        return s

    @property
    def dtype(self):
        return self.x1.dtype

    @property
    def C(self):
        return self.x1.shape[-1]


def as_act(x):
    return x if isinstance(x, Act) else Act(x)


def _f32(t):
    assert t is None or (t.dtype == torch.float32 and t.is_contiguous()), "parameters must be contiguous fp32"
    return t


def conv_gemm(src, W, N, out, *, n_samples, T_in, T_out, Vin, ws=None, bias=None, taps=1, tap_step=0, tap_off=0,
              t_mul=1, t_div=1, ext_in=False, contract_ext=False, add=None, add2=None, bcast=None, bcast_scale=1.0,
              mask=None, stat_sum=None, stat_sq=None, partner=None, out_f32=False, adyn=None, y_out=None):
    """dsg_conv_gemm.  W fp32 with strides ws=(ws_n, ws_k, ws_tap); default = PyTorch conv weight
    [N, K, taps, 1] (or [N, K]).  out_f32: bf16 sources, fp32 `out` (unrounded accumulator; tcgen05 engine only)."""
    src = as_act(src)
    a = L.ConvGemmArgs()
    a.src = src.struct()
    a.dtype = L.dt(src.dtype)
    K = src.C
    a.K, a.N = K, N
    a.W = L.ptr(_f32(W))
    if ws is None:
        ws = (K * taps, taps, 1)
    a.ws_n, a.ws_k, a.ws_tap = ws
    a.bias = L.ptr(_f32(bias))
    a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div = taps, tap_step, tap_off, t_mul, t_div
    a.n_samples, a.T_in, a.T_out, a.Vin = n_samples, T_in, T_out, Vin
    a.ext_in, a.contract_ext = int(ext_in), int(contract_ext)
    assert out.shape[-1] == N
    if out_f32:
        assert src.dtype == torch.bfloat16 and out.dtype == torch.float32
        a.out_f32 = 1
    else:
        assert out.dtype == src.dtype
    a.out, a.ld_out = L.ptr(out), _ld(out)
    if adyn is not None:
        # fused spatial graph convolution: rows are contracted per frame with adyn[n,u,w,k] inside the kernel (include/dsgcn_b200.h)
        assert adyn.dtype == torch.bfloat16 and adyn.is_contiguous() and adyn.shape == (n_samples, Vin, Vin, K)
        a.adyn = L.ptr(adyn)
        if y_out is not None:
            assert y_out.dtype == torch.bfloat16 and y_out.shape[-1] == K
            a.y_out, a.ld_y = L.ptr(y_out), _ld(y_out)
    if add is not None:
        assert add.dtype == src.dtype
        a.add, a.ld_add = L.ptr(add), _ld(add)
    if add2 is not None:
        assert add2.dtype == src.dtype
        a.add2, a.ld_add2 = L.ptr(add2), _ld(add2)
    if bcast is not None:
        a.bcast, a.bcast_scale = L.ptr(_f32(bcast)), float(bcast_scale)
    if mask is not None:
        mask = as_act(mask)
        assert mask.dtype == src.dtype
        a.has_mask, a.mask = 1, mask.struct()
    if stat_sum is not None:
        assert stat_sum.dtype == torch.float64 and stat_sq.dtype == torch.float64
        a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    if partner is not None:
        assert partner.dtype == src.dtype
        a.partner, a.ld_partner = L.ptr(partner), _ld(partner)
    wpack = None
    if src.dtype == torch.bfloat16 and taps == 1 and K % 8 == 0 and N % 8 == 0:
        nb = int(L.lib().dsg_conv_gemm_wpack_bytes(K, N))       # 0 in the simulator build
        if nb > 0:
            wpack = torch.empty(nb, dtype=torch.uint8, device=out.device)     # stream-ordered: free to die after the call
            a.wpack = L.ptr(wpack)
    es = out.element_size()
    rows_out = n_samples * T_out * (Vin + int(ext_in) - int(contract_ext))
    nbytes = n_samples * T_in * Vin * K * es * (2 if src.x2 is not None else 1) + rows_out * N * es
    for extra in (add, add2, partner):
        if extra is not None:
            nbytes += rows_out * N * es
    if mask is not None:
        nbytes += rows_out * N * es
    tag = None
    if L.profile is not None:
        tag = dict(K=K, N=N, rows_out=rows_out, taps=taps, ext_in=int(ext_in), cext=int(contract_ext), x2=src.x2 is not None, fused=adyn is not None,
                   mask=mask is not None, stats=stat_sum is not None, ws=tuple(ws), t_mul=t_mul, t_div=t_div, dt=str(src.dtype))
    L.call("dsg_conv_gemm", C.byref(a), L.stream(), nbytes=nbytes, tag=tag)
    return out


def conv_wgrad(A, B, dW, *, n_samples, T_in, T_out, Vin, ws=None, db=None, taps=1, tap_step=0, tap_off=0,
               t_mul=1, t_div=1, ext_in=False):
    A, B = as_act(A), as_act(B)
    assert A.dtype == B.dtype
    a = L.ConvWgradArgs()
    a.A, a.B = A.struct(), B.struct()
    a.dtype = L.dt(A.dtype)
    a.K, a.N = A.C, B.C
    a.dW = L.ptr(_f32(dW))
    if ws is None:
        ws = (a.K * taps, taps, 1)
    a.ws_n, a.ws_k, a.ws_tap = ws
    a.db = L.ptr(_f32(db))
    a.taps, a.tap_step, a.tap_off, a.t_mul, a.t_div = taps, tap_step, tap_off, t_mul, t_div
    a.n_samples, a.T_in, a.T_out, a.Vin, a.ext_in = n_samples, T_in, T_out, Vin, int(ext_in)
    es = A.x1.element_size()
    rows_out = n_samples * T_out * (Vin + int(ext_in))
    nbytes = n_samples * T_in * Vin * a.K * es * (2 if A.x2 is not None else 1) + rows_out * a.N * es * (2 if B.x2 is not None else 1)
    tag = dict(K=a.K, N=a.N, rows_out=rows_out, taps=taps, ext_in=int(ext_in)) if L.profile is not None else None
    with L.side_stream():       # weight gradients are off the critical path: overlap them with the data-gradient chain
        L.keepalive.extend((A, B, dW, db))
        L.call("dsg_conv_wgrad", C.byref(a), L.stream(), nbytes=nbytes, tag=tag)


def bn_job(mode, Cn, *, sum=None, sq=None, count=1.0, gamma=None, beta=None, running_mean=None, running_var=None,
           save_mean=None, save_invstd=None, a=None, b=None, c=None, dgamma=None, dbeta=None, momentum=0.1, eps=1e-5):
    j = L.BnJob()
    j.mode, j.C = mode, Cn
    j.sum, j.sq, j.count = L.ptr(sum), L.ptr(sq), float(count)
    j.gamma, j.beta = L.ptr(gamma), L.ptr(beta)
    j.running_mean, j.running_var = L.ptr(running_mean), L.ptr(running_var)
    j.save_mean, j.save_invstd = L.ptr(save_mean), L.ptr(save_invstd)
    j.a, j.b, j.c = L.ptr(a), L.ptr(b), L.ptr(c)
    j.dgamma, j.dbeta = L.ptr(dgamma), L.ptr(dbeta)
    j.momentum, j.eps = momentum, eps
    return j


def bn_finalize(jobs):
    if not jobs:
        return
    arr = (L.BnJob * len(jobs))(*jobs)
    L.call("dsg_bn_finalize", arr, len(jobs), L.stream())


def tmean(x, n_samples, T, V, with_bf16=False, out=None):
    """x [n*T*V, C] -> xm [n, V, C] fp32 (and, with_bf16, a bf16 copy: operand of the tensor-core topology GEMMs).
    `out`: caller-allocated (xm[, xb]) — callers that launch on the side stream allocate on the main one."""
    Cn = x.shape[-1]
    xm = out[0] if out is not None else torch.empty((n_samples, V, Cn), dtype=torch.float32, device=x.device)
    if with_bf16:
        xb = out[1] if out is not None else torch.empty((n_samples, V, Cn), dtype=torch.bfloat16, device=x.device)
        L.call("dsg_tmean2", L.ptr(x), L.dt(x), _ld(x), n_samples, T, V, Cn, L.ptr(xm), L.ptr(xb), L.stream())
        return xm, xb
    L.call("dsg_tmean", L.ptr(x), L.dt(x), _ld(x), n_samples, T, V, Cn, L.ptr(xm), L.stream())
    return xm


def topology_args(H, n, V, R, node_type, edge_type, A, alpha, beta, We, be, S, plain=False, subset_wise=True):
    a = L.TopologyArgs()
    a.H, a.ld_h = L.ptr(H), _ld(H)
    a.n_samples, a.V, a.R = n, V, R
    a.variant, a.subset_wise = int(plain), int(subset_wise)
    if not plain:
        assert node_type.dtype == torch.int32 and edge_type.dtype == torch.int32
        a.node_type, a.edge_type = L.ptr(node_type), L.ptr(edge_type)
        a.We, a.be = L.ptr(_f32(We)), L.ptr(_f32(be))
    a.A, a.alpha, a.beta = L.ptr(_f32(A)), L.ptr(_f32(alpha)), L.ptr(_f32(beta))
    a.S = L.ptr(S)
    return a


def topology_fwd(H, n, V, R, node_type, edge_type, A, alpha, beta, We, be, adyn, S, plain=False, subset_wise=True):
    a = topology_args(H, n, V, R, node_type, edge_type, A, alpha, beta, We, be, S, plain, subset_wise)
    a.adyn, a.adyn_dtype = L.ptr(adyn), L.dt(adyn)
    L.call("dsg_topology_fwd", C.byref(a), L.stream())


def topology_bwd(H, n, V, R, node_type, edge_type, A, alpha, beta, We, be, S, dadyn, dH, dA, dalpha, dbeta, dWe, dbe, dH_bf16=None,
                 plain=False, subset_wise=True):
    a = topology_args(H, n, V, R, node_type, edge_type, A, alpha, beta, We, be, S, plain, subset_wise)
    assert dadyn.dtype == torch.float32 and dH.dtype == torch.float32 and _ld(dH) == _ld(H)
    a.dadyn, a.dH = L.ptr(dadyn), L.ptr(dH)
    if dH_bf16 is not None:
        assert dH_bf16.dtype == torch.bfloat16 and dH_bf16.is_contiguous() and dH_bf16.shape == dH.shape and dH.is_contiguous()
        a.dH_bf16 = L.ptr(dH_bf16)
    a.dA, a.dalpha, a.dbeta = L.ptr(_f32(dA)), L.ptr(_f32(dalpha)), L.ptr(_f32(dbeta))
    if not plain:
        a.dWe, a.dbe = L.ptr(_f32(dWe)), L.ptr(_f32(dbe))
    L.call("dsg_topology_bwd", C.byref(a), L.stream())


def ctr_topology(H, n, V, R, Cn, A, alpha, W4, b4, *, adyn=None, dadyn=None, dH=None, dH_bf16=None, dA=None, dalpha=None, dW4=None,
                 db4=None):
    """dsg_ctr_topology_fwd (adyn given) / _bwd (dadyn given)"""
    a = L.CtrTopologyArgs()
    a.H, a.ld_h = L.ptr(H), _ld(H)
    a.n_samples, a.V, a.R, a.C = n, V, R, Cn
    a.A, a.alpha, a.W4, a.b4 = L.ptr(_f32(A)), L.ptr(_f32(alpha)), L.ptr(_f32(W4)), L.ptr(_f32(b4))
    if adyn is not None:
        a.adyn, a.adyn_dtype = L.ptr(adyn), L.dt(adyn)
        L.call("dsg_ctr_topology_fwd", C.byref(a), L.stream())
        return
    assert dadyn.dtype == torch.float32 and dH.dtype == torch.float32 and _ld(dH) == _ld(H)
    a.dadyn, a.dH = L.ptr(dadyn), L.ptr(dH)
    if dH_bf16 is not None:
        assert dH_bf16.dtype == torch.bfloat16 and dH_bf16.is_contiguous() and dH_bf16.shape == dH.shape and dH.is_contiguous()
        a.dH_bf16 = L.ptr(dH_bf16)
    a.dA, a.dalpha, a.dW4, a.db4 = L.ptr(_f32(dA)), L.ptr(_f32(dalpha)), L.ptr(_f32(dW4)), L.ptr(_f32(db4))
    L.call("dsg_ctr_topology_bwd", C.byref(a), L.stream())


def graph_agg(src, out, *, mode, n_samples, T, V, KC, adyn=None, A=None, Ksub=0, mask=None, stat_sum=None, stat_sq=None,
              partner=None):
    src = as_act(src)
    a = L.GraphAggArgs()
    a.src, a.dtype, a.mode = src.struct(), L.dt(src.dtype), mode
    a.n_samples, a.T, a.V, a.KC, a.Ksub = n_samples, T, V, KC, Ksub
    if adyn is not None:
        assert adyn.dtype == src.dtype and adyn.is_contiguous()
    a.adyn, a.A = L.ptr(adyn), L.ptr(_f32(A))
    assert out.dtype == src.dtype
    a.out, a.ld_out = L.ptr(out), _ld(out)
    if mask is not None:
        mask = as_act(mask)
        a.has_mask, a.mask = 1, mask.struct()
    if stat_sum is not None:
        a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    if partner is not None:
        a.partner, a.ld_partner = L.ptr(partner), _ld(partner)
    es = out.element_size()
    rows = n_samples * T * V
    nbytes = rows * src.C * es * (2 if src.x2 is not None else 1) + rows * out.shape[-1] * es
    if adyn is not None:
        nbytes += adyn.numel() * es
    if mask is not None:
        nbytes += rows * out.shape[-1] * es
    L.call("dsg_graph_agg", C.byref(a), L.stream(), nbytes=nbytes)
    return out


def graph_agg_dadj(p, dy, dadj, *, n_samples, T, V, KC, is_static=False, Ksub=0):
    p, dy = as_act(p), as_act(dy)
    a = L.GraphAggDadjArgs()
    a.p, a.dy, a.dtype, a.is_static = p.struct(), dy.struct(), L.dt(p.dtype), int(is_static)
    a.n_samples, a.T, a.V, a.KC, a.Ksub = n_samples, T, V, KC, Ksub
    assert dadj.dtype == torch.float32 and dadj.is_contiguous()
    a.dadj = L.ptr(dadj)
    L.call("dsg_graph_agg_dadj", C.byref(a), L.stream())


def pointwise(src, out, *, out_dtype=None, mask=None, stat_sum=None, stat_sq=None, partner=None):
    src = as_act(src)
    a = L.PointwiseArgs()
    a.src, a.dtype = src.struct(), L.dt(src.dtype)
    a.C, a.rows = src.C, _nrows(src.x1)
    if out is not None:
        a.out, a.ld_out, a.out_dtype = L.ptr(out), _ld(out), L.dt(out)
    else:
        a.out_dtype = a.dtype
    if mask is not None:
        mask = as_act(mask)
        assert mask.dtype == src.dtype
        a.has_mask, a.mask = 1, mask.struct()
    if stat_sum is not None:
        a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    if partner is not None:
        a.partner, a.ld_partner, a.partner_dtype = L.ptr(partner), _ld(partner), L.dt(partner)
    es = src.x1.element_size()
    nbytes = a.rows * a.C * es * (2 if src.x2 is not None else 1)
    if out is not None:
        nbytes += a.rows * a.C * out.element_size()
    if mask is not None:
        nbytes += a.rows * a.C * es
    if partner is not None:
        nbytes += a.rows * a.C * partner.element_size()
    L.call("dsg_pointwise", C.byref(a), L.stream(), nbytes=nbytes)
    return out


def ms_combine_args(dtype, n, T_in, T_out, stride, V, has_ext, Cn, ranges, b, o, add_coeff):
    a = L.MsCombineArgs()
    a.dtype = L.dt(dtype)
    a.n_samples, a.T_in, a.T_out, a.stride, a.V, a.has_ext, a.C = n, T_in, T_out, stride, V, int(has_ext), Cn
    (a.conv_lo, a.conv_hi), (a.max_lo, a.max_hi), (a.pass_lo, a.pass_hi) = ranges
    a.b = as_act(b).struct()
    if o is not None:
        a.o, a.ld_o = L.ptr(o), _ld(o)
    a.add_coeff = L.ptr(_f32(add_coeff))
    return a


def ms_combine_fwd(b, o, feat, oglob, *, n, T_in, T_out, stride, V, has_ext, ranges, add_coeff, stat_sum=None, stat_sq=None):
    a = ms_combine_args(feat.dtype, n, T_in, T_out, stride, V, has_ext, feat.shape[-1], ranges, b, o, add_coeff)
    a.feat, a.ld_feat = L.ptr(feat), _ld(feat)
    a.oglob = L.ptr(oglob)
    if stat_sum is not None:
        a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    L.call("dsg_ms_combine_fwd", C.byref(a), L.stream())


def ms_combine_bwd(b, dfeat, d_o, e, oglob, b_raw, *, n, T_in, T_out, stride, V, has_ext, ranges, add_coeff, e_sum, e_sq, dadd_coeff, parts=3,
                   d_o_full=False):
    dfeat = as_act(dfeat)
    a = ms_combine_args(dfeat.dtype, n, T_in, T_out, stride, V, has_ext, dfeat.C, ranges, b, None, add_coeff)
    a.dfeat = dfeat.struct()
    a.d_o, a.ld_do = L.ptr(d_o), _ld(d_o)
    a.d_o_full = int(d_o_full)
    if d_o_full:
        assert d_o.shape[-1] >= dfeat.C
    a.e, a.ld_e = L.ptr(e), _ld(e)
    a.oglob = L.ptr(oglob)
    a.b_raw, a.ld_b = L.ptr(b_raw), _ld(b_raw)
    a.e_sum, a.e_sq = L.ptr(e_sum), L.ptr(e_sq)
    a.dadd_coeff = L.ptr(_f32(dadd_coeff))
    if parts == 3:
        L.call("dsg_ms_combine_bwd", C.byref(a), L.stream())
    else:
        L.call("dsg_ms_combine_bwd_part", C.byref(a), parts, L.stream())


def ms_temporal_args(b, layout, weights, *, n, T_in, T_out, stride, V, has_ext, add_coeff):
    """layout: [(kind, lo, hi, cfg)] as functional.ms_layout; weights: {branch index: (W, bias, dW, db)} for conv branches."""
    b = as_act(b)
    a = L.MsTemporalArgs()
    a.n_samples, a.T_in, a.T_out, a.stride, a.V, a.has_ext, a.C = n, T_in, T_out, stride, V, int(has_ext), b.C
    if len(layout) > 8:
        return None
    a.n_branches = len(layout)
    for j, (kind, lo, hi, cfg) in enumerate(layout):
        br = a.br[j]
        br.lo, br.hi = lo, hi
        if kind == "conv":
            if cfg[0] != 3:
                return None
            br.kind, br.dilation = 0, cfg[1]
            W, bias, dW, db = weights[j]
            br.W, br.bias, br.dW, br.db = L.ptr(_f32(W)), L.ptr(_f32(bias)), L.ptr(_f32(dW)), L.ptr(_f32(db))
        else:
            br.kind, br.dilation = (1 if kind == "max" else 2), 1
    if b.dtype != torch.bfloat16:
        return None
    a.b = b.struct()
    a.add_coeff = L.ptr(_f32(add_coeff))
    return a


def ms_temporal_supported(a):
    if a is None or not L.lib().dsg_ms_temporal_supported(C.byref(a)):
        return False
    nb = int(L.lib().dsg_ms_temporal_wpack_bytes(C.byref(a)))
    a._wpack = torch.empty(max(nb, 16), dtype=torch.uint8, device=torch.device("cuda", torch.cuda.current_device()))
    a.wpack = L.ptr(a._wpack)
    return True


def ms_temporal_fwd(a, feat, oglob, stat_sum, stat_sq):
    a.feat, a.ld_feat = L.ptr(feat), _ld(feat)
    a.oglob = L.ptr(oglob)
    a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    es = feat.element_size()
    nbytes = a.n_samples * (a.T_in * (a.V + a.has_ext) + a.T_out * a.V) * a.C * es
    L.call("dsg_ms_temporal_fwd", C.byref(a), L.stream(), nbytes=nbytes)


def ms_temporal_bwd(a, dfeat, e, oglob, e_sum, e_sq, dadd_coeff, data=True):
    dfeat = as_act(dfeat)
    a.dfeat = dfeat.struct()
    a.e, a.ld_e = L.ptr(e), _ld(e)
    a.oglob = L.ptr(oglob)
    a.e_sum, a.e_sq = L.ptr(e_sum), L.ptr(e_sq)
    a.dadd_coeff = L.ptr(_f32(dadd_coeff))
    es = e.element_size()
    nbytes = a.n_samples * (2 * a.T_in * (a.V + a.has_ext) + (2 if dfeat.x2 is not None else 1) * a.T_out * a.V) * a.C * es
    if data:
        L.call("dsg_ms_temporal_bwd_data", C.byref(a), L.stream(), nbytes=nbytes)
    with L.side_stream():
        L.keepalive.extend((a, dfeat, e, oglob))
        L.call("dsg_ms_temporal_bwd_weight", C.byref(a), L.stream(), nbytes=nbytes)


def ms_conv(src, out, layout, weights, *, n, T_in, T_out, stride, Vr, transposed, mask=None, partner=None, stat_sum=None, stat_sq=None):
    """dsg_ms_conv: every dilated (3 x 1) conv branch of a multi-scale unit in one launch (or their data gradient).
    `layout`/`weights` as ms_temporal_args; src/out [n*T*Vr, >= conv span] bf16 with absolute channel indexing.
    Returns False when the engine declines the shape (the caller runs the per-branch path)."""
    if src.dtype != torch.bfloat16 or not L.is_device_build():
        return False
    a = L.MsConvArgs()
    a.n_samples, a.T_in, a.T_out, a.stride, a.Vr, a.transposed = n, T_in, T_out, stride, Vr, int(transposed)
    nb = 0
    for j, (kind, lo, hi, cfg) in enumerate(layout):
        if kind != "conv":
            continue
        if cfg[0] != 3 or nb >= 8:
            return False
        br = a.br[nb]
        br.kind, br.lo, br.hi, br.dilation = 0, lo, hi, cfg[1]
        W, bias = weights[j][0], weights[j][1]
        br.W, br.bias = L.ptr(_f32(W)), L.ptr(_f32(bias))
        nb += 1
    if nb == 0:
        return False
    a.n_branches = nb
    a.src, a.ld_src = L.ptr(src), _ld(src)
    a.out, a.ld_out = L.ptr(out), _ld(out)
    if mask is not None:
        mask = as_act(mask)
        a.has_mask, a.mask = 1, mask.struct()
    if partner is not None:
        a.partner, a.ld_partner = L.ptr(partner), _ld(partner)
    if stat_sum is not None:
        a.stat_sum, a.stat_sq = L.ptr(stat_sum), L.ptr(stat_sq)
    nbytes_w = int(L.lib().dsg_ms_conv_wpack_bytes(C.byref(a)))
    if nbytes_w <= 0:
        return False
    wpack = torch.empty(nbytes_w, dtype=torch.uint8, device=out.device)
    a.wpack = L.ptr(wpack)
    handled = C.c_int(0)
    span = a.br[nb - 1].hi - a.br[0].lo
    es = 2
    rows_src = n * (T_out if transposed else T_in) * Vr
    rows_dst = n * (T_in if transposed else T_out) * Vr
    nbytes = (rows_src + rows_dst * (1 + int(mask is not None) + int(partner is not None))) * span * es
    L.call("dsg_ms_conv", C.byref(a), C.byref(handled), L.stream(), nbytes=nbytes)
    return bool(handled.value)


def ms_conv_wgrad(H, d_o, layout, wgrads, *, n, T_in, T_out, stride, Vr):
    """dsg_ms_conv_wgrad: weight / bias gradients of every dilated (3 x 1) conv branch in one TMA-fed tcgen05 launch, on the side
    stream like the other weight gradients.  `wgrads`: {branch index: (dW, db)} pre-zeroed fp32 accumulators.  Returns False when
    the engine declines (the caller runs dsg_ms_temporal_bwd_weight)."""
    if H.dtype != torch.bfloat16 or d_o.dtype != torch.bfloat16 or not L.is_device_build():
        return False
    a = L.MsConvArgs()
    a.n_samples, a.T_in, a.T_out, a.stride, a.Vr, a.transposed = n, T_in, T_out, stride, Vr, 0
    nb = 0
    for j, (kind, lo, hi, cfg) in enumerate(layout):
        if kind != "conv":
            continue
        if cfg[0] != 3 or nb >= 8:
            return False
        br = a.br[nb]
        br.kind, br.lo, br.hi, br.dilation = 0, lo, hi, cfg[1]
        dW, db = wgrads[j]
        br.dW, br.db = L.ptr(_f32(dW)), L.ptr(_f32(db))
        nb += 1
    if nb == 0:
        return False
    a.n_branches = nb
    a.src, a.ld_src = L.ptr(H), _ld(H)
    a.out, a.ld_out = L.ptr(d_o), _ld(d_o)
    handled = C.c_int(0)
    span = a.br[nb - 1].hi - a.br[0].lo
    nbytes = (n * T_in * Vr + n * T_out * Vr) * span * 2
    with L.side_stream():
        L.keepalive.extend((a, H, d_o, wgrads))
        L.call("dsg_ms_conv_wgrad", C.byref(a), C.byref(handled), L.stream(), nbytes=nbytes)
    return bool(handled.value)


def head_ce_fwd(pooled, W, b, label=None):
    """dsg_head_ce_fwd: logits [N, K] (+ per-sample {cross-entropy, top-1 hit, top-5 hit} [N, 3] when labels are given)"""
    N, Cn = pooled.shape
    K = W.shape[0]
    assert pooled.dtype == torch.float32 and pooled.is_contiguous() and W.shape[1] == Cn
    logits = torch.empty(N, K, dtype=torch.float32, device=pooled.device)
    stats = None
    if label is not None:
        assert label.dtype == torch.int64 and label.is_contiguous() and label.numel() == N
        stats = torch.empty(N, 3, dtype=torch.float32, device=pooled.device)
    L.call("dsg_head_ce_fwd", L.ptr(pooled), L.ptr(_f32(W)), L.ptr(_f32(b)), L.ptr(label), N, Cn, K, L.ptr(logits), L.ptr(stats), L.stream())
    return logits, stats


def head_ce_bwd(logits, label, pooled, W, gscale, dW=None, db=None, need_dpooled=True):
    """dsg_head_ce_bwd: returns (dlogits, dpooled); accumulates into dW / db when given"""
    N, K = logits.shape
    Cn = W.shape[1]
    assert gscale.dtype == torch.float32 and gscale.numel() == 1
    dlogits = torch.empty_like(logits)
    dpooled = torch.empty(N, Cn, dtype=torch.float32, device=logits.device) if need_dpooled else None
    L.call("dsg_head_ce_bwd", L.ptr(logits), L.ptr(label), L.ptr(pooled), L.ptr(_f32(W)), L.ptr(gscale), N, Cn, K, L.ptr(dlogits),
           L.ptr(dpooled), L.ptr(_f32(dW)), L.ptr(_f32(db)), L.stream())
    return dlogits, dpooled


def sgd_step(p, grad, buf, lr, momentum, wd, nesterov, grad_scale=1.0):
    """`lr`: a Python float, or a one-element fp32 tensor on the device (read by the kernel: CUDA-graph friendly schedules)."""
    assert p.is_contiguous() and grad.is_contiguous() and buf.is_contiguous()
    assert p.dtype == grad.dtype == buf.dtype == torch.float32 and p.numel() == grad.numel() == buf.numel()
    if isinstance(lr, torch.Tensor):
        assert lr.dtype == torch.float32 and lr.numel() == 1
        L.call("dsg_sgd_step_dev", L.ptr(p), L.ptr(grad), L.ptr(buf), p.numel(), L.ptr(lr), momentum, wd, int(nesterov), grad_scale, L.stream())
    else:
        L.call("dsg_sgd_step", L.ptr(p), L.ptr(grad), L.ptr(buf), p.numel(), lr, momentum, wd, int(nesterov), grad_scale, L.stream())
