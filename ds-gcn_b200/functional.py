"""Forward / backward orchestration of the DS-GCN units on top of the C-ABI kernels.

Every function works on channels-last 2-D activations `[n*T*V, C]` (compute dtype fp32 or bf16)
and on the *modules'* own parameters (real nn.Conv2d / nn.BatchNorm2d children — SURVEY.md §8b
attribute contract).  BatchNorm never runs as a separate pass: the producer kernel accumulates the
batch statistics in its epilogue, `bn_finalize` turns them into per-channel coefficients and the
consumer kernel applies `a*x+b` (+ReLU, +residual) in its prologue.  Backward mirrors that with
the masked gradient `e` and `dy = ca*e + cb*y + cc`.

Reference semantics: pyskl/models/gcns/utils/gcn.py:2217-2365 (dgphgcn1), :75-94 (unit_gcn),
pyskl/models/gcns/utils/tcn.py:31-32 (unit_tcn), :162-177 (mstcn), :407-428 (dgmstcn),
pyskl/models/gcns/dgstgcn.py:61-65 (DGBlock).
"""
import contextlib
import math
import os

import torch

from . import ops
from .ops import Act

BN_EPS_DEFAULT = 1e-5


# ------------------------------------------------------------------------------------------------
# small helpers
# ------------------------------------------------------------------------------------------------

class _Pending:
    """num_batches_tracked buffers to bump with one multi-tensor add at the end of a forward."""
    stack = []

    @classmethod
    def add(cls, bn):
        if bn.num_batches_tracked is None:
            return
        if cls.stack:
            cls.stack[-1].append(bn.num_batches_tracked)
        else:
            bn.num_batches_tracked.add_(1)


class defer_bn_counters:
    """Context manager: collect the BatchNorm `num_batches_tracked` increments of a whole forward pass and
    apply them with a single torch._foreach_add_ (one launch instead of ~100)."""

    def __enter__(self):
        self.items = []
        _Pending.stack.append(self.items)
        return self

    def __exit__(self, *exc):
        _Pending.stack.pop()
        if self.items and exc[0] is None:
            torch._foreach_add_(self.items, 1)
        return False


def _w2(conv):
    return conv.weight


def _zeros64(n, dev):
    return torch.zeros(n, dtype=torch.float64, device=dev)


class stat_arena:
    """BatchNorm statistics accumulators (fp64 sums the kernels add into) for one forward + backward pass of a backbone, carved
    from ONE pre-zeroed buffer: one memset per step instead of ~100 small fills.  Slices are handed out once and never reused
    before the next `with stat_arena(dev)` (the next forward) zeroes the buffer again; when the arena is exhausted or absent
    (units used stand-alone) a fresh torch.zeros is returned."""
    _pool = {}            # device -> [buffer, next free element]
    SIZE = 1 << 17        # fp64 elements (1 MB): a DS-GCN training step uses ~45k

    def __init__(self, dev):
        self.dev = dev

    def __enter__(self):
        ent = stat_arena._pool.get(self.dev)
        if ent is None:
            ent = [torch.zeros(stat_arena.SIZE, dtype=torch.float64, device=self.dev), 0]
            stat_arena._pool[self.dev] = ent
        else:
            ent[0].zero_()
            ent[1] = 0
        return self

    def __exit__(self, *exc):
        return False          # stays active: the backward pass of this forward draws from the same buffer

    @staticmethod
    def take(rows, C, dev):
        ent = stat_arena._pool.get(dev)
        n = rows * ((C + 1) // 2 * 2)                       # 16-byte aligned rows
        if ent is None or ent[1] + n > stat_arena.SIZE:
            return torch.zeros(rows, C, dtype=torch.float64, device=dev)
        o = ent[1]
        ent[1] = o + n
        return ent[0][o:o + n].view(rows, n // rows)[:, :C]


def _empty32(n, dev):
    return torch.empty(n, dtype=torch.float32, device=dev)


# ---- parameter packing hooks (parallel.GradBuckets) -------------------------------------------------------------------
# When GradBuckets has packed a model, a parameter's gradient lives in a pre-zeroed flat buffer (`p.grad` is a view of it and
# the parameter carries `_dsg_sink`), and parameters that are always used concatenated (pre|down, conv1|conv2|conv1_se, the
# branch 1x1 convolutions) are adjacent in the flat buffers, so the concatenation and the gradient of the concatenation are
# zero-copy views (`module._dsg_flat[key] = (param_view, grad_view)`).  The kernels then accumulate straight into the flat
# gradient buffer: no torch.cat, no per-parameter zero fill, no accumulate-add, no flatten before the all-reduce.
# Without GradBuckets (unit tests, plain torch optimizers) every helper falls back to fresh tensors.

def cat_params(m, key, params, shape):
    fl = getattr(m, "_dsg_flat", None)
    if fl is not None and key in fl:
        return fl[key][0]
    if len(params) == 1:
        return params[0].view(shape)
    return torch.cat([p.reshape(-1) for p in params]).view(shape)


def _sink(p):
    return p.grad if (getattr(p, "_dsg_sink", False) and p.grad is not None) else None


def grad_like(p, zero=True):
    """gradient accumulator for parameter `p`: its flat-buffer view (pre-zeroed by GradBuckets.zero_grad) or a fresh tensor"""
    g = _sink(p)
    if g is not None:
        return g
    return torch.zeros_like(p) if zero else torch.empty_like(p)


def grad_cat(m, key, params, shape, grads):
    """gradient accumulator for cat(params) viewed as `shape`; registers the per-parameter views in `grads`"""
    fl = getattr(m, "_dsg_flat", None)
    if fl is not None and key in fl and all(_sink(p) is not None for p in params):
        g = fl[key][1]
        for p in params:
            grads[p] = p.grad
        return g
    if len(params) == 1 and _sink(params[0]) is not None:       # a "group" of one (no `down` branch): the parameter's own flat view
        grads[params[0]] = params[0].grad
        return params[0].grad.view(shape)
    g = torch.zeros(shape, dtype=torch.float32, device=params[0].device)
    flat, o = g.view(-1), 0
    for p in params:
        grads[p] = flat[o:o + p.numel()].view(p.shape)
        o += p.numel()
    return g


def bn_uses_batch_stats(bn):
    """nn.BatchNorm semantics (torch/nn/modules/batchnorm.py): batch statistics iff the BatchNorm ITSELF is in training mode or keeps
    no running estimates — the flags of the real child module decide, not the parent unit's (frozen-BN fine-tuning calls
    bn.eval() on individual layers)."""
    return bn.training or bn.running_mean is None


class BNCoef:
    """Per-channel coefficient arrays for a (possibly concatenated) set of BatchNorms over `C` channels.
    `bns`: the BatchNorm modules whose channels the buffer covers (statistics are accumulated iff any of them needs them)."""

    def __init__(self, C, dev, bns):
        training = any(bn_uses_batch_stats(b) for b in bns)
        self.C, self.dev, self.training = C, dev, training
        self.batch = {}                  # (lo, hi) -> this slice normalised with batch statistics
        self.a, self.b = _empty32(C, dev), _empty32(C, dev)
        self.mean, self.invstd = _empty32(C, dev), _empty32(C, dev)
        self.stats = stat_arena.take(2, C, dev) if training else None
        self.jobs = []

    @property
    def ssum(self):
        return self.stats[0] if self.training else None

    @property
    def ssq(self):
        return self.stats[1] if self.training else None

    def add_bn(self, bn, lo, hi, count):
        """forward job for nn.BatchNorm2d `bn` on channels [lo,hi)"""
        use_batch = bn_uses_batch_stats(bn)
        self.batch[(lo, hi)] = use_batch
        sl = slice(lo, hi)
        if use_batch:
            if bn.momentum is None and bn.running_mean is not None:
                raise NotImplementedError("BatchNorm momentum=None (cumulative moving average) is not built; the DS-GCN configs use 0.1")
            self.jobs.append(ops.bn_job(0, hi - lo, sum=self.stats[0, sl], sq=self.stats[1, sl], count=count, gamma=bn.weight,
                                        beta=bn.bias, running_mean=bn.running_mean, running_var=bn.running_var,
                                        save_mean=self.mean[sl], save_invstd=self.invstd[sl], a=self.a[sl], b=self.b[sl],
                                        momentum=bn.momentum if bn.momentum is not None else 0.1, eps=bn.eps))
            if bn.running_mean is not None:
                _Pending.add(bn)
        else:
            self.jobs.append(ops.bn_job(1, hi - lo, gamma=bn.weight, beta=bn.bias, running_mean=bn.running_mean,
                                        running_var=bn.running_var, save_mean=self.mean[sl], save_invstd=self.invstd[sl],
                                        a=self.a[sl], b=self.b[sl], eps=bn.eps))

    def add_identity(self, lo, hi):
        sl = slice(lo, hi)
        self.jobs.append(ops.bn_job(4, hi - lo, a=self.a[sl], b=self.b[sl]))

    def run(self):
        ops.bn_finalize(self.jobs)
        self.jobs = []


class BNBack:
    """Backward coefficient arrays (ca, cb, cc) for the same channel layout as a BNCoef."""

    def __init__(self, fwd):
        C, dev = fwd.C, fwd.dev
        self.fwd = fwd
        self.ca, self.cb, self.cc = _empty32(C, dev), _empty32(C, dev), _empty32(C, dev)
        self.stats = stat_arena.take(2, C, dev)
        self.jobs = []

    @property
    def ssum(self):
        return self.stats[0]

    @property
    def ssq(self):
        return self.stats[1]

    def add_bn(self, bn, lo, hi, count, grads):
        sl = slice(lo, hi)
        dg, db = grad_like(bn.weight, zero=False), grad_like(bn.bias, zero=False)
        grads[bn.weight], grads[bn.bias] = dg, db
        self.jobs.append(ops.bn_job(2 if self.fwd.batch[(lo, hi)] else 3, hi - lo, sum=self.stats[0, sl], sq=self.stats[1, sl], count=count,
                                    gamma=bn.weight, save_mean=self.fwd.mean[sl], save_invstd=self.fwd.invstd[sl],
                                    a=self.ca[sl], b=self.cb[sl], c=self.cc[sl], dgamma=dg, dbeta=db))

    def add_identity(self, lo, hi):
        sl = slice(lo, hi)
        self.jobs.append(ops.bn_job(4, hi - lo, a=self.ca[sl], b=self.cb[sl], c=self.cc[sl]))

    def run(self):
        ops.bn_finalize(self.jobs)
        self.jobs = []

    def dy(self, e, y, lo=None, hi=None):
        """activation source for dy = ca*e + cb*y + cc on channels [lo,hi) (e, y already sliced)"""
        sl = slice(lo, hi)
        return Act(e, self.ca[sl], self.cc[sl], y, self.cb[sl])


# ------------------------------------------------------------------------------------------------
# dgphgcn1  (spatial unit with the dynamic semantic adjacency)
# ------------------------------------------------------------------------------------------------

def _topo_params(m):
    """weights / biases of the topology-feature convolutions in the order the kernels read H"""
    if getattr(m, "plain", False):
        return [m.conv1.weight, m.conv2.weight], [m.conv1.bias, m.conv2.bias]
    return [m.conv1.weight, m.conv2.weight, m.conv1_se.weight], [m.conv1.bias, m.conv2.bias, m.conv1_se.bias]


def dgphgcn1_forward(m, x, n, T, V, save):
    """x [n*T*V, C_in] -> out [n*T*V, C_out].  `save`: dict filled for backward (or None)."""
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, R = m.in_channels, m.out_channels, m.mid_channels
    KC = 3 * R
    has_down = m.has_down
    training = m.training
    nt, et = m._tables(dev)     # (None, None) for the plain flag set

    # ---- topology branch: temporal mean -> 9R features per joint -> per-sample adjacency
    # (bf16 mode: the three feature convolutions run on the tensor core from a bf16 copy of the temporal mean, with the
    #  accumulator stored UNROUNDED in fp32 — H feeds differences, tanh and softmax)
    tc_topo = dt == torch.bfloat16 and Cin % 8 == 0 and R % 8 == 0 and ops.L.is_device_build()
    plain = getattr(m, "plain", False)       # DG-GCN flag set (dggcn; dghgcn / dgphgcn / dgphgcn1 with the attention flags off)
    tw, tb = _topo_params(m)
    HC = (6 if plain else 9) * R
    Wt = cat_params(m, "Wt", tw, (HC, Cin))
    bt = cat_params(m, "bt", tb, (HC,))
    H = torch.empty(n * V, HC, dtype=torch.float32, device=dev)
    adyn = torch.empty(n, V, V, KC, dtype=dt, device=dev)
    S = torch.empty(n, 3, V, V, dtype=torch.float32, device=dev)
    xm = torch.empty(n, V, Cin, dtype=torch.float32, device=dev)
    xmb = torch.empty(n, V, Cin, dtype=torch.bfloat16, device=dev) if tc_topo else None
    We, be = (None, None) if plain else (m.edge_linears.weight.view(15 * R, R), m.edge_linears.bias)
    # The topology chain (three small, latency-bound launches) is independent of the pre / down GEMM: it runs on the side stream
    # next to it (fork / join edges under graph capture) and is joined in front of the contraction that reads adyn.
    with (ops.L.side_stream() if TOPO_SIDE else contextlib.nullcontext()):
        ops.L.keepalive.extend((x, Wt, bt, H, adyn, S, xm, xmb))
        if tc_topo:
            ops.tmean(x, n, T, V, with_bf16=True, out=(xm, xmb))                # [n,V,Cin] fp32 + bf16
            xm2 = xmb.view(n * V, Cin)
            ops.conv_gemm(xm2, Wt, HC, H, n_samples=n, T_in=1, T_out=1, Vin=V, bias=bt, out_f32=True)
        else:
            ops.tmean(x, n, T, V, out=(xm,))                                    # [n,V,Cin] fp32
            xm2 = xm.view(n * V, Cin)
            ops.conv_gemm(xm2, Wt, HC, H, n_samples=n, T_in=1, T_out=1, Vin=V, bias=bt)
        ops.topology_fwd(H, n, V, R, nt, et, m.A, m.alpha, m.beta, We, be, adyn, S, plain=plain, subset_wise=bool(m.subset_wise))

    # ---- pre (+down) 1x1 convolutions in one GEMM, BatchNorm statistics in the epilogue
    Npd = KC + (Cout if has_down else 0)
    if has_down:
        Wpd = cat_params(m, "Wpd", [m.pre[0].weight, m.down[0].weight], (Npd, Cin))
        bpd = cat_params(m, "bpd", [m.pre[0].bias, m.down[0].bias], (Npd,))
    else:
        Wpd, bpd = m.pre[0].weight.view(KC, Cin), m.pre[0].bias
    PD = torch.empty(rows, Npd, dtype=dt, device=dev)
    c_pd = BNCoef(Npd, dev, [m.pre[1]] + ([m.down[1]] if has_down else []))
    ops.conv_gemm(x, Wpd, Npd, PD, n_samples=n, T_in=T, T_out=T, Vin=V, bias=bpd, stat_sum=c_pd.ssum, stat_sq=c_pd.ssq)
    c_pd.add_bn(m.pre[1], 0, KC, rows)
    if has_down:
        c_pd.add_bn(m.down[1], KC, Npd, rows)
    c_pd.run()
    if TOPO_SIDE:
        ops.L.join_side()

    # ---- y[n,t,w,kc] = sum_u relu(bn(pre))[n,t,u,kc] * adyn[n,u,w,kc];  z = post(y), BatchNorm statistics in the epilogue
    P_act = Act(PD[:, :KC], c_pd.a[:KC], c_pd.b[:KC], relu=True)
    Z = torch.empty(rows, Cout, dtype=dt, device=dev)
    c_z = BNCoef(Cout, dev, [m.bn])
    Wpost = m.post.weight.view(Cout, KC)
    if FUSED_AGG and dt == torch.bfloat16 and KC <= 64 and KC % 8 == 0 and Cout <= 128 and ops.L.is_device_build():
        # north-star kernel (a): the contraction runs INSIDE the post GEMM (operand producer of the tensor core); Y is written only
        # when backward will need it (weight gradient of `post`) and is never read back in the forward pass
        Y = torch.empty(rows, KC, dtype=dt, device=dev) if save is not None else None
        ops.conv_gemm(P_act, Wpost, Cout, Z, n_samples=n, T_in=T, T_out=T, Vin=V, bias=m.post.bias, stat_sum=c_z.ssum, stat_sq=c_z.ssq,
                      adyn=adyn, y_out=Y)
    else:
        Y = torch.empty(rows, KC, dtype=dt, device=dev)
        ops.graph_agg(P_act, Y, mode=0, n_samples=n, T=T, V=V, KC=KC, adyn=adyn)
        ops.conv_gemm(Y, Wpost, Cout, Z, n_samples=n, T_in=T, T_out=T, Vin=V, bias=m.post.bias, stat_sum=c_z.ssum, stat_sq=c_z.ssq)
    c_z.add_bn(m.bn, 0, Cout, rows)
    c_z.run()

    # ---- out = relu(bn(z) + down(x))
    out = torch.empty(rows, Cout, dtype=dt, device=dev)
    if has_down:
        src = Act(Z, c_z.a, c_z.b, PD[:, KC:], c_pd.a[KC:], c_pd.b[KC:], relu=True)
    else:
        src = Act(Z, c_z.a, c_z.b, x, relu=True)
    ops.pointwise(src, out)
    if save is not None:
        save.update(x=x, xm2=xm2, Wt=Wt, H=H, adyn=adyn, S=S, Wpd=Wpd, PD=PD, c_pd=c_pd, Y=Y, Z=Z, c_z=c_z, out=out,
                    dims=(n, T, V))
    return out


def dgphgcn1_backward_tail(m, sv):
    """What the producer of `dout` needs to apply this unit's output ReLU mask and accumulate the BN-backward sums of `bn` in its
    own epilogue (mstcn_backward(tail=...)): destination buffer, mask tensor, partner, statistics.  Pass the result back as
    dgphgcn1_backward(..., pre=...)."""
    n, T, V = sv["dims"]
    x, out, Z, c_z = sv["x"], sv["out"], sv["Z"], sv["c_z"]
    rows = n * T * V
    Cout, KC = m.out_channels, 3 * m.mid_channels
    E = torch.empty(rows, KC + Cout, dtype=x.dtype, device=x.device) if m.has_down else None
    E4 = E[:, KC:] if m.has_down else torch.empty(rows, Cout, dtype=x.dtype, device=x.device)
    b_z = BNBack(c_z)
    return dict(out=E4, mask=out, partner=Z, stat_sum=b_z.ssum, stat_sq=b_z.ssq, E=E, b_z=b_z)


def dgphgcn1_backward(m, sv, dout, grads, extra_add=None, pre=None):
    """dout: gradient w.r.t. the unit output [rows, C_out].  Fills `grads` {param: grad}; returns dx.
    `extra_add` (optional, [rows, C_in]) is added to dx inside the last kernel's epilogue.
    `pre`: the dict of dgphgcn1_backward_tail when the producer of dout already wrote e4 = dout * [out > 0] and the sums."""
    n, T, V = sv["dims"]
    x, PD, Y, Z, out, adyn = sv["x"], sv["PD"], sv["Y"], sv["Z"], sv["out"], sv["adyn"]
    c_pd, c_z = sv["c_pd"], sv["c_z"]
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, R = m.in_channels, m.out_channels, m.mid_channels
    KC = 3 * R
    has_down = m.has_down
    Npd = KC + (Cout if has_down else 0)
    nt, et = m._tables(dev)

    # ---- e4 = dout * [out > 0], BatchNorm-backward sums for `bn` (and `down.1`)
    b_pd = BNBack(c_pd)
    if pre is not None:
        E, E4, b_z = pre["E"], pre["out"], pre["b_z"]
    else:
        E = torch.empty(rows, Npd, dtype=dt, device=dev) if has_down else None
        E4 = E[:, KC:] if has_down else torch.empty(rows, Cout, dtype=dt, device=dev)
        b_z = BNBack(c_z)
        ops.pointwise(dout, E4, mask=out, stat_sum=b_z.ssum, stat_sq=b_z.ssq, partner=Z)
    E5 = E[:, :KC] if has_down else torch.empty(rows, KC, dtype=dt, device=dev)
    b_z.add_bn(m.bn, 0, Cout, rows, grads)
    if has_down:
        ops.pointwise(E4, None, stat_sum=b_pd.ssum[KC:], stat_sq=b_pd.ssq[KC:], partner=PD[:, KC:])
    b_z.run()
    dZ = b_z.dy(E4, Z)

    # every parameter-gradient accumulator of the unit: views of the flat gradient buffer when the model is packed
    # (GradBuckets), else fresh zeros
    dWpost, dbpost, dA, dal, dbe = (grad_like(q) for q in (m.post.weight, m.post.bias, m.A, m.alpha, m.beta))
    plain = getattr(m, "plain", False)
    dWe, dbe_l = (None, None) if plain else (grad_like(m.edge_linears.weight), grad_like(m.edge_linears.bias))
    tw, tb = _topo_params(m)
    HC = (6 if plain else 9) * R
    dWt = grad_cat(m, "Wt", tw, (HC, Cin), grads)
    dbt = grad_cat(m, "bt", tb, (HC,), grads)
    pd_w = [m.pre[0].weight] + ([m.down[0].weight] if has_down else [])
    pd_b = [m.pre[0].bias] + ([m.down[0].bias] if has_down else [])
    dWpd, dbpd = grad_cat(m, "Wpd", pd_w, (Npd, Cin), grads), grad_cat(m, "bpd", pd_b, (Npd,), grads)

    # ---- post conv backward
    dY = torch.empty(rows, KC, dtype=dt, device=dev)
    Wpost = m.post.weight.view(Cout, KC)
    ops.conv_gemm(dZ, Wpost, KC, dY, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, KC, 0))
    ops.conv_wgrad(Y, dZ, dWpost, db=dbpost, n_samples=n, T_in=T, T_out=T, Vin=V)
    grads[m.post.weight], grads[m.post.bias] = dWpost, dbpost

    # ---- adjacency contraction backward: dadyn, and e5 = dP * [P > 0] with BN-backward sums for pre.1
    P_act = Act(PD[:, :KC], c_pd.a[:KC], c_pd.b[:KC], relu=True)
    dadyn = torch.empty(n, V, V, KC, dtype=torch.float32, device=dev)
    ops.graph_agg_dadj(P_act, dY, dadyn, n_samples=n, T=T, V=V, KC=KC)
    ops.graph_agg(dY, E5, mode=1, n_samples=n, T=T, V=V, KC=KC, adyn=adyn, mask=Act(PD[:, :KC], c_pd.a[:KC], c_pd.b[:KC]),
                  stat_sum=b_pd.ssum[:KC], stat_sq=b_pd.ssq[:KC], partner=PD[:, :KC])
    b_pd.add_bn(m.pre[1], 0, KC, rows, grads)
    if has_down:
        b_pd.add_bn(m.down[1], KC, Npd, rows, grads)
    b_pd.run()

    # ---- topology backward
    H, S, Wt, xm2 = sv["H"], sv["S"], sv["Wt"], sv["xm2"]
    dH = torch.empty_like(H)
    tc_topo = xm2.dtype == torch.bfloat16
    # bf16 compute mode: the kernel writes a bf16 copy of dH (operand of the tensor-core GEMMs when C_in % 8 == 0) and — keyed on that
    # buffer — uses the forward kernel's hardware tanh and the per-node-type form; block 0 (C_in = 3) asks for the copy too, so its
    # backward matches its forward
    dHb = torch.empty(H.shape, dtype=torch.bfloat16, device=dev) if (tc_topo or adyn.dtype == torch.bfloat16) else None
    We, be_l = (None, None) if plain else (m.edge_linears.weight.view(15 * R, R), m.edge_linears.bias)
    ops.topology_bwd(H, n, V, R, nt, et, m.A, m.alpha, m.beta, We, be_l, S,
                     dadyn, dH, dA, dal, dbe, dWe, dbe_l, dH_bf16=dHb, plain=plain, subset_wise=bool(m.subset_wise))
    grads[m.A], grads[m.alpha], grads[m.beta] = dA, dal, dbe
    if not plain:
        grads[m.edge_linears.weight], grads[m.edge_linears.bias] = dWe, dbe_l
    dxm = torch.empty(n * V, Cin, dtype=torch.float32, device=dev)
    if tc_topo:
        ops.conv_wgrad(xm2, dHb, dWt, db=dbt, n_samples=n, T_in=1, T_out=1, Vin=V)
        ops.conv_gemm(dHb, Wt, Cin, dxm, n_samples=n, T_in=1, T_out=1, Vin=V, ws=(1, Cin, 0), out_f32=True)
    else:
        ops.conv_wgrad(xm2, dH, dWt, db=dbt, n_samples=n, T_in=1, T_out=1, Vin=V)
        ops.conv_gemm(dH, Wt, Cin, dxm, n_samples=n, T_in=1, T_out=1, Vin=V, ws=(1, Cin, 0))

    # ---- dx = [dP_raw | dD_raw] @ [Wpre; Wdown] (+ e4 when the residual is the identity) + dxm/T (+ extra)
    Wpd = sv["Wpd"]
    dx = torch.empty(rows, Cin, dtype=dt, device=dev)
    if has_down:
        dPD = b_pd.dy(E, PD)
        ops.conv_gemm(dPD, Wpd, Cin, dx, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0), add=extra_add, bcast=dxm,
                      bcast_scale=1.0 / T)
    else:
        dPD = b_pd.dy(E5, PD)
        ops.conv_gemm(dPD, Wpd, Cin, dx, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0), add=E4, add2=extra_add, bcast=dxm,
                      bcast_scale=1.0 / T)
    ops.conv_wgrad(x, dPD, dWpd, db=dbpd, n_samples=n, T_in=T, T_out=T, Vin=V)
    return dx


# ------------------------------------------------------------------------------------------------
# mstcn / dgmstcn  (multi-scale temporal unit)
# ------------------------------------------------------------------------------------------------

def ms_layout(m):
    """channel ranges of the branches: [(kind, lo, hi, cfg)], kind in {'conv','max','1x1'}"""
    out, lo = [], 0
    widths = getattr(m, "branch_widths", None)       # MSTCN (msg3d_utils.py:75-76): the LAST branch takes the remainder
    for j, cfg in enumerate(m.ms_cfg):
        w = widths[j] if widths is not None else (m.rem_mid_channels if j == 0 else m.mid_channels)
        kind = "1x1" if cfg == "1x1" else ("max" if cfg[0] == "max" else "conv")
        out.append((kind, lo, lo + w, cfg))
        lo += w
    return out, lo


def _ms_ranges(layout):
    def span(kind):
        items = [(lo, hi) for k, lo, hi, _ in layout if k == kind]
        if not items:
            return (0, 0)
        for (l0, h0), (l1, h1) in zip(items, items[1:]):
            if h0 != l1:
                raise NotImplementedError("ms_cfg: branches of one kind must be adjacent")
        return (items[0][0], items[-1][1])
    if sum(1 for k, *_ in layout if k == "max") > 1 or sum(1 for k, *_ in layout if k == "1x1") > 1:
        raise NotImplementedError("ms_cfg: at most one 'max' and one '1x1' branch")
    return (span("conv"), span("max"), span("1x1"))


def _ms_fused_args(m, layout, b_act, n, T, T_out, s, V, has_ext, grads):
    """Argument block of the fused tcgen05 branch-stage kernels, or None when the shape is not taken by them."""
    if b_act.dtype != torch.bfloat16 or not ops.L.is_device_build() or getattr(m, "no_transform", False):
        return None
    weights = {}
    for j, (kind, lo, hi, cfg) in enumerate(layout):
        if kind == "conv":
            conv = m.branches[j][3].conv
            dW = db = None
            if grads is not None:
                dW, db = grad_like(conv.weight), grad_like(conv.bias)
                grads[conv.weight], grads[conv.bias] = dW, db
            weights[j] = (conv.weight, conv.bias, dW, db)
    a = ops.ms_temporal_args(b_act, layout, weights, n=n, T_in=T, T_out=T_out, stride=s, V=V, has_ext=has_ext,
                             add_coeff=m.add_coeff if has_ext else None)
    a.wgrads_py = {j: (w[2], w[3]) for j, w in weights.items()}      # the accumulators, for the TMA-fed weight gradient
    return a if ops.ms_temporal_supported(a) else None


TOPO_SIDE = os.environ.get("DSG_TOPO_SIDE", "1") != "0"       # 0: topology chain of the forward pass on the main stream
FUSED_AGG = os.environ.get("DSG_FUSED_AGG", "1") != "0"     # 0: separate adjacency contraction (dsg_graph_agg) + post GEMM
MS_TWGRAD = os.environ.get("DSG_MS_TWGRAD", "1") != "0"   # 0: round-1 temporal weight-gradient kernel (CUDA-core staging)
MS_TAP = os.environ.get("DSG_MS_TAP", "1") != "0"      # 0: the staged single-kernel branch stage (ms_temporal_tc.cuh) instead


def _ms_tap_path(m, layout, ranges, dt):
    """(conv channel width rounded up to 8, {branch: (W, bias)}) when the tap-shifted TMA path applies, else None."""
    if not MS_TAP or dt != torch.bfloat16 or not ops.L.is_device_build() or getattr(m, "no_transform", False):
        return None
    (clo, chi), _, _ = ranges
    if chi <= clo or clo != 0 or m.stride > 2:
        return None
    weights = {}
    for j, (kind, lo, hi, cfg) in enumerate(layout):
        if kind == "conv":
            if cfg[0] != 3:
                return None
            conv = m.branches[j][3].conv
            weights[j] = (conv.weight, conv.bias)
    return (chi + 7) & ~7, weights


def mstcn_forward(m, g, n, T, V, save, res=None, final_relu=False):
    """g [n*T*V, C_in] -> [n*T_out*V, C_out] = bn(transform(branches(g))) (+ res) (relu).
    `res`: optional Act-like tuple (x, a, b) added before the final ReLU (DGBlock residual)."""
    dev, dt = g.device, g.dtype
    training = m.training
    has_ext = m.has_ext
    Vp, s = V + int(has_ext), m.stride
    T_out = (T - 1) // s + 1
    Cin, Cout = m.in_channels, m.out_channels
    layout, Ct = ms_layout(m)
    ranges = _ms_ranges(layout)

    # ---- all branch 1x1 convolutions as one GEMM over the (V+1)-joint tensor
    convs = [m.branches[j] if kind == "1x1" and not isinstance(m.branches[j], torch.nn.Sequential) else m.branches[j][0]
             for j, (kind, *_) in enumerate(layout)]
    Wbr = cat_params(m, "Wbr", [c.weight for c in convs], (Ct, Cin))
    bbr = cat_params(m, "bbr", [c.bias for c in convs], (Ct,))
    rows_b = n * T * Vp
    B = torch.empty(rows_b, Ct, dtype=dt, device=dev)
    c_b = BNCoef(Ct, dev, [m.branches[j][1] for j, (kind, *_) in enumerate(layout) if kind != "1x1"])
    ops.conv_gemm(g, Wbr, Ct, B, n_samples=n, T_in=T, T_out=T, Vin=V, ext_in=has_ext, bias=bbr, stat_sum=c_b.ssum, stat_sq=c_b.ssq)
    for j, (kind, lo, hi, _) in enumerate(layout):
        if kind == "1x1":
            c_b.add_identity(lo, hi)
        else:
            c_b.add_bn(m.branches[j][1], lo, hi, rows_b)
    c_b.run()

    rows_o, rows_f = n * T_out * Vp, n * T_out * V
    feat = torch.empty(rows_f, Ct, dtype=dt, device=dev)
    oglob = torch.empty(n * T_out, Ct, dtype=torch.float32, device=dev) if has_ext else None
    no_tr = getattr(m, "no_transform", False)
    feat_bns = m._feat_bns(layout) if no_tr else [(m.transform[0], 0, Ct)]
    c_t = BNCoef(Ct, dev, [b for b, _, _ in feat_bns])
    add_coeff = m.add_coeff if has_ext else None
    if has_ext and add_coeff.numel() < V:
        raise ValueError("add_coeff is shorter than the number of joints")
    H = None
    tap = _ms_tap_path(m, layout, ranges, dt)
    if tap is not None:
        # ---- tap-shifted TMA path: relu(bn(B)) of the conv channels is materialised once (the 4-D tensor map's zero fill then
        #      IS the convolution's zero padding), all conv branches run as one tcgen05 launch, and a streaming pass adds the
        #      max-pool / pass-through branches, local + global * add_coeff and the statistics of transform.0
        chw, weights = tap
        H = torch.empty(rows_b, chw, dtype=dt, device=dev)
        ops.pointwise(Act(B[:, :chw], c_b.a[:chw], c_b.b[:chw], relu=True), H)
        O = torch.empty(rows_o, chw, dtype=dt, device=dev)
        if ops.ms_conv(H, O, layout, weights, n=n, T_in=T, T_out=T_out, stride=s, Vr=Vp, transposed=False):
            ops.ms_combine_fwd(Act(B, c_b.a, c_b.b), O, feat, oglob, n=n, T_in=T, T_out=T_out, stride=s, V=V, has_ext=has_ext,
                               ranges=ranges, add_coeff=add_coeff, stat_sum=c_t.ssum, stat_sq=c_t.ssq)
        else:
            tap = H = None
    fused = _ms_fused_args(m, layout, Act(B, c_b.a, c_b.b), n, T, T_out, s, V, has_ext, None) if tap is None else None
    if tap is not None:
        pass
    elif fused is not None:
        # ---- one tcgen05 kernel: dilated convs (implicit GEMM), max-pool, pass-through, local + global*add_coeff
        ops.ms_temporal_fwd(fused, feat, oglob, c_t.ssum, c_t.ssq)
    else:
        # ---- per-branch path: dilated (k x 1) convolutions of the conv branches ...
        O = torch.empty(rows_o, Ct, dtype=dt, device=dev)
        for j, (kind, lo, hi, cfg) in enumerate(layout):
            if kind != "conv":
                continue
            k, d = cfg
            pad = (k + (k - 1) * (d - 1) - 1) // 2
            conv = m.branches[j][3].conv
            ops.conv_gemm(Act(B[:, lo:hi], c_b.a[lo:hi], c_b.b[lo:hi], relu=True), conv.weight, hi - lo, O[:, lo:hi],
                          n_samples=n, T_in=T, T_out=T_out, Vin=Vp, bias=conv.bias, taps=k, tap_step=d, tap_off=-pad, t_mul=s)
        # ---- ... then max-pool / pass-through branches, local + global*add_coeff, statistics for transform.0
        ops.ms_combine_fwd(Act(B, c_b.a, c_b.b), O, feat, oglob, n=n, T_in=T, T_out=T_out, stride=s, V=V, has_ext=has_ext,
                           ranges=ranges, add_coeff=add_coeff, stat_sum=c_t.ssum, stat_sq=c_t.ssq)
    for bn, lo, hi in feat_bns:
        c_t.add_bn(bn, lo, hi, rows_f)
    c_t.run()
    if no_tr:
        # ---- MSTCN (msg3d_utils.py:135-147): every branch ends in its own BatchNorm; out = relu(cat + residual), no transform conv.
        #      Inside a block (ctrgcn.py:56-58) the block residual and a second ReLU follow.
        inner = torch.empty(rows_f, Ct, dtype=dt, device=dev)
        own = m._own_residual(g, n, T, V, save)
        if own is not None:
            ops.pointwise(Act(feat, c_t.a, c_t.b, own[0], own[1], own[2], relu=True), inner)
        else:
            ops.pointwise(Act(feat, c_t.a, c_t.b, relu=True), inner)
        out = inner
        if res is not None or final_relu:
            out = torch.empty(rows_f, Ct, dtype=dt, device=dev)
            if res is not None:
                rx, ra, rb = res
                ops.pointwise(Act(inner, None, None, rx, ra, rb, relu=final_relu), out)
            else:
                ops.pointwise(Act(inner, relu=final_relu), out)
        if save is not None:
            save.update(g=g, Wbr=Wbr, B=B, c_b=c_b, feat=feat, oglob=oglob, c_t=c_t, inner=inner, out=out, dims=(n, T, V),
                        final_relu=final_relu, has_outer=out is not inner, H=None, feat_bns=feat_bns)
        return out, T_out

    # ---- transform conv + final BatchNorm (+ residual, ReLU)
    U = torch.empty(rows_f, Cout, dtype=dt, device=dev)
    c_u = BNCoef(Cout, dev, [m.bn])
    ops.conv_gemm(Act(feat, c_t.a, c_t.b, relu=True), m.transform[2].weight.view(Cout, Ct), Cout, U, n_samples=n, T_in=T_out,
                  T_out=T_out, Vin=V, bias=m.transform[2].bias, stat_sum=c_u.ssum, stat_sq=c_u.ssq)
    c_u.add_bn(m.bn, 0, Cout, rows_f)
    c_u.run()
    out = torch.empty(rows_f, Cout, dtype=dt, device=dev)
    if res is not None:
        rx, ra, rb = res
        src = Act(U, c_u.a, c_u.b, rx, ra, rb, relu=final_relu)
    else:
        src = Act(U, c_u.a, c_u.b, relu=final_relu)
    ops.pointwise(src, out)
    if save is not None:
        save.update(g=g, Wbr=Wbr, B=B, c_b=c_b, feat=feat, oglob=oglob, c_t=c_t, U=U, c_u=c_u, out=out, dims=(n, T, V),
                    final_relu=final_relu, H=H)
    return out, T_out


def _ms_branch_backward(m, sv, dfeat, grads, tail=None):
    """Backward of the branch stage shared by mstcn / dgmstcn / MSTCN: from dfeat (gradient w.r.t. the concatenated raw branch
    outputs, an Act) to dg (gradient w.r.t. the unit input); fills the branch parameter gradients."""
    n, T, V = sv["dims"]
    g, B, c_b = sv["g"], sv["B"], sv["c_b"]
    dev, dt = g.device, g.dtype
    has_ext = m.has_ext
    Vp, s = V + int(has_ext), m.stride
    T_out = (T - 1) // s + 1
    Cin = m.in_channels
    layout, Ct = ms_layout(m)
    ranges = _ms_ranges(layout)
    rows_b, rows_o = n * T * Vp, n * T_out * Vp
    convs = [m.branches[j] if kind == "1x1" and not isinstance(m.branches[j], torch.nn.Sequential) else m.branches[j][0]
             for j, (kind, *_) in enumerate(layout)]
    dWbr = grad_cat(m, "Wbr", [c.weight for c in convs], (Ct, Cin), grads)
    dbbr = grad_cat(m, "bbr", [c.bias for c in convs], (Ct,), grads)
    b_b = BNBack(c_b)
    E3 = torch.empty(rows_b, Ct, dtype=dt, device=dev)
    dadd = grad_like(m.add_coeff) if has_ext else None
    if has_ext:
        grads[m.add_coeff] = dadd
    fused = _ms_fused_args(m, layout, Act(B, c_b.a, c_b.b), n, T, T_out, s, V, has_ext, grads)
    H = sv.get("H")
    tap_done = False
    if H is not None and fused is not None:
        # ---- tap-shifted TMA path: a streaming pass routes dfeat to the branch outputs (d_o of the conv range incl. the
        #      joint-mean row, masked gradients of the max / pass ranges, dadd_coeff), then every conv branch's data gradient is
        #      one tcgen05 launch with the ReLU mask (H > 0) and the BN-backward sums in its epilogue
        chw = H.shape[1]
        tap = _ms_tap_path(m, layout, ranges, dt)
        d_o = torch.empty(rows_o, Ct, dtype=dt, device=dev)           # gradient w.r.t. every branch output, joint-mean row included
        ckw = dict(n=n, T_in=T, T_out=T_out, stride=s, V=V, has_ext=has_ext, ranges=ranges, add_coeff=m.add_coeff if has_ext else None,
                   e_sum=b_b.ssum, e_sq=b_b.ssq, dadd_coeff=dadd, d_o_full=True)
        ops.ms_combine_bwd(Act(B, c_b.a, c_b.b), dfeat, d_o, E3, sv["oglob"], B, parts=1, **ckw)
        tap_done = tap is not None and ops.ms_conv(d_o, E3, layout, tap[1], n=n, T_in=T, T_out=T_out, stride=s, Vr=Vp, transposed=True,
                                                   mask=H, partner=B, stat_sum=b_b.ssum, stat_sq=b_b.ssq)
        if not tap_done:
            raise RuntimeError("dsg_ms_conv took the forward of this shape but declined its data gradient")
        # (after the conv branches: dsg_ms_conv pads its last 16-byte channel chunk with zeros, the max range starts inside it)
        ops.ms_combine_bwd(Act(B, c_b.a, c_b.b), dfeat, d_o, E3, sv["oglob"], B, parts=2, **ckw)
        # weight gradients of the conv branches: TMA-fed tcgen05 engine straight from H and d_o (both materialised above), else the
        # round-1 kernel that stages its operands with CUDA cores
        if not (MS_TWGRAD and ops.ms_conv_wgrad(H, d_o, layout, fused.wgrads_py, n=n, T_in=T, T_out=T_out, stride=s, Vr=Vp)):
            ops.ms_temporal_bwd(fused, dfeat, E3, sv["oglob"], b_b.ssum, b_b.ssq, dadd, data=False)
    elif fused is not None:
        # ---- two tcgen05 kernels: data gradient of the whole branch stage (+ masks, BN-backward sums, dadd_coeff),
        #      then the weight gradients of the dilated convolutions
        ops.ms_temporal_bwd(fused, dfeat, E3, sv["oglob"], b_b.ssum, b_b.ssq, dadd)
    else:
        # ---- per-branch path: d_o for the conv branches, masked grads of max / pass branches, dadd_coeff
        d_o = torch.empty(rows_o, Ct, dtype=dt, device=dev)
        ops.ms_combine_bwd(Act(B, c_b.a, c_b.b), dfeat, d_o, E3, sv["oglob"], B, n=n, T_in=T, T_out=T_out, stride=s, V=V,
                           has_ext=has_ext, ranges=ranges, add_coeff=m.add_coeff if has_ext else None, e_sum=b_b.ssum, e_sq=b_b.ssq,
                           dadd_coeff=dadd)
        for j, (kind, lo, hi, cfg) in enumerate(layout):
            if kind != "conv":
                continue
            k, d = cfg
            pad = (k + (k - 1) * (d - 1) - 1) // 2
            w = hi - lo
            conv = m.branches[j][3].conv
            ops.conv_gemm(d_o[:, lo:hi], conv.weight, w, E3[:, lo:hi], n_samples=n, T_in=T_out, T_out=T, Vin=Vp, ws=(k, w * k, 1), taps=k,
                          tap_step=-d, tap_off=pad, t_div=s, mask=Act(B[:, lo:hi], c_b.a[lo:hi], c_b.b[lo:hi]),
                          stat_sum=b_b.ssum[lo:hi], stat_sq=b_b.ssq[lo:hi], partner=B[:, lo:hi])
            dWc, dbc = grad_like(conv.weight), grad_like(conv.bias)
            ops.conv_wgrad(Act(B[:, lo:hi], c_b.a[lo:hi], c_b.b[lo:hi], relu=True), d_o[:, lo:hi], dWc, db=dbc, n_samples=n, T_in=T,
                           T_out=T_out, Vin=Vp, taps=k, tap_step=d, tap_off=-pad, t_mul=s)
            grads[conv.weight], grads[conv.bias] = dWc, dbc
    for j, (kind, lo, hi, _) in enumerate(layout):
        if kind == "1x1":
            b_b.add_identity(lo, hi)
        else:
            b_b.add_bn(m.branches[j][1], lo, hi, rows_b, grads)
    b_b.run()
    dB = b_b.dy(E3, B)

    # ---- branch 1x1 convolutions backward (the joint-mean column folds back into the V joints)
    if tail is not None:
        dg = tail["out"]
        ops.conv_gemm(dB, sv["Wbr"], Cin, dg, n_samples=n, T_in=T, T_out=T, Vin=Vp, ws=(1, Cin, 0), contract_ext=has_ext,
                      mask=tail["mask"], partner=tail["partner"], stat_sum=tail["stat_sum"], stat_sq=tail["stat_sq"])
    else:
        dg = torch.empty(n * T * V, Cin, dtype=dt, device=dev)
        ops.conv_gemm(dB, sv["Wbr"], Cin, dg, n_samples=n, T_in=T, T_out=T, Vin=Vp, ws=(1, Cin, 0), contract_ext=has_ext)
    ops.conv_wgrad(g, dB, dWbr, db=dbbr, n_samples=n, T_in=T, T_out=T, Vin=V, ext_in=has_ext)
    return dg


def _mstcn_notransform_backward(m, sv, dout, grads, tail=None):
    """MSTCN: out = relu(inner + res_block), inner = relu(bn_branch(feat) + own_res).  Returns (dg, E) like mstcn_backward."""
    n, T, V = sv["dims"]
    feat, c_t, inner, out = sv["feat"], sv["c_t"], sv["inner"], sv["out"]
    dev, dt = feat.device, feat.dtype
    rows_f, Ct = feat.shape
    if sv["has_outer"] and sv["final_relu"]:
        E = torch.empty(rows_f, Ct, dtype=dt, device=dev)
        ops.pointwise(dout, E, mask=out)
    else:
        E = dout
    b_t = BNBack(c_t)
    E2 = torch.empty(rows_f, Ct, dtype=dt, device=dev)
    ops.pointwise(E, E2, mask=inner, stat_sum=b_t.ssum, stat_sq=b_t.ssq, partner=feat)
    for bn, lo, hi in sv["feat_bns"]:
        b_t.add_bn(bn, lo, hi, rows_f, grads)
    b_t.run()
    dg = _ms_branch_backward(m, sv, b_t.dy(E2, feat), grads, tail)
    dg = m._own_residual_backward(sv, E2, dg, grads)
    return dg, E


def mstcn_backward(m, sv, dout, grads, tail=None):
    """Returns (dg, E) where E is the gradient w.r.t. the pre-ReLU sum (what flows into the residual).
    `tail` (optional dict: out, mask, partner, stat_sum, stat_sq): the consumer of dg is the spatial unit of the same block, whose
    backward starts with dg * [g > 0] and the BatchNorm-backward sums of its `bn`; with `tail` the last GEMM here applies that mask
    and accumulates those sums in its epilogue and writes straight into the consumer's buffer (no separate pass over dg)."""
    n, T, V = sv["dims"]
    g, B, feat, U, out = sv["g"], sv["B"], sv["feat"], sv.get("U"), sv["out"]
    c_b, c_t, c_u = sv["c_b"], sv["c_t"], sv.get("c_u")
    dev, dt = g.device, g.dtype
    has_ext = m.has_ext
    Vp, s = V + int(has_ext), m.stride
    T_out = (T - 1) // s + 1
    Cin, Cout = m.in_channels, m.out_channels
    layout, Ct = ms_layout(m)
    ranges = _ms_ranges(layout)
    rows_b, rows_o, rows_f = n * T * Vp, n * T_out * Vp, n * T_out * V

    if getattr(m, "no_transform", False):
        return _mstcn_notransform_backward(m, sv, dout, grads, tail)
    # ---- final ReLU mask + BN-backward sums of `bn`
    b_u = BNBack(c_u)
    if sv["final_relu"]:
        E = torch.empty(rows_f, Cout, dtype=dt, device=dev)
        ops.pointwise(dout, E, mask=out, stat_sum=b_u.ssum, stat_sq=b_u.ssq, partner=U)
    else:
        E = dout
        ops.pointwise(dout, None, stat_sum=b_u.ssum, stat_sq=b_u.ssq, partner=U)
    b_u.add_bn(m.bn, 0, Cout, rows_f, grads)
    b_u.run()
    dU = b_u.dy(E, U)

    # ---- transform conv backward; e2 = dfeat_act * [bn_t(feat) > 0] with sums for transform.0
    tr = m.transform[2]
    b_t = BNBack(c_t)
    E2 = torch.empty(rows_f, Ct, dtype=dt, device=dev)
    ops.conv_gemm(dU, tr.weight.view(Cout, Ct), Ct, E2, n_samples=n, T_in=T_out, T_out=T_out, Vin=V, ws=(1, Ct, 0),
                  mask=Act(feat, c_t.a, c_t.b), stat_sum=b_t.ssum, stat_sq=b_t.ssq, partner=feat)
    dWtr, dbtr = grad_like(tr.weight), grad_like(tr.bias)
    ops.conv_wgrad(Act(feat, c_t.a, c_t.b, relu=True), dU, dWtr, db=dbtr, n_samples=n, T_in=T_out, T_out=T_out, Vin=V)
    grads[tr.weight], grads[tr.bias] = dWtr, dbtr
    b_t.add_bn(m.transform[0], 0, Ct, rows_f, grads)
    b_t.run()
    dfeat = b_t.dy(E2, feat)

    return _ms_branch_backward(m, sv, dfeat, grads, tail), E


# ------------------------------------------------------------------------------------------------
# unit_tcn  ((k x 1) conv + BatchNorm; also the strided 1x1 block residual)
# ------------------------------------------------------------------------------------------------

def unit_tcn_raw_forward(m, x, n, T, V, save):
    """conv only: returns (R_raw, BNCoef or None, T_out).  The BatchNorm is applied by the consumer."""
    dev, dt = x.device, x.dtype
    k, s, d = m.kernel_size, m.stride, m.dilation
    pad = (k + (k - 1) * (d - 1) - 1) // 2
    T_out = (T + 2 * pad - d * (k - 1) - 1) // s + 1
    Cout = m.out_channels
    rows = n * T_out * V
    Rr = torch.empty(rows, Cout, dtype=dt, device=dev)
    has_bn = isinstance(m.bn, torch.nn.modules.batchnorm._BatchNorm)
    c_r = BNCoef(Cout, dev, [m.bn]) if has_bn else None
    ops.conv_gemm(x, m.conv.weight, Cout, Rr, n_samples=n, T_in=T, T_out=T_out, Vin=V, bias=m.conv.bias, taps=k, tap_step=d,
                  tap_off=-pad, t_mul=s, stat_sum=c_r.ssum if has_bn else None, stat_sq=c_r.ssq if has_bn else None)
    if has_bn:
        c_r.add_bn(m.bn, 0, Cout, rows)
        c_r.run()
    if save is not None:
        save.update(x=x, R=Rr, c_r=c_r, dims=(n, T, V), T_out=T_out)
    return Rr, c_r, T_out


def unit_tcn_raw_backward(m, sv, E, grads, e_is_masked_sum_done=False):
    """E: gradient w.r.t. bn(conv(x)).  Returns dx."""
    n, T, V = sv["dims"]
    x, Rr, c_r, T_out = sv["x"], sv["R"], sv["c_r"], sv["T_out"]
    dev, dt = x.device, x.dtype
    k, s, d = m.kernel_size, m.stride, m.dilation
    pad = (k + (k - 1) * (d - 1) - 1) // 2
    Cin, Cout = m.in_channels, m.out_channels
    if c_r is not None:
        b_r = BNBack(c_r)
        ops.pointwise(E, None, stat_sum=b_r.ssum, stat_sq=b_r.ssq, partner=Rr)
        b_r.add_bn(m.bn, 0, Cout, n * T_out * V, grads)
        b_r.run()
        dR = b_r.dy(E, Rr)
    else:
        dR = Act(E)
    dx = torch.empty(n * T * V, Cin, dtype=dt, device=dev)
    ops.conv_gemm(dR, m.conv.weight, Cin, dx, n_samples=n, T_in=T_out, T_out=T, Vin=V, ws=(k, Cin * k, 1), taps=k, tap_step=-d,
                  tap_off=pad, t_div=s)
    dW, db = grad_like(m.conv.weight), grad_like(m.conv.bias)
    ops.conv_wgrad(x, dR, dW, db=db, n_samples=n, T_in=T, T_out=T_out, Vin=V, taps=k, tap_step=d, tap_off=-pad, t_mul=s)
    grads[m.conv.weight], grads[m.conv.bias] = dW, db
    return dx


def unit_tcn_forward(m, x, n, T, V, save, res=None, final_relu=False):
    """out = bn(conv(x)) (+ res) (relu) — unit_tcn as a block's temporal unit or stand-alone."""
    rsave = {} if save is not None else None
    Rr, c_r, T_out = unit_tcn_raw_forward(m, x, n, T, V, rsave)
    a, b = (c_r.a, c_r.b) if c_r is not None else (None, None)
    if res is None and not final_relu and c_r is None:
        out = Rr
    else:
        out = torch.empty_like(Rr)
        if res is not None:
            rx, ra, rb = res
            ops.pointwise(Act(Rr, a, b, rx, ra, rb, relu=final_relu), out)
        else:
            ops.pointwise(Act(Rr, a, b, relu=final_relu), out)
    if save is not None:
        save.update(raw=rsave, out=out, final_relu=final_relu)
    return out, T_out


def unit_tcn_backward(m, sv, dout, grads):
    """Returns (dx, E) like mstcn_backward."""
    if sv["final_relu"]:
        E = torch.empty_like(dout)
        ops.pointwise(dout, E, mask=sv["out"])
    else:
        E = dout
    return unit_tcn_raw_backward(m, sv["raw"], E, grads), E


# ------------------------------------------------------------------------------------------------
# unit_gcn  (static adjacency; ST-GCN / ST-GCN++ spatial unit)
# ------------------------------------------------------------------------------------------------

def unit_gcn_A(m):
    if m.adaptive == "offset":
        return m.A + m.PA
    if m.adaptive == "importance":
        return m.A * m.PA
    return m.A


def unit_gcn_forward(m, x, n, T, V, save):
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, K = m.in_channels, m.out_channels, m.num_subsets
    training = m.training
    A_eff = unit_gcn_A(m).detach().contiguous()
    has_down = m.with_res and m.has_down
    c_d = None
    D = None
    if has_down:
        D = torch.empty(rows, Cout, dtype=dt, device=dev)
        c_d = BNCoef(Cout, dev, [m.down[1]])
        ops.conv_gemm(x, m.down[0].weight, Cout, D, n_samples=n, T_in=T, T_out=T, Vin=V, bias=m.down[0].bias,
                      stat_sum=c_d.ssum, stat_sq=c_d.ssq)
        c_d.add_bn(m.down[1], 0, Cout, rows)
        c_d.run()
    c_z = BNCoef(Cout, dev, [m.bn])
    Z = torch.empty(rows, Cout, dtype=dt, device=dev)
    if m.conv_pos == "pre":
        mid = torch.empty(rows, K * Cout, dtype=dt, device=dev)
        ops.conv_gemm(x, m.conv.weight, K * Cout, mid, n_samples=n, T_in=T, T_out=T, Vin=V, bias=m.conv.bias)
        ops.graph_agg(mid, Z, mode=2, n_samples=n, T=T, V=V, KC=Cout, A=A_eff, Ksub=K, stat_sum=c_z.ssum, stat_sq=c_z.ssq)
    else:
        mid = torch.empty(rows, K * Cin, dtype=dt, device=dev)
        ops.graph_agg(x, mid, mode=3, n_samples=n, T=T, V=V, KC=Cin, A=A_eff.transpose(1, 2).contiguous(), Ksub=K)
        ops.conv_gemm(mid, m.conv.weight, Cout, Z, n_samples=n, T_in=T, T_out=T, Vin=V, bias=m.conv.bias,
                      stat_sum=c_z.ssum, stat_sq=c_z.ssq)
    c_z.add_bn(m.bn, 0, Cout, rows)
    c_z.run()
    out = torch.empty(rows, Cout, dtype=dt, device=dev)
    if has_down:
        src = Act(Z, c_z.a, c_z.b, D, c_d.a, c_d.b, relu=True)
    elif m.with_res:
        src = Act(Z, c_z.a, c_z.b, x, relu=True)
    else:
        src = Act(Z, c_z.a, c_z.b, relu=True)
    ops.pointwise(src, out)
    if save is not None:
        save.update(x=x, D=D, c_d=c_d, mid=mid, Z=Z, c_z=c_z, out=out, A_eff=A_eff, dims=(n, T, V))
    return out


def unit_gcn_backward(m, sv, dout, grads, extra_add=None):
    n, T, V = sv["dims"]
    x, D, c_d, mid, Z, c_z, out, A_eff = sv["x"], sv["D"], sv["c_d"], sv["mid"], sv["Z"], sv["c_z"], sv["out"], sv["A_eff"]
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, K = m.in_channels, m.out_channels, m.num_subsets
    has_down = D is not None
    b_z = BNBack(c_z)
    E = torch.empty(rows, Cout, dtype=dt, device=dev)
    ops.pointwise(dout, E, mask=out, stat_sum=b_z.ssum, stat_sq=b_z.ssq, partner=Z)
    b_z.add_bn(m.bn, 0, Cout, rows, grads)
    b_z.run()
    dZ = b_z.dy(E, Z)
    dA = torch.zeros_like(A_eff)
    dx = torch.empty(rows, Cin, dtype=dt, device=dev)
    dW, db = grad_like(m.conv.weight), grad_like(m.conv.bias)
    add = E if (m.with_res and not has_down) else None
    add2 = extra_add
    if has_down:
        b_d = BNBack(c_d)
        ops.pointwise(E, None, stat_sum=b_d.ssum, stat_sq=b_d.ssq, partner=D)
        b_d.add_bn(m.down[1], 0, Cout, rows, grads)
        b_d.run()
        dD = b_d.dy(E, D)
        dxd = torch.empty(rows, Cin, dtype=dt, device=dev)
        ops.conv_gemm(dD, m.down[0].weight, Cin, dxd, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0))
        dWd, dbd = grad_like(m.down[0].weight), grad_like(m.down[0].bias)
        ops.conv_wgrad(x, dD, dWd, db=dbd, n_samples=n, T_in=T, T_out=T, Vin=V)
        grads[m.down[0].weight], grads[m.down[0].bias] = dWd, dbd
        add = dxd
    if m.conv_pos == "pre":
        # z = sum_k mid_k A_k ; mid = conv(x)
        dmid = torch.empty(rows, K * Cout, dtype=dt, device=dev)
        dZm = torch.empty(rows, Cout, dtype=dt, device=dev)
        ops.pointwise(dZ, dZm)                       # materialise dz once: it feeds two contractions
        ops.graph_agg(dZm, dmid, mode=3, n_samples=n, T=T, V=V, KC=Cout, A=A_eff, Ksub=K)
        ops.graph_agg_dadj(mid, dZm, dA, n_samples=n, T=T, V=V, KC=Cout, is_static=True, Ksub=K)
        ops.conv_gemm(dmid, m.conv.weight, Cin, dx, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0), add=add, add2=add2)
        ops.conv_wgrad(x, dmid, dW, db=db, n_samples=n, T_in=T, T_out=T, Vin=V)
    else:
        # mid_k = x A_k ; z = conv(mid)
        dmid = torch.empty(rows, K * Cin, dtype=dt, device=dev)
        ops.conv_gemm(dZ, m.conv.weight, K * Cin, dmid, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, K * Cin, 0))
        ops.conv_wgrad(mid, dZ, dW, db=db, n_samples=n, T_in=T, T_out=T, Vin=V)
        At = A_eff.transpose(1, 2).contiguous()
        dxa = torch.empty(rows, Cin, dtype=dt, device=dev)
        ops.graph_agg(dmid, dxa, mode=2, n_samples=n, T=T, V=V, KC=Cin, A=At, Ksub=K)
        # dA[k,u,w] = sum x[u,c] dmid[w,kC+c]  == static dadj with p := dmid (as [w, kC+c]) transposed roles
        dAt = torch.zeros_like(A_eff)
        ops.graph_agg_dadj(dmid, x, dAt, n_samples=n, T=T, V=V, KC=Cin, is_static=True, Ksub=K)
        dA = dAt.transpose(1, 2).contiguous()
        for extra in (add, add2):
            if extra is not None:
                nxt = torch.empty_like(dxa)
                ops.pointwise(Act(dxa, x2=extra), nxt)
                dxa = nxt
        dx = dxa
    grads[m.conv.weight], grads[m.conv.bias] = dW, db
    # chain dA_eff into the learnable tensors
    if m.adaptive == "init":
        grads[m.A] = dA
    elif m.adaptive == "offset":
        grads[m.PA] = dA
    elif m.adaptive == "importance":
        grads[m.PA] = dA * m.A.detach()
    return dx

# ------------------------------------------------------------------------------------------------
# unit_ctrgcn  (CTR-GCN spatial unit: three channel-wise topology refinement graph convolutions, gcn.py:634-666, :882-930)
# ------------------------------------------------------------------------------------------------

def _ctr_groups(m):
    """parameter groups the kernels read concatenated (adjacent views when the model is packed by GradBuckets)"""
    cs = list(m.convs)
    g = {"Wt": [c.conv1.weight for c in cs] + [c.conv2.weight for c in cs], "bt": [c.conv1.bias for c in cs] + [c.conv2.bias for c in cs],
         "W4": [c.conv4.weight for c in cs], "b4": [c.conv4.bias for c in cs],
         "W3": [c.conv3.weight for c in cs] + ([m.down[0].weight] if m.has_down else []),
         "b3": [c.conv3.bias for c in cs] + ([m.down[0].bias] if m.has_down else [])}
    return g


def unit_ctrgcn_forward(m, x, n, T, V, save):
    """x [n*T*V, C_in] -> relu(bn(sum_k CTRGC_k(x)) + down(x)) [n*T*V, C_out]"""
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, R = m.in_c, m.out_c, m.rel_channels
    KC = 3 * Cout
    grp = _ctr_groups(m)
    has_down = m.has_down
    # ---- topology: x1 = conv1(mean_t x), x2 = conv2(mean_t x) (the convolution commutes with the mean), refined adjacency per
    #      sample and OUTPUT channel
    Wt, bt = cat_params(m, "Wt", grp["Wt"], (6 * R, Cin)), cat_params(m, "bt", grp["bt"], (6 * R,))
    tc_topo = dt == torch.bfloat16 and Cin % 8 == 0 and R % 8 == 0 and ops.L.is_device_build()
    H = torch.empty(n * V, 6 * R, dtype=torch.float32, device=dev)
    if tc_topo:
        xm, xmb = ops.tmean(x, n, T, V, with_bf16=True)
        xm2 = xmb.view(n * V, Cin)
        ops.conv_gemm(xm2, Wt, 6 * R, H, n_samples=n, T_in=1, T_out=1, Vin=V, bias=bt, out_f32=True)
    else:
        xm2 = ops.tmean(x, n, T, V).view(n * V, Cin)
        ops.conv_gemm(xm2, Wt, 6 * R, H, n_samples=n, T_in=1, T_out=1, Vin=V, bias=bt)
    W4, b4 = cat_params(m, "W4", grp["W4"], (3, Cout, R)), cat_params(m, "b4", grp["b4"], (3, Cout))
    adyn = torch.empty(n, V, V, KC, dtype=dt, device=dev)
    ops.ctr_topology(H, n, V, R, Cout, m.A, m.alpha, W4, b4, adyn=adyn)
    # ---- conv3 of the three subsets (+ down) as one GEMM; BatchNorm statistics of `down` in the epilogue
    N3 = KC + (Cout if has_down else 0)
    W3, b3 = cat_params(m, "W3", grp["W3"], (N3, Cin)), cat_params(m, "b3", grp["b3"], (N3,))
    XD = torch.empty(rows, N3, dtype=dt, device=dev)
    c_d = BNCoef(N3, dev, [m.down[1]]) if has_down else None
    ops.conv_gemm(x, W3, N3, XD, n_samples=n, T_in=T, T_out=T, Vin=V, bias=b3,
                  stat_sum=c_d.ssum if has_down else None, stat_sq=c_d.ssq if has_down else None)
    if has_down:
        c_d.add_identity(0, KC)
        c_d.add_bn(m.down[1], KC, N3, rows)
        c_d.run()
    # ---- y_k[n,t,w,c] = sum_u x3_k[n,t,u,c] adyn[n,u,w,k,c]; z = sum_k y_k (a GEMM with [I|I|I]: exact, statistics in its epilogue)
    Y3 = torch.empty(rows, KC, dtype=dt, device=dev)
    ops.graph_agg(XD[:, :KC], Y3, mode=0, n_samples=n, T=T, V=V, KC=KC, adyn=adyn)
    Z = torch.empty(rows, Cout, dtype=dt, device=dev)
    c_z = BNCoef(Cout, dev, [m.bn])
    Wsum = m._sum_weight(dev)
    ops.conv_gemm(Y3, Wsum, Cout, Z, n_samples=n, T_in=T, T_out=T, Vin=V, stat_sum=c_z.ssum, stat_sq=c_z.ssq)
    c_z.add_bn(m.bn, 0, Cout, rows)
    c_z.run()
    out = torch.empty(rows, Cout, dtype=dt, device=dev)
    if has_down:
        src = Act(Z, c_z.a, c_z.b, XD[:, KC:], c_d.a[KC:], c_d.b[KC:], relu=True)
    else:
        src = Act(Z, c_z.a, c_z.b, x, relu=True)
    ops.pointwise(src, out)
    if save is not None:
        save.update(x=x, xm2=xm2, Wt=Wt, H=H, adyn=adyn, W3=W3, XD=XD, c_d=c_d, Z=Z, c_z=c_z, out=out, W4=W4, b4=b4, dims=(n, T, V))
    return out


def unit_ctrgcn_backward(m, sv, dout, grads, extra_add=None):
    n, T, V = sv["dims"]
    x, XD, Z, out, adyn, c_d, c_z = sv["x"], sv["XD"], sv["Z"], sv["out"], sv["adyn"], sv["c_d"], sv["c_z"]
    dev, dt = x.device, x.dtype
    rows = n * T * V
    Cin, Cout, R = m.in_c, m.out_c, m.rel_channels
    KC = 3 * Cout
    has_down = m.has_down
    N3 = KC + (Cout if has_down else 0)
    grp = _ctr_groups(m)
    # ---- e4 = dout * [out > 0], BatchNorm-backward sums of `bn` (and `down.1`)
    E = torch.empty(rows, N3, dtype=dt, device=dev) if has_down else None
    E4 = E[:, KC:] if has_down else torch.empty(rows, Cout, dtype=dt, device=dev)
    b_z = BNBack(c_z)
    ops.pointwise(dout, E4, mask=out, stat_sum=b_z.ssum, stat_sq=b_z.ssq, partner=Z)
    b_z.add_bn(m.bn, 0, Cout, rows, grads)
    b_z.run()
    if has_down:
        b_d = BNBack(c_d)
        ops.pointwise(E4, None, stat_sum=b_d.ssum[KC:], stat_sq=b_d.ssq[KC:], partner=XD[:, KC:])
    # ---- dz broadcast to the three subsets, contraction backward: dadyn and dx3
    dY3 = torch.empty(rows, KC, dtype=dt, device=dev)
    Wsum = m._sum_weight(dev)
    ops.conv_gemm(b_z.dy(E4, Z), Wsum, KC, dY3, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, KC, 0))
    dadyn = torch.empty(n, V, V, KC, dtype=torch.float32, device=dev)
    ops.graph_agg_dadj(XD[:, :KC], dY3, dadyn, n_samples=n, T=T, V=V, KC=KC)
    E5 = E[:, :KC] if has_down else torch.empty(rows, KC, dtype=dt, device=dev)
    ops.graph_agg(dY3, E5, mode=1, n_samples=n, T=T, V=V, KC=KC, adyn=adyn)
    # ---- topology backward
    H, Wt, xm2 = sv["H"], sv["Wt"], sv["xm2"]
    dA, dal = grad_like(m.A), grad_like(m.alpha)
    dW4, db4 = grad_cat(m, "W4", grp["W4"], (3, Cout, R), grads), grad_cat(m, "b4", grp["b4"], (3, Cout), grads)
    dWt, dbt = grad_cat(m, "Wt", grp["Wt"], (6 * R, Cin), grads), grad_cat(m, "bt", grp["bt"], (6 * R,), grads)
    dW3, db3 = grad_cat(m, "W3", grp["W3"], (N3, Cin), grads), grad_cat(m, "b3", grp["b3"], (N3,), grads)
    dH = torch.empty_like(H)
    tc_topo = xm2.dtype == torch.bfloat16
    dHb = torch.empty(H.shape, dtype=torch.bfloat16, device=dev) if tc_topo else None
    ops.ctr_topology(H, n, V, R, Cout, m.A, m.alpha, sv["W4"], sv["b4"], dadyn=dadyn, dH=dH, dH_bf16=dHb, dA=dA, dalpha=dal, dW4=dW4,
                     db4=db4)
    grads[m.A], grads[m.alpha] = dA, dal
    dxm = torch.empty(n * V, Cin, dtype=torch.float32, device=dev)
    dHs = dHb if tc_topo else dH
    ops.conv_wgrad(xm2, dHs, dWt, db=dbt, n_samples=n, T_in=1, T_out=1, Vin=V)
    ops.conv_gemm(dHs, Wt, Cin, dxm, n_samples=n, T_in=1, T_out=1, Vin=V, ws=(1, Cin, 0), out_f32=tc_topo)
    # ---- dx = [dx3 | dD_raw] @ [W3; Wdown] (+ e4 when the residual is the identity) + dxm / T (+ extra)
    dx = torch.empty(rows, Cin, dtype=dt, device=dev)
    if has_down:
        b_d.add_identity(0, KC)
        b_d.add_bn(m.down[1], KC, N3, rows, grads)
        b_d.run()
        dsrc = b_d.dy(E, XD)
        ops.conv_gemm(dsrc, sv["W3"], Cin, dx, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0), add=extra_add, bcast=dxm,
                      bcast_scale=1.0 / T)
    else:
        dsrc = Act(E5)
        ops.conv_gemm(dsrc, sv["W3"], Cin, dx, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, Cin, 0), add=E4, add2=extra_add, bcast=dxm,
                      bcast_scale=1.0 / T)
    ops.conv_wgrad(x, dsrc, dW3, db=db3, n_samples=n, T_in=T, T_out=T, Vin=V)
    return dx
