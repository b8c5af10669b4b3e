"""Drop-in modules: same names, constructor signatures, child-module names and state-dict keys as the
reference, with forward/backward running on the sm_100a kernels.

  unit_gcn   pyskl/models/gcns/utils/gcn.py:22-97        unit_tcn  pyskl/models/gcns/utils/tcn.py:10-37
  dgphgcn1   pyskl/models/gcns/utils/gcn.py:2074-2373    mstcn     pyskl/models/gcns/utils/tcn.py:104-180
  DGBlock    pyskl/models/gcns/dgstgcn.py:12-70          dgmstcn   pyskl/models/gcns/utils/tcn.py:344-431
  DGSTGCN    pyskl/models/gcns/dgstgcn.py:73-170         STGCN     pyskl/models/gcns/stgcn.py:16-153

Parameters live in real nn.Conv2d / nn.BatchNorm2d children with the original names (checkpoints load
unchanged; analysis hooks that walk the children keep working).  Inputs/outputs are the reference's logical
[n, C, t, v] tensors; physically they are channels-last and in the compute dtype (bf16 by default,
fp32 for bit-for-bit-style parity runs — see `set_compute_dtype`).  There is no CPU fallback.
"""
import copy as cp
from math import ceil

import numpy as np
import torch
import torch.nn as nn

from . import functional as Fn
from . import ops
from .graph import Graph

EPS = 1e-4

_compute_dtype = torch.bfloat16


def set_compute_dtype(dtype):
    """torch.bfloat16 (default: bf16 storage, fp32 accumulation/statistics) or torch.float32 (exact path)."""
    global _compute_dtype
    assert dtype in (torch.float32, torch.bfloat16)
    _compute_dtype = dtype


def get_compute_dtype():
    return _compute_dtype


def build_norm_layer(cfg, num_features):
    cfg = cfg if isinstance(cfg, dict) else dict(type=cfg)
    if cfg.get("type") not in ("BN", "BN2d"):
        raise NotImplementedError(f"norm {cfg} (only BatchNorm is on the DS-GCN path)")
    return "bn", nn.BatchNorm2d(num_features, eps=cfg.get("eps", 1e-5))


def build_activation_layer(cfg):
    cfg = cfg if isinstance(cfg, dict) else dict(type=cfg)
    if cfg.get("type") != "ReLU":
        raise NotImplementedError(f"activation {cfg} (only ReLU is on the DS-GCN path)")
    return nn.ReLU()


# ------------------------------------------------------------------------------------------------
# layout helpers + autograd packaging
# ------------------------------------------------------------------------------------------------

def to_rows(x, dtype):
    """logical [n,C,t,v] -> channels-last rows [n*t*v, C] in `dtype` (a view when x already is that)."""
    n, c, t, v = x.shape
    y = x.permute(0, 2, 3, 1)
    if y.dtype != dtype:
        y = y.to(dtype)
    return y.contiguous().view(n * t * v, c)


def from_rows(y, n, t, v):
    return y.view(n, t, v, y.shape[-1]).permute(0, 3, 1, 2)


class _KernelFn(torch.autograd.Function):
    """Runs impl.fwd / impl.bwd (functional.py) and routes parameter gradients."""

    @staticmethod
    def forward(ctx, x, impl, *params):
        n, _, t, v = x.shape
        dtype = _compute_dtype
        need = any(ctx.needs_input_grad)
        save = {} if need else None
        out, t_out = impl.fwd(to_rows(x.detach(), dtype), n, t, v, save)
        ctx.impl, ctx.save, ctx.params, ctx.x_dtype, ctx.dims = impl, save, params, x.dtype, (n, t, t_out, v)
        ctx.dtype = dtype                 # the compute dtype of THIS forward (set_compute_dtype may change before backward)
        y = from_rows(out, n, t_out, v)
        if need:
            # the output aliases the tensor backward uses as its ReLU mask: registering it makes autograd's version counter catch an
            # in-place edit of a block output (relu_, add_) instead of silently corrupting the gradients
            ctx.save_for_backward(y)
        return y

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, dout):
        n, t, t_out, v = ctx.dims
        _ = ctx.saved_tensors             # raises if the output was modified in place
        grads = {}
        dx = ctx.impl.bwd(ctx.save, to_rows(dout, ctx.dtype), grads)
        ops.L.join_side()           # weight-gradient kernels ran on the side stream: rejoin before autograd sees them
        ctx.save = None
        dx = from_rows(dx, n, t, v)
        if dx.dtype != ctx.x_dtype:
            dx = dx.to(ctx.x_dtype)
        # gradients the kernels accumulated straight into the packed flat buffer (parallel.GradBuckets: p.grad is a view of it)
        # are complete here: autograd gets None for them and the bucket is told directly
        out, ready = [], []
        for p in ctx.params:
            g = grads.get(p)
            if g is not None and p.grad is not None and g.data_ptr() == p.grad.data_ptr():
                ready.append(p)
                g = None
            out.append(g)
        for p in ready:
            cb = getattr(p, "_dsg_ready", None)
            if cb is not None:
                cb(p)
        return (dx, None) + tuple(out)


class _Impl:
    def __init__(self, fwd, bwd):
        self.fwd, self.bwd = fwd, bwd


def _run(module, x, fwd, bwd):
    params = tuple(p for p in module.parameters() if p.requires_grad)
    with Fn.defer_bn_counters():
        return _KernelFn.apply(x, _Impl(fwd, bwd), *params)


def _check_input(x):
    if x.dim() != 4:
        raise ValueError(f"expected [n, C, t, v], got {tuple(x.shape)}")
    ops.L.check_tensor(x)


# ------------------------------------------------------------------------------------------------
# spatial units
# ------------------------------------------------------------------------------------------------

class unit_gcn(nn.Module):

    def __init__(self, in_channels, out_channels, A, adaptive='init', conv_pos='pre', with_res=False, norm='BN', act='ReLU'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_subsets = A.size(0)
        assert adaptive in [None, 'init', 'offset', 'importance']
        assert conv_pos in ['pre', 'post']
        self.adaptive, self.conv_pos, self.with_res = adaptive, conv_pos, with_res
        self.norm_cfg = norm if isinstance(norm, dict) else dict(type=norm)
        self.act_cfg = act if isinstance(act, dict) else dict(type=act)
        self.bn = build_norm_layer(self.norm_cfg, out_channels)[1]
        self.act = build_activation_layer(self.act_cfg)
        if adaptive == 'init':
            self.A = nn.Parameter(A.clone())
        else:
            self.register_buffer('A', A)
        if adaptive in ['offset', 'importance']:
            self.PA = nn.Parameter(A.clone())
            if adaptive == 'offset':
                nn.init.uniform_(self.PA, -1e-6, 1e-6)
            else:
                nn.init.constant_(self.PA, 1)
        if conv_pos == 'pre':
            self.conv = nn.Conv2d(in_channels, out_channels * A.size(0), 1)
        else:
            self.conv = nn.Conv2d(A.size(0) * in_channels, out_channels, 1)
        self.has_down = False
        if with_res:
            if in_channels != out_channels:
                self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), build_norm_layer(self.norm_cfg, out_channels)[1])
                self.has_down = True
            else:
                self.down = lambda x: x

    def _fwd(self, x, n, t, v, save):
        return Fn.unit_gcn_forward(self, x, n, t, v, save), t

    def _bwd(self, save, dout, grads, extra_add=None):
        return Fn.unit_gcn_backward(self, save, dout, grads, extra_add)

    def forward(self, x, A=None):
        _check_input(x)
        if A is not None:     # reference quirk (gcn.py:78-79): a passed A replaces self.A
            self.A = A
        return _run(self, x, self._fwd, self._bwd)

    def init_weights(self):
        pass


def _dg_unit_layers(self, in_channels, out_channels, A, ratio, norm, act, semantic, num_types=5, edge_num=15):
    """children of the dynamic-adjacency spatial units, registered in the reference's order (state-dict contract):
    gcn.py:1483-1512 (dggcn), :1643-1683 (dghgcn), :1863-1906 (dgphgcn), :2137-2215 (dgphgcn1)."""
    num_subsets = A.size(0)
    if ratio is None:
        ratio = 1 / num_subsets
    self.ratio = ratio
    mid_channels = int(ratio * out_channels)
    self.mid_channels = mid_channels
    self.norm_cfg = norm if isinstance(norm, dict) else dict(type=norm)
    self.act_cfg = act if isinstance(act, dict) else dict(type=act)
    self.act = build_activation_layer(self.act_cfg)
    self.A = nn.Parameter(A.clone())
    self.semantic_num = ceil(num_subsets / 3) if semantic else 0
    self.norm_num = num_subsets - self.semantic_num
    self.pre = nn.Sequential(nn.Conv2d(in_channels, mid_channels * num_subsets, 1),
                             build_norm_layer(self.norm_cfg, mid_channels * num_subsets)[1], self.act)
    self.post = nn.Conv2d(mid_channels * num_subsets, out_channels, 1)
    self.tanh, self.relu, self.sigmoid, self.softmax = nn.Tanh(), nn.ReLU(), nn.Sigmoid(), nn.Softmax(-2)
    self.alpha = nn.Parameter(torch.zeros(num_subsets))
    self.beta = nn.Parameter(torch.zeros(num_subsets))
    if semantic:
        self.conv1_se = nn.Conv2d(in_channels, self.semantic_num * mid_channels * num_types, kernel_size=1)
        self.conv2_se = nn.Conv2d(in_channels, self.semantic_num * mid_channels * num_types, kernel_size=1)   # allocated, unused (gcn.py:2253-2254)
    self.conv1 = nn.Conv2d(in_channels, self.norm_num * mid_channels, 1)
    self.conv2 = nn.Conv2d(in_channels, self.norm_num * mid_channels, 1)
    if semantic:
        self.edge_linears = nn.Conv2d(self.semantic_num * mid_channels, edge_num * self.semantic_num * mid_channels, 1)
    self.has_down = in_channels != out_channels
    if self.has_down:
        self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), build_norm_layer(self.norm_cfg, out_channels)[1])
    else:
        self.down = lambda x: x
    self.bn = build_norm_layer(self.norm_cfg, out_channels)[1]
    self._tab = {}


class dgphgcn1(nn.Module):

    def __init__(self, in_channels, out_channels, A, edge_type, node_type, ratio=0.25, decompose=False, ctr='T', ada='T',
                 node_attention=False, edge_attention=False, ada_attention=False, target_specific=False, add_type=False,
                 sub_att=True, stage=True, num_types=5, edge_num=15, subset_wise=True, ada_act='softmax', ctr_act='tanh',
                 norm='BN', act='ReLU'):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        num_subsets = A.size(0)
        self.num_subsets = num_subsets
        self.ctr, self.ada, self.ada_act, self.ctr_act = ctr, ada, ada_act, ctr_act
        self.node_attention, self.edge_attention = node_attention, edge_attention
        self.target_specific, self.ada_attention = target_specific, ada_attention
        self.num_types, self.edge_num = num_types, edge_num
        self.edge_type, self.node_type = edge_type, node_type     # plain attributes, as in the reference (gcn.py:2116-2117)
        self.add_type, self.decompose, self.subset_wise, self.sub_att = add_type, decompose, subset_wise, sub_att
        if stage is False:
            self.node_attention = self.edge_attention = self.target_specific = False
            self.decompose = decompose = False
            self.subset_wise = False
        if num_subsets != 3 or ctr != 'T' or ada != 'T' or ada_act != 'softmax' or ctr_act != 'tanh':
            raise NotImplementedError("the topology kernels cover ctr='T', ada='T', tanh / softmax, 3 subsets")
        full = (self.decompose and self.node_attention and self.edge_attention and self.subset_wise and sub_att
                and not ada_attention and not target_specific and num_types == 5 and edge_num == 15)
        # the plain DG-GCN unit (every attention flag off, e.g. stage=False, gcn.py:2122-2127): same kernels, variant 1
        self.plain = not (self.decompose or self.node_attention or self.edge_attention or ada_attention or target_specific) and sub_att
        if not (full or self.plain):
            raise NotImplementedError(
                "dgphgcn1 kernels cover the DS-GCN configuration (configs/dsstgcn/DSSTGCN_model.py: decompose, node_attention, "
                "edge_attention, subset_wise, sub_att) and the plain DG-GCN flag set (no attention flags); other flag sets are not built")
        _dg_unit_layers(self, in_channels, out_channels, A, ratio, norm, act, semantic=not self.plain, num_types=num_types,
                        edge_num=edge_num)

    def _tables(self, dev):
        """int32 device copies of node_type / edge_type (bit-exact integers; uploaded once per device)."""
        if self.plain:
            return None, None
        key = str(dev)
        if key not in self._tab:
            nt = torch.as_tensor(np.asarray(self.node_type), dtype=torch.int32).reshape(-1)
            et = torch.as_tensor(np.asarray(self.edge_type)).to(torch.int32).reshape(-1)
            V = self.A.shape[-1]
            if nt.numel() != V or et.numel() != V * V:
                raise ValueError("node_type / edge_type do not match the adjacency size")
            self._tab[key] = (nt.to(dev).contiguous(), et.to(dev).contiguous())
        return self._tab[key]

    def _flat_groups(self):
        """parameters the kernels always use concatenated (one GEMM): parallel.GradBuckets lays them out adjacently"""
        tw, tb = Fn._topo_params(self)
        g = {"Wt": tw, "bt": tb}
        if self.has_down:
            g["Wpd"] = [self.pre[0].weight, self.down[0].weight]
            g["bpd"] = [self.pre[0].bias, self.down[0].bias]
        return g

    def _fwd(self, x, n, t, v, save):
        return Fn.dgphgcn1_forward(self, x, n, t, v, save), t

    def _bwd(self, save, dout, grads, extra_add=None, pre=None):
        return Fn.dgphgcn1_backward(self, save, dout, grads, extra_add, pre=pre)

    def _bwd_tail(self, save):
        return Fn.dgphgcn1_backward_tail(self, save)

    def forward(self, x, A=None):     # A is ignored, as in the reference (gcn.py:2222)
        _check_input(x)
        return _run(self, x, self._fwd, self._bwd)

    def init_weights(self):
        pass


class dggcn(dgphgcn1):
    """DG-GCN spatial unit (gcn.py:1445-1584): the dynamic adjacency without the semantic decomposition — same kernels as dgphgcn1
    (dsg_topology_* variant 1).  ctr='T', ada='T', tanh / softmax are built; the 'NA' (per-frame) and None variants raise."""

    def __init__(self, in_channels, out_channels, A, ratio=0.25, ctr='T', ada='T', subset_wise=False, ada_act='softmax',
                 ctr_act='tanh', norm='BN', act='ReLU'):
        nn.Module.__init__(self)
        self._init_plain(in_channels, out_channels, A, ratio, ctr, ada, subset_wise, ada_act, ctr_act, norm, act)

    def _init_plain(self, in_channels, out_channels, A, ratio, ctr, ada, subset_wise, ada_act, ctr_act, norm, act):
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_subsets = A.size(0)
        self.ctr, self.ada, self.ada_act, self.ctr_act = ctr, ada, ada_act, ctr_act
        self.subset_wise = subset_wise
        if self.num_subsets != 3 or ctr != 'T' or ada != 'T' or ada_act != 'softmax' or ctr_act != 'tanh':
            raise NotImplementedError("the topology kernels cover ctr='T', ada='T', tanh / softmax, 3 subsets")
        self.plain = True
        _dg_unit_layers(self, in_channels, out_channels, A, ratio, norm, act, semantic=False)


class dghgcn(dggcn):
    """gcn.py:1586-1806 with its default flags (no node / edge / ada attention, no target-specific branch) is the DG-GCN unit; the
    attention flag sets of this earlier variant are not built (the DS-GCN configs use dgphgcn1)."""

    def __init__(self, in_channels, out_channels, A, edge_type, node_type, ratio=0.25, ctr='T', ada='T', node_attention=False,
                 edge_attention=False, ada_attention=False, target_specific=False, add_type=False, num_types=5, edge_num=15,
                 subset_wise=False, ada_act='softmax', ctr_act='tanh', norm='BN', act='ReLU'):
        nn.Module.__init__(self)
        if node_attention or edge_attention or ada_attention or target_specific:
            raise NotImplementedError(f"{type(self).__name__}: only the flag set without attention branches is built "
                                      "(use dgphgcn1 for the semantic decomposition)")
        self.node_attention, self.edge_attention = node_attention, edge_attention
        self.target_specific, self.ada_attention, self.add_type = target_specific, ada_attention, add_type
        self.num_types, self.edge_num = num_types, edge_num
        self.edge_type, self.node_type = edge_type, node_type
        self._init_plain(in_channels, out_channels, A, ratio, ctr, ada, subset_wise, ada_act, ctr_act, norm, act)


class dgphgcn(dghgcn):
    """gcn.py:1808-2072; without attention flags `part_ratio` only sets attributes (gcn.py:1889-1904)."""

    def __init__(self, in_channels, out_channels, A, edge_type, node_type, ratio=0.25, part_ratio=0.4, **kw):
        _ = kw.get('node_attention', False) & part_ratio      # gcn.py:1892 evaluates `bool & part_ratio`: a float part_ratio (the default!) raises TypeError there too
        super().__init__(in_channels, out_channels, A, edge_type, node_type, ratio=ratio, **kw)
        self.part_ratio = part_ratio
        self.semantic_num = int(self.num_subsets * part_ratio)
        self.norm_num = self.num_subsets - self.semantic_num


def _conv_init(conv):       # init_func.py:15-17
    nn.init.kaiming_normal_(conv.weight, mode='fan_out')
    nn.init.constant_(conv.bias, 0)


class CTRGC(nn.Module):
    """Parameter holder of one channel-wise topology refinement graph convolution (gcn.py:634-666); unit_ctrgcn runs the three of
    a unit together on the kernels (one feature GEMM, one adjacency kernel, one contraction)."""

    def __init__(self, in_channels, out_channels, rel_reduction=8):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.rel_channels = 8 if in_channels <= 16 else in_channels // rel_reduction
        self.conv1 = nn.Conv2d(in_channels, self.rel_channels, kernel_size=1)
        self.conv2 = nn.Conv2d(in_channels, self.rel_channels, kernel_size=1)
        self.conv3 = nn.Conv2d(in_channels, out_channels, kernel_size=1)
        self.conv4 = nn.Conv2d(self.rel_channels, out_channels, kernel_size=1)
        self.tanh = nn.Tanh()
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _conv_init(m)

    def forward(self, x, A=None, alpha=1):
        raise NotImplementedError("CTRGC runs inside unit_ctrgcn (the three subsets share one set of kernel launches)")


class unit_ctrgcn(nn.Module):
    """gcn.py:882-930: relu(bn(sum_k CTRGC_k(x, A[k], alpha)) + down(x))."""

    def __init__(self, in_channels, out_channels, A):
        super().__init__()
        self.inter_c, self.out_c, self.in_c = out_channels // 4, out_channels, in_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.num_subset = A.shape[0]
        if self.num_subset != 3:
            raise NotImplementedError("unit_ctrgcn kernels are built for 3 adjacency subsets")
        self.convs = nn.ModuleList([CTRGC(in_channels, out_channels) for _ in range(self.num_subset)])
        self.rel_channels = self.convs[0].rel_channels
        self.has_down = in_channels != out_channels
        if self.has_down:
            self.down = nn.Sequential(nn.Conv2d(in_channels, out_channels, 1), nn.BatchNorm2d(out_channels))
        else:
            self.down = lambda x: x
        self.A = nn.Parameter(A.clone())
        self.alpha = nn.Parameter(torch.zeros(1))
        self.bn = nn.BatchNorm2d(out_channels)
        self.soft = nn.Softmax(-2)
        self.relu = nn.ReLU(inplace=True)
        self._wsum = {}

    def _sum_weight(self, dev):
        """[I | I | I]: the subset sum z = y_0 + y_1 + y_2 as a GEMM operand (exact in fp32 accumulation)"""
        key = str(dev)
        if key not in self._wsum:
            self._wsum[key] = torch.eye(self.out_c, dtype=torch.float32, device=dev).repeat(1, self.num_subset).contiguous()
        return self._wsum[key]

    def _flat_groups(self):
        return Fn._ctr_groups(self)

    def _fwd(self, x, n, t, v, save):
        return Fn.unit_ctrgcn_forward(self, x, n, t, v, save), t

    def _bwd(self, save, dout, grads, extra_add=None):
        return Fn.unit_ctrgcn_backward(self, save, dout, grads, extra_add)

    def forward(self, x):
        _check_input(x)
        return _run(self, x, self._fwd, self._bwd)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                _conv_init(m)
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        nn.init.constant_(self.bn.weight, 1e-6)
        nn.init.constant_(self.bn.bias, 0)


# ------------------------------------------------------------------------------------------------
# temporal units
# ------------------------------------------------------------------------------------------------

class unit_tcn(nn.Module):

    def __init__(self, in_channels, out_channels, kernel_size=9, stride=1, dilation=1, norm='BN', dropout=0):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm_cfg = norm if isinstance(norm, dict) else dict(type=norm)
        pad = (kernel_size + (kernel_size - 1) * (dilation - 1) - 1) // 2
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size=(kernel_size, 1), padding=(pad, 0), stride=(stride, 1),
                              dilation=(dilation, 1))
        self.bn = build_norm_layer(self.norm_cfg, out_channels)[1] if norm is not None else nn.Identity()
        if dropout:
            raise NotImplementedError("dropout > 0 is not on the DS-GCN path (configs use 0)")
        self.drop = nn.Dropout(dropout, inplace=True)
        self.stride, self.kernel_size, self.dilation = stride, kernel_size, dilation

    def _fwd(self, x, n, t, v, save, res=None, final_relu=False):
        return Fn.unit_tcn_forward(self, x, n, t, v, save, res, final_relu)

    def _bwd(self, save, dout, grads):
        return Fn.unit_tcn_backward(self, save, dout, grads)

    def forward(self, x):
        _check_input(x)
        return _run(self, x, self._fwd, lambda s, d, g: self._bwd(s, d, g)[0])

    def init_weights(self):
        pass


class mstcn(nn.Module):
    has_ext = False

    def __init__(self, in_channels, out_channels, mid_channels=None, dropout=0.,
                 ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ('max', 3), '1x1'], stride=1):
        super().__init__()
        self._build(in_channels, out_channels, mid_channels, dropout, ms_cfg, stride)

    def _build(self, in_channels, out_channels, mid_channels, dropout, ms_cfg, stride):
        self.ms_cfg = ms_cfg
        num_branches = len(ms_cfg)
        self.num_branches, self.in_channels, self.out_channels, self.stride = num_branches, in_channels, out_channels, stride
        self.act = nn.ReLU()
        if mid_channels is None:
            mid_channels = out_channels // num_branches
            rem_mid_channels = out_channels - mid_channels * (num_branches - 1)
        else:
            assert isinstance(mid_channels, float) and mid_channels > 0
            mid_channels = int(out_channels * mid_channels)
            rem_mid_channels = mid_channels
        self.mid_channels, self.rem_mid_channels = mid_channels, rem_mid_channels
        branches = []
        for i, cfg in enumerate(ms_cfg):
            branch_c = rem_mid_channels if i == 0 else mid_channels
            if cfg == '1x1':
                branches.append(nn.Conv2d(in_channels, branch_c, kernel_size=1, stride=(stride, 1)))
                continue
            assert isinstance(cfg, tuple)
            if cfg[0] == 'max':
                if cfg[1] != 3:
                    raise NotImplementedError("only the 3x1 max-pool branch is built")
                branches.append(nn.Sequential(nn.Conv2d(in_channels, branch_c, kernel_size=1), nn.BatchNorm2d(branch_c), self.act,
                                              nn.MaxPool2d(kernel_size=(cfg[1], 1), stride=(stride, 1), padding=(1, 0))))
                continue
            assert isinstance(cfg[0], int) and isinstance(cfg[1], int)
            branches.append(nn.Sequential(nn.Conv2d(in_channels, branch_c, kernel_size=1), nn.BatchNorm2d(branch_c), self.act,
                                          unit_tcn(branch_c, branch_c, kernel_size=cfg[0], stride=stride, dilation=cfg[1], norm=None)))
        self.branches = nn.ModuleList(branches)
        tin_channels = mid_channels * (num_branches - 1) + rem_mid_channels
        self.transform = nn.Sequential(nn.BatchNorm2d(tin_channels), self.act, nn.Conv2d(tin_channels, out_channels, kernel_size=1))
        self.bn = nn.BatchNorm2d(out_channels)
        if dropout:
            raise NotImplementedError("dropout > 0 is not on the DS-GCN path (configs use 0)")
        self.drop = nn.Dropout(dropout, inplace=True)

    def _flat_groups(self):
        convs = [b if isinstance(b, nn.Conv2d) else b[0] for b in self.branches]
        return {"Wbr": [c.weight for c in convs], "bbr": [c.bias for c in convs]}

    def _fwd(self, x, n, t, v, save, res=None, final_relu=False):
        return Fn.mstcn_forward(self, x, n, t, v, save, res, final_relu)

    def _bwd(self, save, dout, grads, tail=None):
        return Fn.mstcn_backward(self, save, dout, grads, tail=tail)

    def forward(self, x):
        _check_input(x)
        return _run(self, x, self._fwd, lambda s, d, g: self._bwd(s, d, g)[0])

    def init_weights(self):
        pass


class dgmstcn(mstcn):
    has_ext = True

    def __init__(self, in_channels, out_channels, mid_channels=None, num_joints=25, dropout=0.,
                 ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ('max', 3), '1x1'], stride=1):
        nn.Module.__init__(self)
        self.num_joints = num_joints
        self.add_coeff = nn.Parameter(torch.zeros(num_joints))
        self._build(in_channels, out_channels, mid_channels, dropout, ms_cfg, stride)


class MSTCN(mstcn):
    """MS-G3D / CTR-GCN multi-scale temporal unit (pyskl/models/gcns/utils/msg3d_utils.py:64-150): 1x1 conv + BN + ReLU into
    (k x 1) dilated convs / a 3x1 max-pool / a strided 1x1, every branch closed by its own BatchNorm, out = relu(cat + residual).
    Runs on the branch-stage kernels of mstcn (no transform conv; the last branch takes the channel remainder)."""
    has_ext = False
    no_transform = True

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilations=[1, 2, 3, 4], residual=True,
                 act_cfg=dict(type='ReLU'), tcn_dropout=0):
        nn.Module.__init__(self)
        self.num_branches = len(dilations) + 2
        branch_channels = out_channels // self.num_branches
        branch_channels_rem = out_channels - branch_channels * (self.num_branches - 1)
        if type(kernel_size) == list:
            assert len(kernel_size) == len(dilations)
        else:
            kernel_size = [kernel_size] * len(dilations)
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, stride
        self.ms_cfg = [(ks, d) for ks, d in zip(kernel_size, dilations)] + [('max', 3), '1x1']
        self.branch_widths = [branch_channels] * (self.num_branches - 1) + [branch_channels_rem]
        self.branches = nn.ModuleList([
            nn.Sequential(nn.Conv2d(in_channels, branch_channels, kernel_size=1, padding=0), nn.BatchNorm2d(branch_channels),
                          build_activation_layer(act_cfg),
                          unit_tcn(branch_channels, branch_channels, kernel_size=ks, stride=stride, dilation=d))
            for ks, d in zip(kernel_size, dilations)])
        self.branches.append(nn.Sequential(
            nn.Conv2d(in_channels, branch_channels, kernel_size=1, padding=0), nn.BatchNorm2d(branch_channels),
            build_activation_layer(act_cfg), nn.MaxPool2d(kernel_size=(3, 1), stride=(stride, 1), padding=(1, 0)),
            nn.BatchNorm2d(branch_channels)))
        self.branches.append(nn.Sequential(
            nn.Conv2d(in_channels, branch_channels_rem, kernel_size=1, padding=0, stride=(stride, 1)),
            nn.BatchNorm2d(branch_channels_rem)))
        if not residual:
            self.residual = lambda x: 0
            self.res_kind = 'none'
        elif in_channels == out_channels and stride == 1:
            self.residual = lambda x: x
            self.res_kind = 'identity'
        else:
            self.residual = unit_tcn(in_channels, out_channels, kernel_size=1, stride=stride)
            self.res_kind = 'conv'
        self.act = build_activation_layer(act_cfg)
        if tcn_dropout:
            raise NotImplementedError("dropout > 0 is not built (the configs use 0)")
        self.drop = nn.Dropout(tcn_dropout)

    def _feat_bns(self, layout):
        """the BatchNorm that closes each branch, with its channel range in the concatenated output"""
        out = []
        for j, (kind, lo, hi, _) in enumerate(layout):
            out.append(((self.branches[j][3].bn if kind == "conv" else self.branches[j][4] if kind == "max" else self.branches[j][1]), lo, hi))
        return out

    def _own_residual(self, x, n, t, v, save):
        if self.res_kind == 'none':
            return None
        if self.res_kind == 'identity':
            return (x, None, None)
        sv = {} if save is not None else None
        rr, c_r, _ = Fn.unit_tcn_raw_forward(self.residual, x, n, t, v, sv)
        if save is not None:
            save["own_res"] = sv
        return (rr, c_r.a, c_r.b)

    def _own_residual_backward(self, save, E, dg, grads):
        if self.res_kind == 'none':
            return dg
        extra = E if self.res_kind == 'identity' else Fn.unit_tcn_raw_backward(self.residual, save["own_res"], E, grads)
        out = torch.empty_like(dg)
        ops.pointwise(ops.Act(dg, None, None, extra), out)
        return out


# ------------------------------------------------------------------------------------------------
# blocks and backbones
# ------------------------------------------------------------------------------------------------

class _STBlock(nn.Module):
    """relu( tcn(gcn(x)) + residual(x) ) — dgstgcn.py:61-65 / stgcn.py:65-68, one fused forward/backward."""

    def _make_residual(self, in_channels, out_channels, stride, residual):
        if not residual:
            self.residual = lambda x: 0
            self.res_kind = 'none'
        elif in_channels == out_channels and stride == 1:
            self.residual = lambda x: x
            self.res_kind = 'identity'
        else:
            self.residual = unit_tcn(in_channels, out_channels, kernel_size=1, stride=stride)
            self.res_kind = 'conv'

    def _fwd(self, x, n, t, v, save):
        sv = dict(gcn={}, tcn={}, res={}) if save is not None else dict(gcn=None, tcn=None, res=None)
        g, _ = self.gcn._fwd(x, n, t, v, sv["gcn"])
        if self.res_kind == 'none':
            res = None
        elif self.res_kind == 'identity':
            res = (x, None, None)
        else:
            rr, c_r, _ = Fn.unit_tcn_raw_forward(self.residual, x, n, t, v, sv["res"])
            res = (rr, c_r.a, c_r.b)
        out, t_out = self.tcn._fwd(g, n, t, v, sv["tcn"], res=res, final_relu=True)
        if save is not None:
            save.update(sv)
        return out, t_out

    def _bwd(self, save, dout, grads):
        # the spatial unit's output mask / BN-backward sums ride in the epilogue of the temporal unit's last backward GEMM
        pre = None
        if isinstance(self.gcn, dgphgcn1) and isinstance(self.tcn, mstcn):
            pre = self.gcn._bwd_tail(save["gcn"])
            dg, E = self.tcn._bwd(save["tcn"], dout, grads, tail=pre)
        else:
            dg, E = self.tcn._bwd(save["tcn"], dout, grads)
        extra = None
        if self.res_kind == 'identity':
            extra = E
        elif self.res_kind == 'conv':
            extra = Fn.unit_tcn_raw_backward(self.residual, save["res"], E, grads)
        if pre is not None:
            return self.gcn._bwd(save["gcn"], dg, grads, extra_add=extra, pre=pre)
        return self.gcn._bwd(save["gcn"], dg, grads, extra_add=extra)

    def forward(self, x, A=None):
        _check_input(x)
        if A is not None and isinstance(self.gcn, unit_gcn):
            self.gcn.A = A
        return _run(self, x, self._fwd, self._bwd)

    def init_weights(self):
        pass


class DGBlock(_STBlock):

    def __init__(self, in_channels, out_channels, A, edge_type, node_type, stride=1, residual=True, **kwargs):
        super().__init__()
        for arg in ['act', 'norm', 'g1x1']:
            if arg in kwargs:
                value = kwargs.pop(arg)
                kwargs['tcn_' + arg] = value
                kwargs['gcn_' + arg] = value
        gcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'gcn_'}
        tcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'tcn_'}
        kwargs = {k: v for k, v in kwargs.items() if k[1:4] != 'cn_'}
        assert len(kwargs) == 0
        tcn_type = tcn_kwargs.pop('type', 'unit_tcn')
        assert tcn_type in ['unit_tcn', 'mstcn', 'dgmstcn', 'dgmsmlp']
        if tcn_type == 'unit_tcn':
            self.tcn = unit_tcn(out_channels, out_channels, 9, stride=stride, **tcn_kwargs)
        elif tcn_type == 'mstcn':
            self.tcn = mstcn(out_channels, out_channels, stride=stride, **tcn_kwargs)
        elif tcn_type == 'dgmstcn':
            self.tcn = dgmstcn(out_channels, out_channels, stride=stride, **tcn_kwargs)
        else:
            raise NotImplementedError("tcn_type='dgmsmlp' is an author experiment outside the DS-GCN configs (SURVEY.md §2 row 3)")
        gcn_type = gcn_kwargs.pop('type', 'dghgcn')
        assert gcn_type in ['dghgcn', 'dgphgcn', 'dgphgcn1', 'dggcn']
        if gcn_type == 'dggcn':
            self.gcn = dggcn(in_channels, out_channels, A, **gcn_kwargs)
        else:
            cls = dict(dghgcn=dghgcn, dgphgcn=dgphgcn, dgphgcn1=dgphgcn1)[gcn_type]
            self.gcn = cls(in_channels, out_channels, A, edge_type, node_type, **gcn_kwargs)
        self.relu = nn.ReLU()
        self._make_residual(in_channels, out_channels, stride, residual)


class STGCNBlock(_STBlock):

    def __init__(self, in_channels, out_channels, A, stride=1, residual=True, **kwargs):
        super().__init__()
        gcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'gcn_'}
        tcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'tcn_'}
        kwargs = {k: v for k, v in kwargs.items() if k[:4] not in ['gcn_', 'tcn_']}
        assert len(kwargs) == 0, f'Invalid arguments: {kwargs}'
        tcn_type = tcn_kwargs.pop('type', 'unit_tcn')
        gcn_type = gcn_kwargs.pop('type', 'unit_gcn')
        if gcn_type != 'unit_gcn' or tcn_type not in ('unit_tcn', 'mstcn'):
            raise NotImplementedError(f"STGCNBlock with gcn_type={gcn_type}, tcn_type={tcn_type} is not built")
        self.gcn = unit_gcn(in_channels, out_channels, A, **gcn_kwargs)
        if tcn_type == 'unit_tcn':
            self.tcn = unit_tcn(out_channels, out_channels, 9, stride=stride, **tcn_kwargs)
        else:
            self.tcn = mstcn(out_channels, out_channels, stride=stride, **tcn_kwargs)
        self.relu = nn.ReLU()
        self._make_residual(in_channels, out_channels, stride, residual)


class CTRGCNBlock(_STBlock):
    """ctrgcn.py:9-66: relu(MSTCN(unit_ctrgcn(x)) + residual(x)); children gcn1 / tcn1 as in the reference."""

    def __init__(self, in_channels, out_channels, A, edge_type=None, node_type=None, semantic_index=False, stride=1, residual=True,
                 kernel_size=5, dilations=[1, 2], tcn_dropout=0, **kwargs):
        super().__init__()
        gcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'gcn_'}
        tcn_kwargs = {k[4:]: v for k, v in kwargs.items() if k[:4] == 'tcn_'}
        kwargs = {k: v for k, v in kwargs.items() if k[:4] not in ['gcn_', 'tcn_']}
        assert len(kwargs) == 0, f'Invalid arguments: {kwargs}'
        tcn_type = tcn_kwargs.pop('type', 'mstcn')
        gcn_type = gcn_kwargs.pop('type', 'unit_ctrhgcn')
        if gcn_type != 'unit_ctrgcn' or tcn_type != 'mstcn':
            raise NotImplementedError(f"CTRGCNBlock with gcn_type={gcn_type}, tcn_type={tcn_type}: only unit_ctrgcn + mstcn (CTR-GCN proper) "
                                      "is built; unit_ctrhgcn / msmlp are author experiments (SURVEY.md §2)")
        self.gcn1 = unit_ctrgcn(in_channels, out_channels, A, **gcn_kwargs)
        self.tcn1 = MSTCN(out_channels, out_channels, kernel_size=kernel_size, stride=stride, dilations=dilations, residual=False,
                          tcn_dropout=tcn_dropout)
        self.relu = nn.ReLU(inplace=True)
        self._make_residual(in_channels, out_channels, stride, residual)

    gcn = property(lambda self: self.gcn1)       # the names _STBlock's fused forward / backward use
    tcn = property(lambda self: self.tcn1)


class _DataBNFn(torch.autograd.Function):
    """data_bn (dgstgcn.py:158-164): rows (n*m, t) x channels v*C+c is already the channels-last activation."""

    @staticmethod
    def forward(ctx, x, bn, weight, bias, mvc=False):
        N, M, T, V, C = x.shape
        dev = x.device
        if mvc:     # ctrgcn.py:112-114: BatchNorm1d over channels m*V*C + v*C + c, statistics over (N, T): rows (n, t), person slot in the channel
            xin = x.detach().permute(0, 2, 1, 3, 4).contiguous().view(N * T, M * V * C)
        else:
            xin = x.detach().contiguous().view(N * M * T, V * C)
        if xin.dtype != torch.float32:
            xin = xin.float()
        nrow, CH = xin.shape
        coef = Fn.BNCoef(CH, dev, [bn])
        if coef.training:
            ops.pointwise(xin, None, stat_sum=coef.ssum, stat_sq=coef.ssq)
        coef.add_bn(bn, 0, CH, nrow)
        coef.run()
        out = torch.empty(nrow, CH, dtype=_compute_dtype, device=dev)
        ops.pointwise(ops.Act(xin, coef.a, coef.b), out)
        ctx.save, ctx.bn, ctx.dims, ctx.need_x = (xin, coef), bn, (N, M, T, V, C), ctx.needs_input_grad[0]
        ctx.x_dtype, ctx.mvc = x.dtype, mvc
        if mvc:
            out = out.view(N, T, M, V, C).permute(0, 2, 1, 3, 4).contiguous()
        return out.view(N * M, T, V, C).permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        N, M, T, V, C = ctx.dims
        xin, coef = ctx.save
        bn = ctx.bn
        nrow, CH = xin.shape
        if ctx.mvc:
            d = dout.permute(0, 2, 3, 1).reshape(N, M, T, V, C).permute(0, 2, 1, 3, 4).contiguous().view(nrow, CH)
        else:
            d = dout.permute(0, 2, 3, 1).contiguous().view(nrow, CH)
        back = Fn.BNBack(coef)
        ops.pointwise(d, None, stat_sum=back.ssum, stat_sq=back.ssq, partner=xin)
        grads = {}
        back.add_bn(bn, 0, CH, nrow, grads)
        back.run()
        dx = None
        if ctx.need_x:
            d32 = d if d.dtype == torch.float32 else d.float()
            dx = torch.empty(nrow, CH, dtype=torch.float32, device=d.device)
            ops.pointwise(ops.Act(d32, back.ca, back.cc, xin, back.cb), dx)
            dx = (dx.view(N, T, M, V, C).permute(0, 2, 1, 3, 4) if ctx.mvc else dx.view(N, M, T, V, C)).to(ctx.x_dtype)
        gw, gb_ = grads[bn.weight], grads[bn.bias]
        for q in (bn.weight, bn.bias):          # written in place into the packed flat gradient buffer (see _KernelFn.backward)
            if q.grad is not None and grads[q].data_ptr() == q.grad.data_ptr():
                cb = getattr(q, "_dsg_ready", None)
                if cb is not None:
                    cb(q)
        if bn.weight.grad is not None and gw.data_ptr() == bn.weight.grad.data_ptr():
            gw = None
        if bn.bias.grad is not None and gb_.data_ptr() == bn.bias.grad.data_ptr():
            gb_ = None
        return dx, None, gw, gb_, None


class _Backbone(nn.Module):

    def _setup(self, block_fn, graph_cfg, in_channels, base_channels, ch_ratio, num_stages, inflate_stages, down_stages,
               data_bn_type, num_person, pretrained, kwargs, pop_first):
        self.data_bn_type = data_bn_type
        self.kwargs = kwargs
        V = self.graph.A.shape[1]
        if data_bn_type == 'MVC':
            self.data_bn = nn.BatchNorm1d(num_person * in_channels * V)
        elif data_bn_type == 'VC':
            self.data_bn = nn.BatchNorm1d(in_channels * V)
        else:
            self.data_bn = nn.Identity()
        lw_kwargs = [cp.deepcopy(kwargs) for _ in range(num_stages)]
        for k, v in kwargs.items():
            if isinstance(v, tuple) and len(v) == num_stages:
                for i in range(num_stages):
                    lw_kwargs[i][k] = v[i]
        for k in pop_first:
            lw_kwargs[0].pop(k, None)
        if 'gcn_stage' in kwargs:
            for i in range(num_stages):
                lw_kwargs[i]['gcn_stage'] = i in kwargs['gcn_stage']
        self.in_channels, self.base_channels, self.ch_ratio = in_channels, base_channels, ch_ratio
        self.inflate_stages, self.down_stages = inflate_stages, down_stages
        modules = []
        if in_channels != base_channels:
            modules = [block_fn(in_channels, base_channels, 1, False, lw_kwargs[0])]
        inflate_times = 0
        for i in range(2, num_stages + 1):
            stride = 1 + (i in down_stages)
            cin = base_channels
            if i in inflate_stages:
                inflate_times += 1
            cout = int(self.base_channels * self.ch_ratio ** inflate_times + EPS)
            base_channels = cout
            modules.append(block_fn(cin, cout, stride, True, lw_kwargs[i - 1]))
        if self.in_channels == self.base_channels:
            num_stages -= 1
        self.num_stages = num_stages
        self.gcn = nn.ModuleList(modules)
        self.pretrained = pretrained

    def init_weights(self):
        if isinstance(self.pretrained, str):
            sd = torch.load(self.pretrained, map_location='cpu')
            sd = sd.get('state_dict', sd)
            sd = {k[len('backbone.'):] if k.startswith('backbone.') else k: v for k, v in sd.items()}
            self.load_state_dict(sd, strict=False)

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError(f"expected [N, M, T, V, C], got {tuple(x.shape)}")
        ops.L.check_tensor(x)
        N, M, T, V, C = x.size()
        with Fn.defer_bn_counters(), Fn.stat_arena(x.device):
            if self.data_bn_type == 'VC':
                h = _DataBNFn.apply(x, self.data_bn, self.data_bn.weight, self.data_bn.bias)
            elif self.data_bn_type == 'MVC':
                h = _DataBNFn.apply(x, self.data_bn, self.data_bn.weight, self.data_bn.bias, True)
            else:
                h = x.reshape(N * M, T, V, C).permute(0, 3, 1, 2)
            for i in range(self.num_stages):
                h = self.gcn[i](h)
        return h.reshape((N, M) + h.shape[1:])


class DGSTGCN(_Backbone):

    def __init__(self, graph_cfg, in_channels=3, base_channels=64, ch_ratio=2, num_stages=10, inflate_stages=[5, 8],
                 down_stages=[5, 8], data_bn_type='VC', num_person=2, pretrained=None, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_cfg)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        if not hasattr(self.graph, 'node_type'):
            raise ValueError(f"layout {graph_cfg.get('layout')} defines no node_type/edge_type (reference: dgstgcn.py:93-95)")
        node_type = torch.tensor(self.graph.node_type, requires_grad=False)
        edge_type = torch.tensor(self.graph.edge_type, dtype=torch.float32, requires_grad=False)
        mk = lambda cin, cout, stride, residual, kw: DGBlock(cin, cout, A.clone(), edge_type, node_type, stride, residual=residual, **kw)
        self._setup(mk, graph_cfg, in_channels, base_channels, ch_ratio, num_stages, inflate_stages, down_stages, data_bn_type,
                    num_person, pretrained, kwargs, pop_first=('tcn_dropout', 'g1x1', 'gcn_g1x1'))


class STGCN(_Backbone):

    def __init__(self, graph_cfg, in_channels=3, base_channels=64, data_bn_type='VC', ch_ratio=2, num_person=2, num_stages=10,
                 inflate_stages=[5, 8], down_stages=[5, 8], pretrained=None, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_cfg)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        mk = lambda cin, cout, stride, residual, kw: STGCNBlock(cin, cout, A.clone(), stride, residual=residual, **kw)
        self._setup(mk, graph_cfg, in_channels, base_channels, ch_ratio, num_stages, inflate_stages, down_stages, data_bn_type,
                    num_person, pretrained, kwargs, pop_first=('tcn_dropout',))


class CTRGCN(nn.Module):
    """ctrgcn.py:69-125 with gcn_type='unit_ctrgcn' (CTR-GCN proper): data_bn over (person, joint, channel), 10 CTRGCNBlocks in `net`."""

    def __init__(self, graph_cfg, in_channels=3, base_channels=64, num_stages=10, inflate_stages=[5, 8], down_stages=[5, 8],
                 semantic_stage=range(1, 11), pretrained=None, num_person=2, **kwargs):
        super().__init__()
        self.graph = Graph(**graph_cfg)
        A = torch.tensor(self.graph.A, dtype=torch.float32, requires_grad=False)
        self.register_buffer('A', A)
        node_type = torch.tensor(self.graph.node_type, requires_grad=False)
        edge_type = torch.tensor(self.graph.edge_type, dtype=torch.float32, requires_grad=False)
        self.num_person, self.base_channels = num_person, base_channels
        self.data_bn = nn.BatchNorm1d(num_person * in_channels * A.size(1))
        kwargs0 = {k: v for k, v in kwargs.items() if k != 'tcn_dropout'}
        modules = [CTRGCNBlock(in_channels, base_channels, A.clone(), edge_type, node_type, 1 in semantic_stage, residual=False, **kwargs0)]
        for i in range(2, num_stages + 1):
            out_channels = base_channels * (1 + (i in inflate_stages))
            stride = 1 + (i in down_stages)
            modules.append(CTRGCNBlock(base_channels, out_channels, A.clone(), edge_type, node_type, i in semantic_stage, stride=stride,
                                       **kwargs))
            base_channels = out_channels
        self.net = nn.ModuleList(modules)
        self.pretrained = pretrained

    def init_weights(self):
        for module in self.net:
            module.init_weights()

    def forward(self, x):
        if x.dim() != 5:
            raise ValueError(f"expected [N, M, T, V, C], got {tuple(x.shape)}")
        ops.L.check_tensor(x)
        N, M, T, V, C = x.size()
        with Fn.defer_bn_counters(), Fn.stat_arena(x.device):
            h = _DataBNFn.apply(x, self.data_bn, self.data_bn.weight, self.data_bn.bias, True)
            for blk in self.net:
                h = blk(h)
        return h.reshape((N, M) + h.shape[1:])
