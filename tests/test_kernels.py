"""Primitive-level parity: every C-ABI entry point against a plain PyTorch fp32 restatement of the same op.

Each test runs twice: `dev=sim` (the kernel sources on the host-side SIMT simulator — CPU-only container) and
`dev=cuda` (marked gpu: the real sm_100a library on the B200).  Tolerances: fp32 path rel 1e-4 (north_star),
bf16 path rel-L2 1e-2.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import dsgcn_b200
from dsgcn_b200 import ops
_lib = ops.L
from oracle import dsgcn_oracle as O

DTYPES = [torch.float32, torch.bfloat16]


def close(got, ref, dtype, what=""):
    _lib.join_side()      # weight-gradient kernels run on the side stream: make the current stream wait for them
    got, ref = got.detach().double().cpu(), ref.detach().double().cpu()
    err = (got - ref).norm() / (ref.norm() + 1e-12)
    tol = 1e-4 if dtype == torch.float32 else 1e-2
    assert err < tol, f"{what}: rel-L2 {err:.3e} > {tol}"


def rnd(*shape, dev, dtype=torch.float32, scale=1.0):
    return (torch.randn(*shape) * scale).to(dtype).to(dev)


def to_cl(x):
    """logical [n,C,t,v] -> rows [(n,t,v), C]"""
    n, c, t, v = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * t * v, c).contiguous()


def from_cl(y, n, t, v):
    return y.reshape(n, t, v, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_gemm_pointwise_prologue_epilogue(dev, dtype):
    torch.manual_seed(0)
    n, T, V, K, N = 2, 7, 5, 19, 70
    x, x2 = rnd(n * T * V, K, dev=dev, dtype=dtype), rnd(n * T * V, K, dev=dev, dtype=dtype)
    W, b = rnd(N, K, dev=dev, scale=0.3), rnd(N, dev=dev)
    a1, b1, a2, b2 = (torch.rand(K, device=dev) + 0.5), rnd(K, dev=dev), (torch.rand(K, device=dev) + 0.5), rnd(K, dev=dev)
    add = rnd(n * T * V, N, dev=dev, dtype=dtype)
    bc = rnd(n, V, N, dev=dev)
    mk = rnd(n * T * V, N, dev=dev, dtype=dtype)
    partner = rnd(n * T * V, N, dev=dev, dtype=dtype)
    out = torch.empty(n * T * V, N, dtype=dtype, device=dev)
    ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
    ops.conv_gemm(ops.Act(x, a1, b1, x2, a2, b2, relu=True), W, N, out, n_samples=n, T_in=T, T_out=T, Vin=V, bias=b,
                  add=add, bcast=bc, bcast_scale=0.25, mask=mk, stat_sum=ss, stat_sq=sq, partner=partner)
    src = torch.relu(x.float() * a1 + b1 + x2.float() * a2 + b2)
    ref = src @ W.t() + b + add.float() + 0.25 * bc[:, None].expand(n, T, V, N).reshape(-1, N)
    ref = ref * (mk.float() > 0)
    close(out, ref, dtype, "out")
    close(ss, ref.sum(0), dtype, "sum")
    close(sq, (ref * partner.float()).sum(0), dtype, "sumprod")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("k,stride,dil", [(3, 1, 1), (3, 1, 4), (3, 2, 2), (3, 2, 3), (9, 1, 1), (9, 2, 1), (1, 2, 1)])
def test_conv_gemm_temporal_conv_and_grads(dev, dtype, k, stride, dil):
    """forward = nn.Conv2d((k,1), stride, dilation, padding); dgrad via the transposed frame map; wgrad."""
    torch.manual_seed(1)
    n, T, V, K, N = 2, 11, 4, 10, 14
    pad = (k + (k - 1) * (dil - 1) - 1) // 2
    x = torch.randn(n, K, T, V)
    w = torch.randn(N, K, k, 1) * 0.3
    b = torch.randn(N)
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    br = b.clone().requires_grad_()
    ref = F.conv2d(xr, wr, br, (stride, 1), (pad, 0), (dil, 1))
    T_out = ref.shape[2]
    gy = torch.randn_like(ref)
    ref.backward(gy)
    xc = to_cl(x).to(dtype).to(dev)
    wd, bd = w.to(dev).contiguous(), b.to(dev)
    out = torch.empty(n * T_out * V, N, dtype=dtype, device=dev)
    ops.conv_gemm(xc, wd, N, out, n_samples=n, T_in=T, T_out=T_out, Vin=V, bias=bd, taps=k, tap_step=dil, tap_off=-pad, t_mul=stride)
    close(from_cl(out, n, T_out, V), ref, dtype, "conv fwd")
    gyc = to_cl(gy).to(dtype).to(dev)
    dx = torch.empty(n * T * V, K, dtype=dtype, device=dev)
    # dx[t] = sum_tap dy[(t + pad - tap*dil)/stride] W[:, :, tap]^T
    ops.conv_gemm(gyc, wd, K, dx, n_samples=n, T_in=T_out, T_out=T, Vin=V, ws=(k, K * k, 1), taps=k, tap_step=-dil, tap_off=pad, t_div=stride)
    close(from_cl(dx, n, T, V), xr.grad, dtype, "conv dgrad")
    dW, db = torch.zeros_like(wd), torch.zeros_like(bd)
    ops.conv_wgrad(xc, gyc, dW, db=db, n_samples=n, T_in=T, T_out=T_out, Vin=V, taps=k, tap_step=dil, tap_off=-pad, t_mul=stride)
    close(dW, wr.grad, dtype, "conv wgrad")
    close(db, br.grad, dtype, "conv bgrad")


@pytest.mark.parametrize("K,N", [(64, 64), (24, 64), (64, 24), (176, 64), (96, 256), (128, 352), (256, 256), (128, 128)])
@pytest.mark.parametrize("variant", ["fwd", "bwd", "bwd_same", "dx", "ext_in", "contract"])
def test_conv_gemm_fast_engines(dev, K, N, variant):
    """bf16 shapes of the network (channels % 8 == 0) take the tcgen05 engines: the persistent ping-pong engine (several
    tiles per CTA, both accumulator phases, one or two K passes, narrow and multiple column tiles, per-pass weight copies)
    or the one-tile-per-CTA engine for wide K.  Every fused prologue / epilogue feature against plain PyTorch."""
    dtype = torch.bfloat16
    big = dev.type == "cuda"
    torch.manual_seed(K * 1000 + N)
    V = 25
    n, T = (96, 52) if big else (2, 3)                  # 124 800 rows = ~1000 tiles on the GPU; tiny on the simulator
    if not big and (K > 64 or N > 64):
        pytest.skip("simulator: small shapes only")
    ext_in, cext = variant == "ext_in", variant == "contract"
    Vin = V + 1 if cext else V
    tc4_before = _lib.lib().dsg_debug_counter(0)
    rows_in = n * T * Vin
    rows_out = n * T * (V + 1 if ext_in else V)
    x = rnd(rows_in, K, dev=dev, dtype=dtype)
    W, b = rnd(N, K, dev=dev, scale=0.2), rnd(N, dev=dev)
    a1, b1 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
    out = torch.empty(rows_out, N, dtype=dtype, device=dev)
    kw = dict(n_samples=n, T_in=T, T_out=T, Vin=Vin)
    if variant == "fwd":
        ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
        add = rnd(rows_out, N, dev=dev, dtype=dtype)
        ops.conv_gemm(ops.Act(x, a1, b1, relu=True), W, N, out, bias=b, add=add, stat_sum=ss, stat_sq=sq, **kw)
        ref = torch.relu(x.float() * a1 + b1) @ W.t() + b + add.float()
        close(ss, ref.sum(0), dtype, "sum")
        close(sq, (ref * ref).sum(0), dtype, "sumsq")
    elif variant == "bwd":
        x2 = rnd(rows_in, K, dev=dev, dtype=dtype)
        a2, b2 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
        mk, partner = rnd(rows_out, N, dev=dev, dtype=dtype), rnd(rows_out, N, dev=dev, dtype=dtype)
        ma, mb = torch.rand(N, device=dev) + 0.5, rnd(N, dev=dev, scale=0.2)
        ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
        Wt = W.t().contiguous()                              # [K, N] read with ws = (1, N, 0): the data-gradient orientation
        ops.conv_gemm(ops.Act(x, a1, b1, x2, a2, b2), Wt, N, out, ws=(1, N, 0), mask=ops.Act(mk, ma, mb), stat_sum=ss, stat_sq=sq,
                      partner=partner, **kw)
        ref = (x.float() * a1 + b1 + x2.float() * a2 + b2) @ W.t()
        ref = ref * ((mk.float() * ma + mb) > 0)
        close(ss, ref.sum(0), dtype, "sum")
        close(sq, (ref * partner.float()).sum(0), dtype, "sumprod")
    elif variant == "bwd_same":
        # the mask source is also the partner (transform / temporal-conv data gradients): staged once in the tail ring
        x2 = rnd(rows_in, K, dev=dev, dtype=dtype)
        a2, b2 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
        mk = rnd(rows_out, N, dev=dev, dtype=dtype)
        ma, mb = torch.rand(N, device=dev) + 0.5, rnd(N, dev=dev, scale=0.2)
        ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
        Wt = W.t().contiguous()
        ops.conv_gemm(ops.Act(x, a1, b1, x2, a2, b2), Wt, N, out, ws=(1, N, 0), mask=ops.Act(mk, ma, mb), stat_sum=ss, stat_sq=sq,
                      partner=mk, **kw)
        ref = (x.float() * a1 + b1 + x2.float() * a2 + b2) @ W.t()
        ref = ref * ((mk.float() * ma + mb) > 0)
        close(ss, ref.sum(0), dtype, "sum")
        close(sq, (ref * mk.float()).sum(0), dtype, "sumprod")
    elif variant == "dx":
        # last data gradient of a spatial unit: two addends and the per-sample broadcast row (dgphgcn1_backward)
        x2 = rnd(rows_in, K, dev=dev, dtype=dtype)
        a2, b2 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
        add, add2 = rnd(rows_out, N, dev=dev, dtype=dtype), rnd(rows_out, N, dev=dev, dtype=dtype)
        bc = rnd(n, V, N, dev=dev)
        Wt = W.t().contiguous()
        ops.conv_gemm(ops.Act(x, a1, b1, x2, a2, b2), Wt, N, out, ws=(1, N, 0), add=add, add2=add2, bcast=bc, bcast_scale=0.25, **kw)
        ref = (x.float() * a1 + b1 + x2.float() * a2 + b2) @ W.t() + add.float() + add2.float()
        ref = ref + 0.25 * bc[:, None].expand(n, T, V, N).reshape(-1, N)
    elif variant == "ext_in":
        ops.conv_gemm(ops.Act(x, a1, b1, relu=True), W, N, out, bias=b, ext_in=True, **kw)
        h = torch.relu(x.float() * a1 + b1).to(dtype).float().view(n * T, V, K)       # the mean is taken over the staged bf16 rows
        h = torch.cat([h, h.mean(1, keepdim=True)], 1).reshape(-1, K)
        ref = h @ W.t() + b
    else:
        bc = rnd(n, V, N, dev=dev)
        ops.conv_gemm(x, W, N, out, contract_ext=True, bcast=bc, bcast_scale=0.5, **kw)
        y = (x.float() @ W.t()).view(n * T, V + 1, N)
        ref = (y[:, :V] + y[:, V:] / V).reshape(-1, N) + 0.5 * bc[:, None].expand(n, T, V, N).reshape(-1, N)
    close(out, ref, dtype, f"{variant} out")
    if big and os.environ.get("DSG_DISABLE_TC4") != "1":
        assert _lib.lib().dsg_debug_counter(0) == tc4_before + 1, "the TMA-fed engine (tc4) declined a network shape"


@pytest.mark.parametrize("K,N", [(64, 64), (24, 64), (64, 176), (256, 96), (128, 288), (256, 256)])
@pytest.mark.parametrize("ext_in", [False, True])
def test_conv_wgrad_fast_engine(dev, K, N, ext_in):
    """bf16 weight / bias gradient on the tcgen05 engine: batched operand staging, register bias sums, vector reductions,
    K and N tiling, the joint-mean row, both operands with fused prologues (A: BN + ReLU; B: two-tensor BN-backward form)."""
    dtype = torch.bfloat16
    big = dev.type == "cuda"
    torch.manual_seed(K * 7 + N)
    V = 25
    n, T = (64, 40) if big else (2, 3)
    if not big and (K > 64 or N > 64):
        pytest.skip("simulator: small shapes only")
    rows_in, rows_out = n * T * V, n * T * (V + 1 if ext_in else V)
    x = rnd(rows_in, K, dev=dev, dtype=dtype)
    a1, b1 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
    e, y = rnd(rows_out, N, dev=dev, dtype=dtype), rnd(rows_out, N, dev=dev, dtype=dtype)
    ca, cb, cc = torch.rand(N, device=dev) + 0.5, rnd(N, dev=dev, scale=0.3), rnd(N, dev=dev, scale=0.1)
    dW, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    ops.conv_wgrad(ops.Act(x, a1, b1, relu=True), ops.Act(e, ca, cc, y, cb), dW, db=db, n_samples=n, T_in=T, T_out=T, Vin=V, ext_in=ext_in)
    A = torch.relu(x.float() * a1 + b1)
    if ext_in:
        A = A.to(dtype).float().view(n * T, V, K)
        A = torch.cat([A, A.mean(1, keepdim=True)], 1).reshape(-1, K)
    B = e.float() * ca + cc + y.float() * cb
    close(dW, B.t() @ A, dtype, "dW")
    close(db, B.sum(0), dtype, "db")


@pytest.mark.parametrize("K,N", [(64, 64), (24, 64), (64, 24), (96, 256), (256, 96), (128, 352), (256, 256), (64, 88)])
@pytest.mark.parametrize("variant", ["plain", "plain1", "act", "ext", "stride2"])
def test_conv_wgrad_tma_engine(dev, K, N, variant):
    """The TMA-fed weight-gradient engine (tc4_wgrad.cuh) on the network's 1x1 shapes: e^T x / y^T x accumulators combined with the
    BatchNorm-backward coefficients after the reduction, ones-MMA column sums (db, the cc term), plain / BN+ReLU / joint-mean /
    strided-frame x operands, partial last tiles, K and N tiling."""
    dtype = torch.bfloat16
    big = dev.type == "cuda"
    if not big:
        pytest.skip("tcgen05 + TMA kernel: GPU only (the simulator build runs the CUDA-core engine, covered above)")
    torch.manual_seed(K * 13 + N)
    V = 25
    n, T = 48, 37                                        # 44 400 rows: the last 128-row tile is partial
    stride = 2 if variant == "stride2" else 1
    ext_in = variant == "ext"
    T_in = T * stride
    rows_in, rows_out = n * T_in * V, n * T * (V + 1 if ext_in else V)
    x = rnd(rows_in, K, dev=dev, dtype=dtype)
    a1, b1 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
    e, y = rnd(rows_out, N, dev=dev, dtype=dtype), rnd(rows_out, N, dev=dev, dtype=dtype)
    ca, cb, cc = torch.rand(N, device=dev) + 0.5, rnd(N, dev=dev, scale=0.3), rnd(N, dev=dev, scale=0.1)
    dW, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
    A_src = ops.Act(x, a1, b1, relu=True) if variant in ("act", "ext") else x
    B_src = e if variant == "plain1" else ops.Act(e, ca, cc, y, cb)
    before = _lib.lib().dsg_debug_counter(1)
    ops.conv_wgrad(A_src, B_src, dW, db=db, n_samples=n, T_in=T_in, T_out=T, Vin=V, ext_in=ext_in, t_mul=stride)
    A = torch.relu(x.float() * a1 + b1) if variant in ("act", "ext") else x.float()
    if ext_in:
        A = A.to(dtype).float().view(n * T, V, K)
        A = torch.cat([A, A.mean(1, keepdim=True)], 1).reshape(-1, K)
    if stride == 2:
        A = A.view(n, T_in, V, K)[:, ::2].reshape(-1, K)
    B = e.float() if variant == "plain1" else e.float() * ca + cc + y.float() * cb
    close(dW, B.t() @ A, dtype, "dW")
    close(db, B.sum(0), dtype, "db")
    if os.environ.get("DSG_DISABLE_TC4") != "1":
        assert _lib.lib().dsg_debug_counter(1) == before + 1, "the TMA-fed weight-gradient engine declined a network shape"


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("K,N", [(88, 3), (19, 5), (64, 8)])
def test_conv_gemm_skinny_output(dev, dtype, K, N):
    """N <= 8 (the data gradient towards the 3-channel network input, gcn.py:2165 backward) runs as a row stream."""
    torch.manual_seed(K + N)
    n, T, V = 3, 7, 25
    rows = n * T * V
    x, x2 = rnd(rows, K, dev=dev, dtype=dtype), rnd(rows, K, dev=dev, dtype=dtype)
    a1, b1 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
    a2, b2 = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.2)
    Wt = rnd(K, N, dev=dev, scale=0.3)                       # [K, N]: the transposed (data-gradient) orientation
    add, bc = rnd(rows, N, dev=dev, dtype=dtype), rnd(n, V, N, dev=dev)
    out = torch.empty(rows, N, dtype=dtype, device=dev)
    ops.conv_gemm(ops.Act(x, a1, b1, x2, a2, b2), Wt, N, out, n_samples=n, T_in=T, T_out=T, Vin=V, ws=(1, N, 0), add=add, bcast=bc,
                  bcast_scale=1.0 / T)
    ref = (x.float() * a1 + b1 + x2.float() * a2 + b2) @ Wt + add.float() + bc[:, None].expand(n, T, V, N).reshape(-1, N) / T
    close(out, ref, dtype, "skinny out")


@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_gemm_joint_mean_row(dev, dtype):
    """ext_in appends mean_v (tcn.py:409); contract_ext is its gradient; wgrad sees the extended rows."""
    torch.manual_seed(2)
    n, T, V, K, N = 3, 5, 25, 12, 20
    x = torch.randn(n, K, T, V)
    w = torch.randn(N, K) * 0.3
    xr, wr = x.clone().requires_grad_(), w.clone().requires_grad_()
    xe = torch.cat([xr, xr.mean(-1, keepdim=True)], -1)
    ref = torch.einsum("nctv,oc->notv", xe, wr)
    gy = torch.randn_like(ref)
    ref.backward(gy)
    xc, wd = to_cl(x).to(dtype).to(dev), w.to(dev)
    out = torch.empty(n * T * (V + 1), N, dtype=dtype, device=dev)
    ops.conv_gemm(xc, wd, N, out, n_samples=n, T_in=T, T_out=T, Vin=V, ext_in=True)
    close(from_cl(out, n, T, V + 1), ref, dtype, "ext fwd")
    gyc = to_cl(gy).to(dtype).to(dev)
    dx = torch.empty(n * T * V, K, dtype=dtype, device=dev)
    ops.conv_gemm(gyc, wd, K, dx, n_samples=n, T_in=T, T_out=T, Vin=V + 1, ws=(1, K, 0), contract_ext=True)
    close(from_cl(dx, n, T, V), xr.grad, dtype, "ext dgrad")
    dW = torch.zeros_like(wd)
    ops.conv_wgrad(xc, gyc, dW, n_samples=n, T_in=T, T_out=T, Vin=V, ext_in=True)
    close(dW, wr.grad, dtype, "ext wgrad")


@pytest.mark.parametrize("V", [25, 17])
@pytest.mark.parametrize("dtype", DTYPES)
def test_conv_gemm_contract_ext_bn_backward_form(dev, dtype, V):
    """The last backward GEMM of dgmstcn: dg = fold(ca*e + cb*y + cc) @ W with the joint-mean row folded back (tcn.py:409 backward),
    ReLU mask + BatchNorm-backward sums of the consumer in the epilogue.  The constant cc reaches every joint row (1 + 1/V) times.
    V = 25: frame slot = one warp (fold on the accumulator in the epilogue of the TMA engine); V = 17: operand-side fold."""
    torch.manual_seed(11)
    n, T, K, N = 3, 6, 64, 32
    rows_in, rows_out = n * T * (V + 1), n * T * V
    e, y = rnd(rows_in, K, dev=dev, dtype=dtype), rnd(rows_in, K, dev=dev, dtype=dtype)
    ca, cb, cc = torch.rand(K, device=dev) + 0.5, rnd(K, dev=dev, scale=0.3), rnd(K, dev=dev, scale=2.0)
    W = rnd(K, N, dev=dev, scale=0.2)                      # backward weight [K_in_of_forward? no: dB channels K -> dg channels N]
    mask, partner = rnd(rows_out, N, dev=dev, dtype=dtype), rnd(rows_out, N, dev=dev, dtype=dtype)
    out = torch.empty(rows_out, N, dtype=dtype, device=dev)
    ssum, ssq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
    # W is passed as the forward weight [K, N] read transposed (ws = (1, N, 0)): out = dy @ W
    ops.conv_gemm(ops.Act(e, ca, cc, y, cb), W, N, out, n_samples=n, T_in=T, T_out=T, Vin=V + 1, ws=(1, N, 0), contract_ext=True,
                  mask=mask, stat_sum=ssum, stat_sq=ssq, partner=partner)
    dy = (e.float() * ca + y.float() * cb + cc).view(n * T, V + 1, K)
    folded = (dy[:, :V] + dy[:, V:] / V).reshape(rows_out, K)
    ref = (folded @ W) * (mask.float() > 0)
    close(out, ref, dtype, "contract_ext dgrad")
    close(ssum, ref.double().sum(0), dtype, "sum e")
    close(ssq, (ref.double() * partner.double()).sum(0), dtype, "sum e*partner")


@pytest.mark.parametrize("shape", [(37, 48, 13), (128, 256, 60), (5, 7, 3)])
def test_head_ce_kernels(dev, shape):
    """dsg_head_ce_fwd / _bwd: fc_cls + cross-entropy + top-1 / top-5 against nn.functional (heads/simple_head.py:93-96,
    losses/cross_entropy_loss.py:77-80, top_k_accuracy)."""
    N, Cn, K = shape
    torch.manual_seed(N + K)
    x, W, b = torch.randn(N, Cn), torch.randn(K, Cn) * 0.2, torch.randn(K) * 0.1
    y = torch.randint(0, K, (N,))
    xr, Wr, br = x.clone().requires_grad_(), W.clone().requires_grad_(), b.clone().requires_grad_()
    lr = F.linear(xr, Wr, br)
    loss_r = F.cross_entropy(lr, y)
    (loss_r * 1.7).backward()
    pred = lr.detach().topk(min(5, K), dim=1).indices
    hit = pred.eq(y.view(-1, 1))
    d = lambda t: t.to(dev)
    logits, stats = ops.head_ce_fwd(d(x), d(W), d(b), d(y))
    close(logits, lr, torch.float32, "logits")
    means = stats.mean(0)
    close(means[0], loss_r, torch.float32, "cross-entropy")
    assert float(means[1]) == pytest.approx(float(hit[:, :1].any(1).float().mean()), abs=1e-6)
    assert float(means[2]) == pytest.approx(float(hit.any(1).float().mean()), abs=1e-6)
    dW, db = torch.zeros(K, Cn, device=dev), torch.zeros(K, device=dev)
    gscale = torch.tensor([1.7 / N], device=dev)
    dlogits, dx = ops.head_ce_bwd(logits, d(y), d(x), d(W), gscale, dW=dW, db=db)
    close(dx, xr.grad, torch.float32, "dpooled")
    close(dW, Wr.grad, torch.float32, "dW")
    close(db, br.grad, torch.float32, "db")
    # scores only (inference): no labels, no statistics
    logits2, none = ops.head_ce_fwd(d(x), d(W), d(b))
    assert none is None
    close(logits2, lr, torch.float32, "logits (no labels)")


def test_bn_finalize_matches_batch_norm(dev):
    torch.manual_seed(3)
    Cn, M = 37, 500
    y = torch.randn(M, Cn) * 2 + 0.7
    gamma, beta = torch.rand(Cn) + 0.5, torch.randn(Cn)
    rm, rv = torch.randn(Cn) * 0.1, torch.rand(Cn) + 0.5
    yr = y.clone().requires_grad_()
    g_, b_ = gamma.clone().requires_grad_(), beta.clone().requires_grad_()
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ref = F.batch_norm(yr, rm_ref, rv_ref, g_, b_, True, 0.1, 1e-5)
    e = torch.randn_like(ref)
    ref.backward(e)
    d = lambda t: t.to(dev)
    ss, sq = d(y.double().sum(0)), d((y.double() ** 2).sum(0))
    a, b, sm, si = (torch.empty(Cn, device=dev) for _ in range(4))
    rmd, rvd = d(rm.clone()), d(rv.clone())
    ops.bn_finalize([ops.bn_job(0, Cn, sum=ss, sq=sq, count=M, gamma=d(gamma), beta=d(beta), running_mean=rmd, running_var=rvd,
                                save_mean=sm, save_invstd=si, a=a, b=b)])
    close(y.to(dev) * a + b, ref, torch.float32, "bn fwd")
    close(rmd, rm_ref, torch.float32, "running_mean")
    close(rvd, rv_ref, torch.float32, "running_var")
    s1, s2 = d(e.double().sum(0)), d((e.double() * y.double()).sum(0))
    ca, cb, cc, dg, db = (torch.empty(Cn, device=dev) for _ in range(5))
    ops.bn_finalize([ops.bn_job(2, Cn, sum=s1, sq=s2, count=M, gamma=d(gamma), save_mean=sm, save_invstd=si, a=ca, b=cb, c=cc,
                                dgamma=dg, dbeta=db)])
    close(e.to(dev) * ca + y.to(dev) * cb + cc, yr.grad, torch.float32, "bn bwd dx")
    close(dg, g_.grad, torch.float32, "dgamma")
    close(db, b_.grad, torch.float32, "dbeta")
    # eval mode
    ops.bn_finalize([ops.bn_job(1, Cn, gamma=d(gamma), beta=d(beta), running_mean=d(rm), running_var=d(rv), a=a, b=b)])
    close(y.to(dev) * a + b, F.batch_norm(y, rm, rv, gamma, beta, False, 0.1, 1e-5), torch.float32, "bn eval")


@pytest.mark.parametrize("dtype", DTYPES)
def test_tmean_and_pointwise(dev, dtype):
    torch.manual_seed(4)
    n, T, V, Cn = 3, 9, 5, 70
    x = rnd(n * T * V, Cn, dev=dev, dtype=dtype)
    xm = ops.tmean(x, n, T, V)
    close(xm, x.float().reshape(n, T, V, Cn).mean(1), torch.float32, "tmean")
    xv = rnd(n * 11 * V, 72, dev=dev, dtype=dtype)[:, :64]            # 16-byte path: channel slice of a wider buffer, T % 4 != 0
    close(ops.tmean(xv, n, 11, V), xv.float().reshape(n, 11, V, 64).mean(1), torch.float32, "tmean vec")
    x2 = rnd(n * T * V, Cn, dev=dev, dtype=dtype)
    a1, b1 = torch.rand(Cn, device=dev) + 0.5, rnd(Cn, dev=dev)
    mk = rnd(n * T * V, Cn, dev=dev, dtype=dtype)
    partner = rnd(n * T * V, Cn, dev=dev)     # fp32 partner with a bf16 source is allowed
    out = torch.empty(n * T * V, Cn, dtype=torch.float32, device=dev)
    ss, sq = torch.zeros(Cn, dtype=torch.float64, device=dev), torch.zeros(Cn, dtype=torch.float64, device=dev)
    ops.pointwise(ops.Act(x, a1, b1, x2, relu=True), out, mask=mk, stat_sum=ss, stat_sq=sq, partner=partner)
    ref = torch.relu(x.float() * a1 + b1 + x2.float()) * (mk.float() > 0)
    close(out, ref, dtype, "pointwise")
    close(ss, ref.sum(0), dtype)
    close(sq, (ref * partner).sum(0), dtype)
    # same op with everything in the activation dtype and 16-byte aligned (C % 8 == 0): the vectorised path for bf16
    Cv = 72
    xv, x2v, mkv, pv = (rnd(n * T * V, Cv, dev=dev, dtype=dtype) for _ in range(4))
    av, bv, a2v, b2v = torch.rand(Cv, device=dev) + 0.5, rnd(Cv, dev=dev), torch.rand(Cv, device=dev) + 0.5, rnd(Cv, dev=dev)
    outv = torch.empty(n * T * V, Cv, dtype=dtype, device=dev)
    ss, sq = torch.zeros(Cv, dtype=torch.float64, device=dev), torch.zeros(Cv, dtype=torch.float64, device=dev)
    ops.pointwise(ops.Act(xv, av, bv, x2v, a2v, b2v, relu=True), outv, mask=ops.Act(mkv, av, bv), stat_sum=ss, stat_sq=sq, partner=pv)
    refv = torch.relu(xv.float() * av + bv + x2v.float() * a2v + b2v) * ((mkv.float() * av + bv) > 0)
    close(outv, refv, dtype, "pointwise vec")
    close(ss, refv.sum(0), dtype)
    close(sq, (refv * pv.float()).sum(0), dtype)
    ss.zero_(); sq.zero_()
    ops.pointwise(xv, None, stat_sum=ss, stat_sq=sq)
    close(ss, xv.float().sum(0), dtype)
    close(sq, (xv.float() ** 2).sum(0), dtype)


def _topo_setup(layout, R, n, cin, seed):
    torch.manual_seed(seed)
    V, _, _, nt, et = O.graph_tables(layout)
    sd = {"conv1.weight": torch.randn(2 * R, cin, 1, 1) * 0.3, "conv1.bias": torch.randn(2 * R) * 0.1,
          "conv2.weight": torch.randn(2 * R, cin, 1, 1) * 0.3, "conv2.bias": torch.randn(2 * R) * 0.1,
          "conv1_se.weight": torch.randn(5 * R, cin, 1, 1) * 0.3, "conv1_se.bias": torch.randn(5 * R) * 0.1,
          "edge_linears.weight": torch.randn(15 * R, R, 1, 1) * 0.3, "edge_linears.bias": torch.randn(15 * R) * 0.1,
          "alpha": torch.randn(3), "beta": torch.randn(3), "A": torch.randn(3, V, V) * 0.02 + 0.04}
    xm = torch.randn(n, cin, V)
    return V, nt, et, sd, xm


@pytest.mark.parametrize("layout,R", [("nturgb+d", 8), ("coco", 16), ("nturgb+d", 32), ("nturgb+d/scrambled", 8)])
@pytest.mark.parametrize("adyn_dtype", DTYPES)
def test_topology_fwd_bwd(dev, layout, R, adyn_dtype):
    """The reference's tables make the edge type a function of the two node types: the kernels then work per node type (typed path).
    "/scrambled": an edge-type table that is NOT of that form must take the per-pair path and still match the oracle."""
    n, cin = 3, 12
    scrambled = layout.endswith("/scrambled")
    V, nt, et, sd, xm = _topo_setup(layout.split("/")[0], R, n, cin, 5)
    if scrambled:
        et = np.random.RandomState(3).randint(0, 15, size=(V, V))
    for v in sd.values():
        v.requires_grad_()
    xmr = xm.clone().requires_grad_()
    ref = O.dgphgcn1_topology(xmr, sd, nt, et, R)            # [n,3,R,V,V]
    g = torch.randn_like(ref)
    ref.backward(g)
    d = lambda t: t.detach().to(dev).contiguous()
    Wcat = torch.cat([sd["conv1.weight"], sd["conv2.weight"], sd["conv1_se.weight"]])[:, :, 0, 0]
    bcat = torch.cat([sd["conv1.bias"], sd["conv2.bias"], sd["conv1_se.bias"]])
    xm_rows = d(xm.permute(0, 2, 1).reshape(n * V, cin))
    H = torch.empty(n * V, 9 * R, device=dev)
    ops.conv_gemm(xm_rows, d(Wcat), 9 * R, H, n_samples=n, T_in=1, T_out=1, Vin=V, bias=d(bcat))
    ntd = torch.tensor(nt, dtype=torch.int32, device=dev)
    etd = torch.tensor(np.asarray(et), dtype=torch.int32, device=dev).reshape(-1)
    adyn = torch.empty(n, V, V, 3 * R, dtype=adyn_dtype, device=dev)
    S = torch.empty(n, 3, V, V, device=dev)
    We, be = d(sd["edge_linears.weight"][:, :, 0, 0]), d(sd["edge_linears.bias"])
    common = (H, n, V, R, ntd, etd, d(sd["A"]), d(sd["alpha"]), d(sd["beta"]), We, be)
    ops.topology_fwd(*common, adyn, S)
    ref_l = ref.detach().permute(0, 3, 4, 1, 2).reshape(n, V, V, 3 * R)
    close(adyn, ref_l, adyn_dtype, "adyn")
    dadyn = d(g.permute(0, 3, 4, 1, 2).reshape(n, V, V, 3 * R))
    dH = torch.empty_like(H)
    dA, dal, dbe_, dWe, dbe = (torch.zeros_like(t) for t in (d(sd["A"]), d(sd["alpha"]), d(sd["beta"]), We, be))
    # bf16 compute mode = the caller asks for the bf16 copy of dH: hardware tanh and the per-node-type form of the edge-typed
    # linear (exact tanhf on the simulator); the fp32 outputs are compared either way
    dHb = torch.empty(H.shape, dtype=torch.bfloat16, device=dev) if adyn_dtype != torch.float32 else None
    gt = adyn_dtype if dev.type == "cuda" else torch.float32
    ops.topology_bwd(*common, S, dadyn, dH, dA, dal, dbe_, dWe, dbe, dH_bf16=dHb)
    close(dA, sd["A"].grad, gt, "dA")
    close(dal, sd["alpha"].grad, gt, "dalpha")
    close(dbe_, sd["beta"].grad, gt, "dbeta")
    close(dWe, sd["edge_linears.weight"].grad[:, :, 0, 0], gt, "dWe")
    close(dbe, sd["edge_linears.bias"].grad, gt, "dbe")
    if adyn_dtype != torch.float32:
        close(dHb, dH, adyn_dtype, "dH bf16 copy")
        Wc = torch.cat([sd["conv1.weight"], sd["conv2.weight"], sd["conv1_se.weight"]])[:, :, 0, 0].detach()
        close(dH.reshape(n, V, 9 * R).cpu() @ Wc, xmr.grad.permute(0, 2, 1), gt, "dxm (dH through the feature convolutions)")
        return
    # dH -> weight grads and dxm through the generic GEMMs
    dW = torch.zeros(9 * R, cin, device=dev)
    db = torch.zeros(9 * R, device=dev)
    ops.conv_wgrad(xm_rows, dH, dW, db=db, n_samples=n, T_in=1, T_out=1, Vin=V)
    close(dW[:2 * R], sd["conv1.weight"].grad[:, :, 0, 0], torch.float32, "dconv1")
    close(dW[2 * R:4 * R], sd["conv2.weight"].grad[:, :, 0, 0], torch.float32, "dconv2")
    close(dW[4 * R:], sd["conv1_se.weight"].grad[:, :, 0, 0], torch.float32, "dconv1_se")
    close(db[4 * R:], sd["conv1_se.bias"].grad, torch.float32, "dconv1_se bias")
    dxm = torch.empty(n * V, cin, device=dev)
    ops.conv_gemm(dH, d(Wcat), cin, dxm, n_samples=n, T_in=1, T_out=1, Vin=V, ws=(1, cin, 0))
    close(dxm.reshape(n, V, cin).permute(0, 2, 1), xmr.grad, torch.float32, "dxm")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("V,KC", [(25, 24), (17, 48), (25, 96)])
def test_graph_agg_dynamic(dev, dtype, V, KC):
    torch.manual_seed(6)
    n, T = 2, 10
    p = torch.randn(n, T, V, KC)
    adyn = torch.randn(n, V, V, KC) * 0.3
    a1, b1 = torch.rand(KC) + 0.5, torch.randn(KC) * 0.2
    pd, ad = p.to(dtype).to(dev).reshape(-1, KC), adyn.to(dtype).to(dev)
    pq, aq = pd.float().cpu().reshape(n, T, V, KC), ad.float().cpu()
    act = torch.relu(pq * a1 + b1)
    ref = torch.einsum("ntuc,nuwc->ntwc", act, aq)
    out = torch.empty(n * T * V, KC, dtype=dtype, device=dev)
    ops.graph_agg(ops.Act(pd, a1.to(dev), b1.to(dev), relu=True), out, mode=0, n_samples=n, T=T, V=V, KC=KC, adyn=ad)
    close(out.reshape(n, T, V, KC), ref, dtype, "agg fwd")
    # transposed + mask + BN-backward sums
    dy = torch.randn(n, T, V, KC).to(dtype).to(dev).reshape(-1, KC)
    ss, sq = torch.zeros(KC, dtype=torch.float64, device=dev), torch.zeros(KC, dtype=torch.float64, device=dev)
    e = torch.empty_like(dy)
    ops.graph_agg(dy, e, mode=1, n_samples=n, T=T, V=V, KC=KC, adyn=ad, mask=ops.Act(pd, a1.to(dev), b1.to(dev)),
                  stat_sum=ss, stat_sq=sq, partner=pd)
    dyq = dy.float().cpu().reshape(n, T, V, KC)
    ref_e = torch.einsum("ntwc,nuwc->ntuc", dyq, aq) * (act > 0)
    close(e.reshape(n, T, V, KC), ref_e, dtype, "agg dP")
    close(ss, ref_e.sum((0, 1, 2)), dtype)
    close(sq, (ref_e * pq).sum((0, 1, 2)), dtype)
    dadj = torch.empty(n, V, V, KC, device=dev)
    ops.graph_agg_dadj(ops.Act(pd, a1.to(dev), b1.to(dev), relu=True), dy, dadj, n_samples=n, T=T, V=V, KC=KC)
    close(dadj, torch.einsum("ntuc,ntwc->nuwc", act, dyq), dtype, "dadyn")


@pytest.mark.parametrize("dtype", DTYPES)
def test_graph_agg_static(dev, dtype):
    torch.manual_seed(7)
    n, T, V, Cn, K = 2, 6, 25, 40, 3
    x = torch.randn(n, T, V, K * Cn).to(dtype)
    A = torch.randn(K, V, V) * 0.3
    xq = x.float().reshape(n, T, V, K, Cn)
    ref = torch.einsum("ntukc,kuw->ntwc", xq, A)
    xd, Ad = x.to(dev).reshape(-1, K * Cn), A.to(dev)
    out = torch.empty(n * T * V, Cn, dtype=dtype, device=dev)
    ops.graph_agg(xd, out, mode=2, n_samples=n, T=T, V=V, KC=Cn, A=Ad, Ksub=K)
    close(out.reshape(n, T, V, Cn), ref, dtype, "static agg")
    dy = torch.randn(n, T, V, Cn).to(dtype)
    dyd = dy.to(dev).reshape(-1, Cn)
    dx = torch.empty(n * T * V, K * Cn, dtype=dtype, device=dev)
    ops.graph_agg(dyd, dx, mode=3, n_samples=n, T=T, V=V, KC=Cn, A=Ad, Ksub=K)
    close(dx.reshape(n, T, V, K, Cn), torch.einsum("ntwc,kuw->ntukc", dy.float(), A), dtype, "static agg dx")
    dA = torch.zeros(K, V, V, device=dev)
    ops.graph_agg_dadj(xd, dyd, dA, n_samples=n, T=T, V=V, KC=Cn, is_static=True, Ksub=K)
    close(dA, torch.einsum("ntukc,ntwc->kuw", xq, dy.float()), dtype, "static dA")


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("stride,has_ext", [(1, True), (2, True), (1, False), (2, False)])
def test_ms_combine(dev, dtype, stride, has_ext):
    torch.manual_seed(8)
    n, T, V, Cn = 2, 9, 5, 30
    Vp = V + int(has_ext)
    T_out = (T - 1) // stride + 1
    ranges = ((0, 18), (18, 24), (24, 30))
    B = torch.randn(n, T, Vp, Cn).to(dtype).float().requires_grad_()
    Oc = torch.randn(n, T_out, Vp, Cn).to(dtype).float().requires_grad_()
    a1 = torch.cat([torch.rand(24) + 0.5, torch.ones(6)])
    b1 = torch.cat([torch.randn(24) * 0.3, torch.zeros(6)])
    addc = torch.randn(V).requires_grad_()
    h = B * a1 + b1
    mx = F.max_pool2d(torch.relu(h[..., 18:24]).permute(0, 3, 1, 2), (3, 1), (stride, 1), (1, 0)).permute(0, 2, 3, 1)
    ps = h[:, ::stride, :, 24:]
    o_all = torch.cat([Oc[..., :18], mx, ps], -1)
    feat_ref = o_all[:, :, :V] + (o_all[:, :, V:] * addc[None, None, :, None] if has_ext else 0)
    gy = torch.randn_like(feat_ref).to(dtype).float()
    feat_ref.backward(gy)
    d = lambda t: t.detach().to(dtype).to(dev)
    Bd, Od = d(B).reshape(-1, Cn), d(Oc).reshape(-1, Cn)
    feat = torch.empty(n * T_out * V, Cn, dtype=dtype, device=dev)
    oglob = torch.zeros(n * T_out, Cn, device=dev)
    ss, sq = torch.zeros(Cn, dtype=torch.float64, device=dev), torch.zeros(Cn, dtype=torch.float64, device=dev)
    bact = ops.Act(Bd, a1.to(dev), b1.to(dev))
    kw = dict(n=n, T_in=T, T_out=T_out, stride=stride, V=V, has_ext=has_ext, ranges=ranges, add_coeff=addc.detach().to(dev))
    ops.ms_combine_fwd(bact, Od, feat, oglob, stat_sum=ss, stat_sq=sq, **kw)
    close(feat.reshape(n, T_out, V, Cn), feat_ref, dtype, "feat")
    close(ss, feat_ref.sum((0, 1, 2)), dtype)
    close(sq, (feat_ref ** 2).sum((0, 1, 2)), dtype)
    d_o = torch.zeros(n * T_out * Vp, Cn, dtype=dtype, device=dev)
    e = torch.zeros(n * T * Vp, Cn, dtype=dtype, device=dev)
    es, eq = torch.zeros(Cn, dtype=torch.float64, device=dev), torch.zeros(Cn, dtype=torch.float64, device=dev)
    dadd = torch.zeros(V, device=dev)
    ops.ms_combine_bwd(bact, d(gy).reshape(-1, Cn), d_o, e, oglob, Bd, e_sum=es, e_sq=eq, dadd_coeff=dadd, **kw)
    close(d_o.reshape(n, T_out, Vp, Cn)[..., :18], Oc.grad[..., :18], dtype, "d_o")
    # E holds d/d(h) for max (masked by relu) and pass ranges: dB = E * a1
    close(e.reshape(n, T, Vp, Cn)[..., 18:] * a1[18:].to(dev), B.grad[..., 18:], dtype, "E")
    eref = (B.grad / a1)[..., 18:24]
    close(es[18:24], eref.sum((0, 1, 2)), dtype)
    close(eq[18:24], (eref * B.detach()[..., 18:24]).sum((0, 1, 2)), dtype)
    if has_ext:
        close(dadd, addc.grad, dtype, "dadd_coeff")


@pytest.mark.parametrize("stride", [1, 2])
@pytest.mark.parametrize("has_ext", [False, True])
def test_ms_mix_vector_kernels(dev, stride, has_ext):
    """The 16-byte vector forms of the streaming halves (ms_mix.cuh, bf16): channel ranges that do not fall on chunk boundaries
    (conv | max | pass = 44 | 10 | 10 like a 64-channel dgmstcn), the BatchNorm-backward two-tensor dfeat, full-width dO, and the
    per-input-frame pass run AFTER something else has written the conv range (read-modify-write of shared chunks)."""
    torch.manual_seed(11 + stride)
    dtype = torch.bfloat16
    n, T, V, Cn = 3, 9, 25, 64
    Vp = V + int(has_ext)
    T_out = (T - 1) // stride + 1
    ranges = ((0, 44), (44, 54), (54, 64))
    B = torch.randn(n, T, Vp, Cn).to(dtype).float().requires_grad_()
    Oc = torch.randn(n, T_out, Vp, 48).to(dtype).float().requires_grad_()
    a1 = torch.cat([torch.rand(54) + 0.5, torch.ones(10)])
    b1 = torch.cat([torch.randn(54) * 0.3, torch.zeros(10)])
    addc = torch.randn(V).requires_grad_()
    h = B * a1 + b1
    mx = F.max_pool2d(torch.relu(h[..., 44:54]).permute(0, 3, 1, 2), (3, 1), (stride, 1), (1, 0)).permute(0, 2, 3, 1)
    ps = h[:, ::stride, :, 54:]
    o_all = torch.cat([Oc[..., :44], mx, ps], -1)
    feat_ref = o_all[:, :, :V] + (o_all[:, :, V:] * addc[None, None, :, None] if has_ext else 0)
    # dfeat in the BatchNorm-backward form ca*e + cb*y + cc
    e2, y2 = torch.randn_like(feat_ref).to(dtype).float(), torch.randn_like(feat_ref).to(dtype).float()
    ca, cb, cc = torch.rand(Cn) + 0.5, torch.randn(Cn) * 0.2, torch.randn(Cn) * 0.1
    gy = e2 * ca + y2 * cb + cc
    feat_ref.backward(gy)
    d = lambda t: t.detach().to(dtype).to(dev)
    Bd, Od = d(B).reshape(-1, Cn), d(Oc).reshape(-1, 48)
    feat = torch.empty(n * T_out * V, Cn, dtype=dtype, device=dev)
    oglob = torch.zeros(n * T_out, Cn, device=dev)
    ss, sq = torch.zeros(Cn, dtype=torch.float64, device=dev), torch.zeros(Cn, dtype=torch.float64, device=dev)
    bact = ops.Act(Bd, a1.to(dev), b1.to(dev))
    kw = dict(n=n, T_in=T, T_out=T_out, stride=stride, V=V, has_ext=has_ext, ranges=ranges, add_coeff=addc.detach().to(dev) if has_ext else None)
    ops.ms_combine_fwd(bact, Od, feat, oglob if has_ext else None, stat_sum=ss, stat_sq=sq, **kw)
    close(feat.reshape(n, T_out, V, Cn), feat_ref, dtype, "feat")
    fq = feat.float().cpu()
    close(ss, fq.sum(0), dtype)
    close(sq, (fq ** 2).sum(0), dtype)
    if has_ext:
        close(oglob.reshape(n, T_out, Cn), o_all[:, :, V], dtype, "oglob")
    d_o = torch.zeros(n * T_out * Vp, Cn, dtype=dtype, device=dev)
    e = torch.full((n * T * Vp, Cn), 3.0, dtype=dtype, device=dev)       # 3.0 stands for "written by the conv data gradient"
    es, eq = torch.zeros(Cn, dtype=torch.float64, device=dev), torch.zeros(Cn, dtype=torch.float64, device=dev)
    dadd = torch.zeros(V, device=dev)
    dfeat = ops.Act(d(e2).reshape(-1, Cn), ca.to(dev), cc.to(dev), d(y2).reshape(-1, Cn), cb.to(dev))
    bkw = dict(e_sum=es, e_sq=eq, dadd_coeff=dadd if has_ext else None, d_o_full=True, **kw)
    ops.ms_combine_bwd(bact, dfeat, d_o, e, oglob if has_ext else None, Bd, parts=1, **bkw)
    ops.ms_combine_bwd(bact, dfeat, d_o, e, oglob if has_ext else None, Bd, parts=2, **bkw)
    close(d_o.reshape(n, T_out, Vp, Cn)[..., :44], Oc.grad[..., :44], dtype, "d_o")
    er = e.reshape(n, T, Vp, Cn)
    assert torch.all(er[..., :44] == 3.0), "the conv range of a shared chunk must survive the max / pass pass"
    close(er[..., 44:] * a1[44:].to(dev), B.grad[..., 44:], dtype, "E")
    eref = (B.grad / a1)[..., 44:54]
    close(es[44:54], eref.sum((0, 1, 2)), dtype)
    close(eq[44:54], (eref * B.detach()[..., 44:54]).sum((0, 1, 2)), dtype)
    if has_ext:
        close(dadd, addc.grad, dtype, "dadd_coeff")


def test_sgd_step_matches_torch(dev):
    torch.manual_seed(9)
    p = torch.randn(1000)
    pr = p.clone().requires_grad_()
    opt = torch.optim.SGD([pr], lr=0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
    pd, buf = p.clone().to(dev), torch.zeros(1000, device=dev)
    for _ in range(3):
        g = torch.randn(1000)
        pr.grad = g.clone()
        opt.step()
        ops.sgd_step(pd, g.to(dev), buf, 0.1, 0.9, 5e-4, True)
    close(pd, pr, torch.float32, "sgd")


@pytest.mark.gpu
@pytest.mark.parametrize("Vr", [26, 18])
@pytest.mark.parametrize("widths", [(14, 10, 10, 10), (23, 21, 21, 21), (46, 42, 42, 42)])
@pytest.mark.parametrize("stride", [1, 2])
def test_ms_conv_wgrad_tma(widths, stride, Vr):
    """dsg_ms_conv_wgrad: weight / bias gradients of all dilated (3 x 1) conv branches in one TMA-fed tcgen05 launch (operands read
    MN-major from tap-shifted 4-D tensor-map loads; zero fill = zero padding), against autograd of F.conv2d."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    torch.manual_seed(sum(widths) + stride + Vr)
    n, T = 5, 21                                             # odd T: the last frames of a sample fill a partial tile
    T_out = (T - 1) // stride + 1
    layout, lo = [], 0
    for w, d in zip(widths, (1, 2, 3, 4)):
        layout.append(("conv", lo, lo + w, (3, d)))
        lo += w
    chw, Ct = (lo + 7) & ~7, lo + 2 * widths[-1]             # H holds the conv channels, d_o every branch output
    H = torch.randn(n * T * Vr, chw, device=dev).to(torch.bfloat16)
    d_o = torch.randn(n * T_out * Vr, Ct, device=dev).to(torch.bfloat16)
    wgrads = {j: (torch.zeros(w, w, 3, 1, device=dev), torch.zeros(w, device=dev)) for j, w in enumerate(widths)}
    c0 = _lib.lib().dsg_debug_counter(4)
    assert ops.ms_conv_wgrad(H, d_o, layout, wgrads, n=n, T_in=T, T_out=T_out, stride=stride, Vr=Vr)
    assert _lib.lib().dsg_debug_counter(4) == c0 + 1
    _lib.join_side()
    torch.cuda.synchronize()
    for j, (kind, lo, hi, (k, d)) in enumerate(layout):
        w = hi - lo
        x = H[:, lo:hi].float().view(n, T, Vr, w).permute(0, 3, 1, 2)
        g = d_o[:, lo:hi].float().view(n, T_out, Vr, w).permute(0, 3, 1, 2)
        wt = torch.zeros(w, w, 3, 1, device=dev, requires_grad=True)
        F.conv2d(x, wt, None, (stride, 1), (d, 0), (d, 1)).backward(g)
        close(wgrads[j][0], wt.grad, torch.float32, f"dW branch {j}")
        close(wgrads[j][1], g.sum((0, 2, 3)), torch.float32, f"db branch {j}")


@pytest.mark.gpu
@pytest.mark.parametrize("widths", [(14, 10, 10, 10), (23, 21, 21, 21), (46, 42, 42, 42)])
@pytest.mark.parametrize("stride", [1, 2])
def test_ms_conv_tap_shifted_tma(widths, stride):
    """dsg_ms_conv: all dilated (3 x 1) conv branches in one tcgen05 launch from tap-shifted 4-D TMA loads (zero fill = zero
    padding), forward and data gradient (ReLU mask + BN-backward sums in the epilogue), against F.conv2d / autograd."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    torch.manual_seed(sum(widths) + stride)
    n, T, Vr = 5, 21, 26                                     # odd T: the last frames of a sample fill a partial tile
    T_out = (T - 1) // stride + 1
    dil = (1, 2, 3, 4)
    layout, lo = [], 0
    for w, d in zip(widths, dil):
        layout.append(("conv", lo, lo + w, (3, d)))
        lo += w
    span, chw = lo, (lo + 7) & ~7
    Ct = chw + 16
    layout += [("max", span, span + 8, ("max", 3)), ("1x1", span + 8, Ct, "1x1")]
    Ws = {j: (torch.randn(w, w, 3, 1, device=dev) * 0.2, torch.randn(w, device=dev)) for j, w in enumerate(widths)}
    h = torch.relu(torch.randn(n, T, Vr, chw, device=dev)).to(torch.bfloat16)
    # ---- forward
    O = torch.full((n * T_out * Vr, chw), float("nan"), dtype=torch.bfloat16, device=dev)
    before = _lib.lib().dsg_debug_counter(3)
    assert ops.ms_conv(h.view(-1, chw), O, layout, Ws, n=n, T_in=T, T_out=T_out, stride=stride, Vr=Vr, transposed=False)
    assert _lib.lib().dsg_debug_counter(3) == before + 1
    hf = h.float().permute(0, 3, 1, 2).requires_grad_()      # [n, C, T, V]
    refs = []
    for j, (w, d) in enumerate(zip(widths, dil)):
        lo_j = layout[j][1]
        refs.append(F.conv2d(hf[:, lo_j:lo_j + w], Ws[j][0], Ws[j][1], (stride, 1), (d, 0), (d, 1)))
    ref = torch.cat(refs, 1)                                 # [n, span, T_out, V]
    got = O.view(n, T_out, Vr, chw)[..., :span].permute(0, 3, 1, 2)
    close(got, ref, torch.bfloat16, "ms_conv forward")
    # ---- data gradient with mask, partner and statistics
    Cfull = Ct
    d_o = torch.randn(n, T_out, Vr, chw, device=dev).to(torch.bfloat16)
    Braw = torch.randn(n * T * Vr, Cfull, device=dev).to(torch.bfloat16)
    E = torch.full((n * T * Vr, Cfull), 7.0, dtype=torch.bfloat16, device=dev)
    ss, sq = torch.zeros(Cfull, dtype=torch.float64, device=dev), torch.zeros(Cfull, dtype=torch.float64, device=dev)
    assert ops.ms_conv(d_o.view(-1, chw), E, layout, Ws, n=n, T_in=T, T_out=T_out, stride=stride, Vr=Vr, transposed=True,
                       mask=h.view(-1, chw), partner=Braw, stat_sum=ss, stat_sq=sq)
    ref.backward(d_o.float()[..., :span].permute(0, 3, 1, 2))
    dref = hf.grad.permute(0, 2, 3, 1)[..., :span] * (h.float()[..., :span] > 0)           # [n, T, V, span]
    gotE = E.view(n, T, Vr, Cfull)
    close(gotE[..., :span], dref, torch.bfloat16, "ms_conv data gradient")
    assert torch.all(gotE[..., chw:] == 7.0), "channels outside the conv span's 16-byte chunks must not be written"
    assert torch.all(gotE[..., span:chw] == 0.0)             # the last chunk is padded with zeros (the caller writes those channels afterwards)
    dq = gotE[..., :span].float().reshape(-1, span)
    close(ss[:span], dq.sum(0), torch.bfloat16, "sum e")
    close(sq[:span], (dq * Braw.float()[:, :span]).sum(0), torch.bfloat16, "sum e*b")
    assert float(ss[span:].abs().sum()) == 0.0


@pytest.mark.gpu
@pytest.mark.parametrize("V,KC,N,T", [(25, 24, 64, 100), (25, 48, 128, 50), (17, 24, 64, 30), (25, 48, 128, 7)])
@pytest.mark.parametrize("store_y", [False, True])
def test_conv_gemm_fused_adjacency_contraction(V, KC, N, T, store_y):
    """North-star kernel (a): the per-sample, per-channel adjacency contraction as the operand producer of the post 1x1 convolution
    (gcn.py:2350-2363) — relu(bn(pre)) rows -> y[t,w,c] = sum_u p[t,u,c] adyn[u,w,c] -> z = y W^T + b with BatchNorm statistics, one
    kernel, against einsum + matmul.  T not a multiple of the 4-frame tile exercises the per-sample tail tiles."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    dtype = torch.bfloat16
    torch.manual_seed(V * 100 + KC + T)
    n = 37                                                   # several samples per CTA and CTAs that start mid-sample
    rows = n * T * V
    PD = rnd(rows, KC + 40, dev=dev, dtype=dtype)            # the contraction reads a channel slice of a wider buffer (pre | down)
    a1, b1 = torch.rand(KC, device=dev) + 0.5, rnd(KC, dev=dev, scale=0.2)
    adyn = (torch.randn(n, V, V, KC, device=dev) * 0.3).to(dtype)
    W, b = rnd(N, KC, dev=dev, scale=0.2), rnd(N, dev=dev)
    Z = torch.full((rows, N), float("nan"), dtype=dtype, device=dev)
    Y = torch.full((rows, KC), float("nan"), dtype=dtype, device=dev) if store_y else None
    ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
    before = _lib.lib().dsg_debug_counter(2)
    ops.conv_gemm(ops.Act(PD[:, :KC], a1, b1, relu=True), W, N, Z, n_samples=n, T_in=T, T_out=T, Vin=V, bias=b, stat_sum=ss, stat_sq=sq,
                  adyn=adyn, y_out=Y)
    assert _lib.lib().dsg_debug_counter(2) == before + 1
    p = torch.relu(PD[:, :KC].float() * a1 + b1).view(n, T, V, KC)
    y_ref = torch.einsum("ntuc,nuwc->ntwc", p, adyn.float())
    z_ref = y_ref.to(dtype).float().reshape(rows, KC) @ W.t() + b        # the tensor core multiplies the bf16-rounded y
    close(Z, z_ref, dtype, "fused z")
    zq = Z.float()
    close(ss, zq.sum(0), dtype, "sum z")
    close(sq, (zq * zq).sum(0), dtype, "sum z^2")
    if store_y:
        close(Y.view(n, T, V, KC), y_ref, dtype, "side output y")
