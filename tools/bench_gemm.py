"""Micro-benchmark of dsg_conv_gemm / dsg_conv_wgrad at the network's shapes (kernel work; bench.py is the number of record).
usage: python tools/bench_gemm.py [--cases fwd64,fwd128,...] [--reps 20] [--wgrad]
Each case runs `reps` times between CUDA events on inputs larger than L2 where possible; prints ms and algorithmic GB/s."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dsgcn_b200  # noqa: E402
from dsgcn_b200 import ops  # noqa: E402

CASES = {
    # name: (K, N, rows, variant)
    "fwd64": (64, 64, 640000, "fwd"), "pre64": (64, 24, 640000, "fwd"), "post64": (24, 64, 640000, "fwd"),
    "fwd128": (128, 128, 320000, "fwd"), "fwd256": (256, 256, 160000, "fwd"), "pd7": (128, 352, 320000, "fwd"),
    "act64": (64, 64, 640000, "act"), "act256": (256, 256, 160000, "act"),
    "ext64": (64, 64, 640000, "ext"), "ext256": (256, 256, 160000, "ext"),
    "bwd64": (64, 64, 640000, "bwd"), "bwd128": (128, 128, 320000, "bwd"), "bwd256": (256, 256, 160000, "bwd"),
    "cx64": (64, 64, 640000, "cx"), "cx256": (256, 256, 160000, "cx"),
    "dx64": (88, 64, 640000, "dx"),
}


def run(name, reps, wgrad):
    K, N, rows, var = CASES[name]
    dev = torch.device("cuda")
    V, T = 25, 100
    n = rows // (T * V)
    rows = n * T * V
    bf = torch.bfloat16
    g = torch.Generator(device="cuda").manual_seed(0)
    rnd = lambda *s, dt=bf: torch.randn(*s, device=dev, generator=g).to(dt)
    kw = dict(n_samples=n, T_in=T, T_out=T, Vin=V)
    W, b = rnd(N, K, dt=torch.float32) * 0.1, rnd(N, dt=torch.float32)
    ss, sq = torch.zeros(N, dtype=torch.float64, device=dev), torch.zeros(N, dtype=torch.float64, device=dev)
    a1, b1 = torch.rand(K, device=dev) + 0.5, rnd(K, dt=torch.float32) * 0.1
    if wgrad:
        x = rnd(rows, K)
        e, y = rnd(rows, N), rnd(rows, N)
        ca, cb, cc = torch.rand(N, device=dev) + 0.5, rnd(N, dt=torch.float32) * 0.1, rnd(N, dt=torch.float32) * 0.1
        dW, db = torch.zeros(N, K, device=dev), torch.zeros(N, device=dev)
        A = ops.Act(x, a1, b1, relu=True) if var == "act" else x
        fn = lambda: ops.conv_wgrad(A, ops.Act(e, ca, cc, y, cb), dW, db=db, **kw)
        nbytes = rows * (K + 2 * N) * 2
    elif var in ("fwd", "act"):
        x = rnd(rows, K)
        out = torch.empty(rows, N, dtype=bf, device=dev)
        src = ops.Act(x, a1, b1, relu=True) if var == "act" else x
        fn = lambda: ops.conv_gemm(src, W, N, out, bias=b, stat_sum=ss, stat_sq=sq, **kw)
        nbytes = rows * (K + N) * 2
    elif var == "ext":
        x = rnd(rows, K)
        out = torch.empty(n * T * (V + 1), N, dtype=bf, device=dev)
        fn = lambda: ops.conv_gemm(x, W, N, out, bias=b, ext_in=True, stat_sum=ss, stat_sq=sq, **kw)
        nbytes = rows * K * 2 + out.numel() * 2
    elif var == "bwd":
        e, y = rnd(rows, K), rnd(rows, K)
        ca, cb, cc = torch.rand(K, device=dev) + 0.5, rnd(K, dt=torch.float32) * 0.1, rnd(K, dt=torch.float32) * 0.1
        feat = rnd(rows, N)
        ma, mb = torch.rand(N, device=dev) + 0.5, rnd(N, dt=torch.float32) * 0.1
        out = torch.empty(rows, N, dtype=bf, device=dev)
        Wt = W.t().contiguous()
        fn = lambda: ops.conv_gemm(ops.Act(e, ca, cc, y, cb), Wt, N, out, ws=(1, N, 0), mask=ops.Act(feat, ma, mb), stat_sum=ss, stat_sq=sq,
                                   partner=feat, **kw)
        nbytes = rows * (2 * K + 2 * N) * 2
    elif var == "cx":
        rin = n * T * (V + 1)
        e, y = rnd(rin, K), rnd(rin, K)
        ca, cb, cc = torch.rand(K, device=dev) + 0.5, rnd(K, dt=torch.float32) * 0.1, rnd(K, dt=torch.float32) * 0.1
        out = torch.empty(rows, N, dtype=bf, device=dev)
        Wt = W.t().contiguous()
        kw2 = dict(n_samples=n, T_in=T, T_out=T, Vin=V + 1)
        fn = lambda: ops.conv_gemm(ops.Act(e, ca, cc, y, cb), Wt, N, out, ws=(1, N, 0), contract_ext=True, **kw2)
        nbytes = rin * 2 * K * 2 + rows * N * 2
    else:   # dx: [E|PD] -> dx with two addends and the per-sample broadcast
        e, y = rnd(rows, K), rnd(rows, K)
        ca, cb, cc = torch.rand(K, device=dev) + 0.5, rnd(K, dt=torch.float32) * 0.1, rnd(K, dt=torch.float32) * 0.1
        add, add2 = rnd(rows, N), rnd(rows, N)
        bc = rnd(n, V, N, dt=torch.float32)
        out = torch.empty(rows, N, dtype=bf, device=dev)
        Wt = W.t().contiguous()
        fn = lambda: ops.conv_gemm(ops.Act(e, ca, cc, y, cb), Wt, N, out, ws=(1, N, 0), add=add, add2=add2, bcast=bc, bcast_scale=0.01, **kw)
        nbytes = rows * (2 * K + 3 * N) * 2
    for _ in range(3):
        fn()
    ops.L.join_side()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    ops.L.join_side()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name:8s} {'wgrad' if wgrad else 'gemm '} K={K:3d} N={N:3d} rows={rows:7d}  {ms:7.4f} ms  {nbytes / ms / 1e6:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--cases", default=",".join(CASES))
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--wgrad", action="store_true")
    a = ap.parse_args()
    ops.L.side_enabled = False
    for c in a.cases.split(","):
        run(c, a.reps, a.wgrad)
