"""Run-to-run reproducibility probe: the same small DGSTGCN, same inputs, two forward/backward passes; prints the largest relative
difference of any parameter gradient between the passes (atomic accumulation order is the only legitimate source)."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import dsgcn_b200
from dsgcn_b200 import modules as M
NS = dict(gcn_type="dgphgcn1", gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True, gcn_subset_wise=True,
          gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn", graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
          tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])
dev = torch.device("cuda:0")
for dtype, bc, ratio in ((torch.float32, 16, 0.25), (torch.bfloat16, 16, 0.25), (torch.bfloat16, 64, 0.125)):
    torch.manual_seed(0); np.random.seed(0)
    M.set_compute_dtype(dtype)
    m = M.DGSTGCN(base_channels=bc, gcn_ratio=ratio, **NS).to(dev).train()
    with torch.no_grad():
        for n_, p_ in m.named_parameters():
            if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                p_.normal_(0, 0.5)
    x = torch.randn(6, 2, 24, 25, 3, device=dev)
    gy = None
    runs = []
    for r in range(3):
        m.zero_grad(set_to_none=True)
        y = m(x)
        if gy is None:
            gy = torch.randn_like(y)
        y.backward(gy)
        torch.cuda.synchronize()
        runs.append(({k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}, y.detach().clone()))
    gmax = max(float(g.norm()) for g in runs[0][0].values())
    worst = ("", 0.0)
    for k in runs[0][0]:
        if float(runs[0][0][k].norm()) < 1e-3 * gmax:       # (biases in front of a BatchNorm: true gradient zero, rounding noise only)
            continue
        for r in (1, 2):
            d = float((runs[r][0][k] - runs[0][0][k]).norm()) / (float(runs[0][0][k].norm()) + 1e-6 * gmax)
            if d > worst[1]:
                worst = (k, d)
    print(f"{dtype} base {bc}: forward diff {float((runs[1][1].float() - runs[0][1].float()).norm() / runs[0][1].float().norm()):.2e}, "
          f"worst gradient diff {worst[1]:.2e} at {worst[0]}")
