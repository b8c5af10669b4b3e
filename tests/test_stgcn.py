"""Config 5: ST-GCN++ (unit_gcn adaptive='init' + with_res, mstcn) and vanilla ST-GCN (unit_gcn + 9x1 unit_tcn)
backbones on the same kernel library — pyskl/models/gcns/stgcn.py:16-153.  Golden vectors come from the unmodified
reference (tests/golden/make_golden.py section 5); the oracle restatement is pinned on them too."""
import os

import numpy as np
import pytest
import torch

from dsgcn_b200 import modules as M
from oracle import dsgcn_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")
CASES = {
    "stgcnpp": (dict(gcn_adaptive="init", gcn_with_res=True, tcn_type="mstcn", graph_cfg=dict(layout="nturgb+d", mode="spatial")),
                dict(gcn_kw=dict(adaptive="init", with_res=True), tcn_type="mstcn")),
    "stgcn": (dict(graph_cfg=dict(layout="coco", mode="stgcn_spatial")), dict(gcn_kw=dict(adaptive="init"), tcn_type="unit_tcn")),
}


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _load(name):
    z = np.load(os.path.join(G, f"{name}_small.npz"))
    sd = {k[3:]: torch.from_numpy(z[k].copy()) for k in z.files if k.startswith("sd|")}
    return z, sd


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_stgcn_vs_golden(name):
    z, sd = _load(name)
    x = torch.from_numpy(z["x"])
    okw = CASES[name][1]
    y = O.stgcn_forward(x, {k: v.clone() for k, v in sd.items()}, training=False, base_channels=12, **okw)
    assert rel(y, torch.from_numpy(z["y_eval"])) < 1e-5
    sdt = {k: v.clone() for k, v in sd.items()}
    for k, v in sdt.items():
        if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
            v.requires_grad_()
    y = O.stgcn_forward(x, sdt, training=True, base_channels=12, **okw)
    assert rel(y, torch.from_numpy(z["y_train"])) < 1e-5
    y.backward(torch.from_numpy(z["gy"]))
    gmax = max(float(np.linalg.norm(z[k])) for k in z.files if k.startswith("grad|"))
    for k in z.files:
        if k.startswith("grad|"):
            g, r = sdt[k[5:]].grad, torch.from_numpy(z[k])
            assert (g - r).norm() <= 2e-3 * r.norm() + 1e-5 * gmax, k    # 10 layers deep, fp32 re-association + ReLU-mask flips


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("name", list(CASES))
def test_stgcn_vs_golden(dev, dtype, name):
    z, sd = _load(name)
    np.random.seed(0)
    m = M.STGCN(base_channels=12, **CASES[name][0])
    assert list(m.state_dict().keys()) == list(sd.keys())          # checkpoint drop-in
    m.load_state_dict(sd)
    m.to(dev)
    x = torch.from_numpy(z["x"]).to(dev)
    M.set_compute_dtype(dtype)
    try:
        m.eval()
        with torch.no_grad():
            y = m(x)
        assert y.shape == z["y_eval"].shape
        assert rel(y, torch.from_numpy(z["y_eval"])) < (1e-4 if dtype == torch.float32 else 1.5e-2)
        m.train()
        y = m(x)
        lim = 1e-4
        if dtype == torch.bfloat16:
            # tiny train-mode batches (300 values per BatchNorm channel in the last blocks) amplify bf16 rounding: calibrate on
            # what PyTorch's own bf16 autocast does to the oracle on the same inputs (diagnostic printed; see tests/test_parity_bf16.py
            # for the fixed 1e-2 bound at the benchmarked batch)
            with torch.no_grad(), torch.autocast("cpu", dtype=torch.bfloat16):
                ya = O.stgcn_forward(torch.from_numpy(z["x"]), {k: v.clone() for k, v in sd.items()}, training=True, base_channels=12,
                                     **CASES[name][1])
            ac = rel(ya.float(), torch.from_numpy(z["y_train"]))
            print(f"stgcn {name} bf16 train: ours {rel(y, torch.from_numpy(z['y_train'])):.3e}, torch autocast {ac:.3e}")
            lim = max(1.5e-2, 1.5 * ac)
        assert rel(y, torch.from_numpy(z["y_train"])) < lim
        y.backward(torch.from_numpy(z["gy"]).to(dev).to(y.dtype))
        params = dict(m.named_parameters())
        gmax = max(float(np.linalg.norm(z[k])) for k in z.files if k.startswith("grad|"))
        if dtype == torch.float32:
            for k in z.files:
                if k.startswith("grad|"):
                    g, r = params[k[5:]].grad.detach().cpu(), torch.from_numpy(z[k])
                    assert (g - r).norm() <= 3e-3 * r.norm() + 1e-4 * gmax, k
        else:
            # same calibration for the gradient direction (a 10-block, 12-channel toy net in train mode is chaotic in bf16: the
            # unit-level tests hold the per-unit bf16 gradients)
            keys = [k for k in z.files if k.startswith("grad|")]
            sda = {k: v.clone() for k, v in sd.items()}
            for k, v in sda.items():
                if v.is_floating_point() and not k.endswith(("running_mean", "running_var")):
                    v.requires_grad_()
            with torch.autocast("cpu", dtype=torch.bfloat16):
                ya = O.stgcn_forward(torch.from_numpy(z["x"]), sda, training=True, base_channels=12, **CASES[name][1])
            ya.float().backward(torch.from_numpy(z["gy"]))
            refv = torch.cat([torch.from_numpy(z[k]).double().reshape(-1) for k in keys])
            acv = torch.cat([sda[k[5:]].grad.double().reshape(-1) for k in keys])
            mine = torch.cat([params[k[5:]].grad.detach().double().cpu().reshape(-1) for k in keys])
            cos = lambda a, b: float(torch.dot(a, b) / (a.norm() * b.norm()))
            print(f"stgcn {name} bf16 gradient cosine: ours {cos(mine, refv):.3f}, torch autocast {cos(acv, refv):.3f}")
            assert cos(mine, refv) > min(0.95, cos(acv, refv) - 0.15)
    finally:
        M.set_compute_dtype(torch.bfloat16)
