"""Generates tests/golden/*.npz|json by running the UNMODIFIED reference (imported from /root/reference through
oracle/ref_loader.py).  Run in the authoring container:  python tests/golden/make_golden.py
The fixtures pin the oracle restatement (tests/test_oracle.py) and the CUDA path (tests/test_backbone.py) on
machines where /root/reference is absent (the GPU box)."""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader as rl  # noqa: E402
from oracle import dsgcn_oracle as O  # noqa: E402

SMALL = dict(base_channels=16, gcn_ratio=0.25)   # small-width copy of the north-star backbone for fast fixtures


def main():
    ns = rl.load()
    torch.set_num_threads(4)
    only_new = "--only-new" in sys.argv       # keep committed fixtures byte-identical, add the missing ones
    _savez = np.savez_compressed

    def savez(path, **kw):
        if only_new and os.path.exists(path):
            return
        _savez(path, **kw)
    np.savez_compressed = savez
    # 1. graph tables (integers must be bit-exact)
    tabs = {}
    for layout in ("nturgb+d", "coco", "openpose"):
        for mode, kw in (("spatial", {}), ("stgcn_spatial", {}), ("stgcn_spatial", {"max_hop": 2}), ("binary_adj", {})):
            g = ns.Graph(layout=layout, mode=mode, **kw)
            tabs[f"{layout}|{mode}|{kw.get('max_hop', 1)}|A"] = g.A
        if layout != "openpose":
            g = ns.Graph(layout=layout, mode="spatial")
            tabs[f"{layout}|node_type"] = np.asarray(g.node_type, dtype=np.int64)
            tabs[f"{layout}|edge_type"] = g.edge_type
    np.random.seed(7)
    tabs["nturgb+d|random|seed7"] = ns.Graph(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02).A
    np.savez_compressed(os.path.join(HERE, "graph_tables.npz"), **tabs)

    # 2. state-dict contract of the north-star backbone
    torch.manual_seed(0); np.random.seed(0)
    full = ns.DGSTGCN(**rl.NORTH_STAR_BACKBONE)
    keys = {k: list(v.shape) for k, v in full.state_dict().items()}
    json.dump(dict(num_parameters=sum(p.numel() for p in full.parameters()), keys=keys),
              open(os.path.join(HERE, "dgstgcn_state_dict_keys.json"), "w"), indent=0)

    # 3. seeded forward/backward vectors of a small-width backbone (all branches live)
    torch.manual_seed(1); np.random.seed(1)
    cfg = dict(rl.NORTH_STAR_BACKBONE); cfg.update(SMALL)
    m = ns.DGSTGCN(**cfg)
    sd = m.state_dict(); O.randomize_state(sd, 3); m.load_state_dict(sd)
    x = torch.randn(2, 2, 16, 25, 3)
    out = {"x": x.numpy()}
    for k, v in m.state_dict().items():
        out["sd|" + k] = v.numpy().copy()
    m.eval()
    with torch.no_grad():
        out["y_eval"] = m(x).numpy()
    m.train()
    y = m(x)
    gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(5))
    y.backward(gy)
    out["y_train"], out["gy"] = y.detach().numpy(), gy.numpy()
    for k, p in m.named_parameters():
        if p.grad is None:
            out["nograd|" + k] = np.zeros(0)
        elif k.split(".")[-1] in ("alpha", "beta", "add_coeff", "A") or k.startswith("data_bn") or "edge_linears" in k or k.startswith("gcn.0.") or k.startswith("gcn.9.tcn.transform"):
            out["grad|" + k] = p.grad.numpy().copy()
    for k, v in m.state_dict().items():
        if k.endswith("running_mean") or k.endswith("running_var"):
            if k.startswith("gcn.0.") or k.startswith("data_bn") or k.startswith("gcn.4."):
                out["after|" + k] = v.numpy().copy()
    np.savez_compressed(os.path.join(HERE, "dgstgcn_small.npz"), **out)

    # 4. one full-width DGBlock (block 8 shape: 256->256, R=32) on a short clip
    torch.manual_seed(2)
    g = ns.Graph(layout="nturgb+d", mode="random")
    A = torch.tensor(g.A, dtype=torch.float32)
    blk = ns.DGBlock(256, 256, A, torch.tensor(g.edge_type, dtype=torch.float32), torch.tensor(g.node_type), 1,
                     gcn_type="dgphgcn1", gcn_ratio=0.125, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True,
                     gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn")
    sd = blk.state_dict(); O.randomize_state(sd, 4)
    for v in sd.values():          # big tensors are stored as fp16: make that lossless *before* running the reference
        if v.dim() >= 2 and v.numel() > 20000:
            v.copy_(v.half().float())
    blk.load_state_dict(sd)
    xb = torch.randn(2, 256, 6, 25)
    blk.eval()
    ob = {"x": xb.numpy()}
    with torch.no_grad():
        ob["y_eval"] = blk(xb).numpy()
    for k, v in blk.state_dict().items():
        ob["sd|" + k] = v.numpy().astype(np.float16 if v.dim() >= 2 and v.numel() > 20000 else v.numpy().dtype)
    np.savez_compressed(os.path.join(HERE, "dgblock_256.npz"), **ob)
    # 5. config 5: small-width ST-GCN++ (configs/stgcn++/STGCN++_model.py:1-9) and vanilla ST-GCN (k=9 unit_tcn), fwd + bwd
    for name, kw in (("stgcnpp", dict(gcn_adaptive="init", gcn_with_res=True, tcn_type="mstcn", graph_cfg=dict(layout="nturgb+d", mode="spatial"))),
                     ("stgcn", dict(graph_cfg=dict(layout="coco", mode="stgcn_spatial")))):
        torch.manual_seed(6)
        ms = ns.STGCN(base_channels=12, **kw)
        sd = ms.state_dict(); O.randomize_state(sd, 7); ms.load_state_dict(sd)
        V = ms.gcn[0].gcn.A.shape[-1]
        xs = torch.randn(2, 2, 12, V, 3)
        og = {"x": xs.numpy()}
        for k, v in ms.state_dict().items():
            og["sd|" + k] = v.numpy().copy()
        ms.eval()
        with torch.no_grad():
            og["y_eval"] = ms(xs).numpy()
        ms.train()
        y = ms(xs)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(8))
        y.backward(gy)
        og["y_train"], og["gy"] = y.detach().numpy(), gy.numpy()
        for k, p in ms.named_parameters():
            if p.grad is not None and (k.startswith("gcn.0.") or k.startswith("gcn.4.") or k.startswith("gcn.8.gcn") or k.startswith("data_bn")):
                og["grad|" + k] = p.grad.numpy().copy()
        np.savez_compressed(os.path.join(HERE, f"{name}_small.npz"), **og)
    # 6. config-5 variants on the same kernel library: the plain DG-GCN unit (dggcn + mstcn, the original DG-STGCN block),
    #    MSTCN (msg3d_utils.py) and CTR-GCN (unit_ctrgcn + MSTCN).  Small widths; forward (eval + train) and backward.
    def run_model(m, x, sel):
        og = {"x": x.numpy()}
        for k, v in m.state_dict().items():
            og["sd|" + k] = v.numpy().copy()
        m.eval()
        with torch.no_grad():
            og["y_eval"] = m(x).numpy()
        m.train()
        xr = x.clone().requires_grad_()
        y = m(xr)
        gy = torch.randn(y.shape, generator=torch.Generator().manual_seed(8))
        y.backward(gy)
        og["y_train"], og["gy"], og["gx"] = y.detach().numpy(), gy.numpy(), xr.grad.numpy()
        for k, p in m.named_parameters():
            if p.grad is not None and sel(k):
                og["grad|" + k] = p.grad.numpy().copy()
        return og

    def settle(m, x, iters=30):
        """run the reference in train mode so the running statistics describe the activations (eval mode then stays in range)"""
        m.train()
        with torch.no_grad():
            for _ in range(iters):
                m(x)

    torch.manual_seed(9); np.random.seed(9)
    g = ns.Graph(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02)
    A = torch.tensor(g.A, dtype=torch.float32)
    for name, kw in (("dggcn_block", dict(gcn_type="dggcn", gcn_ratio=0.25, tcn_type="mstcn")),
                     ("dggcn_block_sw", dict(gcn_type="dggcn", gcn_ratio=0.25, gcn_subset_wise=True, tcn_type="dgmstcn"))):
        blk = ns.DGBlock(16, 24, A.clone(), torch.tensor(g.edge_type, dtype=torch.float32), torch.tensor(g.node_type), 2, **kw)
        sd = blk.state_dict(); O.randomize_state(sd, 10); blk.load_state_dict(sd)
        xb = torch.randn(3, 16, 12, 25, generator=torch.Generator().manual_seed(11))
        settle(blk, xb)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), **run_model(blk, xb, lambda k: True))
    torch.manual_seed(12)
    ms = ns.MSTCN(20, 20, kernel_size=5, stride=2, dilations=[1, 2], residual=True)
    sd = ms.state_dict(); O.randomize_state(sd, 13); ms.load_state_dict(sd)
    xb = torch.randn(3, 20, 12, 25, generator=torch.Generator().manual_seed(14))
    settle(ms, xb)
    np.savez_compressed(os.path.join(HERE, "mstcn_msg3d.npz"), **run_model(ms, xb, lambda k: True))
    torch.manual_seed(15); np.random.seed(15)
    ct = ns.CTRGCN(graph_cfg=dict(layout="nturgb+d", mode="spatial"), base_channels=16, gcn_type="unit_ctrgcn")
    sd = ct.state_dict(); O.randomize_state(sd, 16)
    for k in sd:
        if k.endswith("gcn1.alpha"):
            sd[k] = sd[k] * 0.3
    ct.load_state_dict(sd)
    xs = torch.randn(2, 2, 12, 25, 3, generator=torch.Generator().manual_seed(17))
    settle(ct, xs)
    np.savez_compressed(os.path.join(HERE, "ctrgcn_small.npz"),
                        **run_model(ct, xs, lambda k: k.startswith(("net.0.", "net.4.", "net.9.gcn1", "data_bn"))))
    for f in os.listdir(HERE):
        print(f, os.path.getsize(os.path.join(HERE, f)))


if __name__ == "__main__":
    main()
