"""Input stage: dsgcn_b200.pipeline (torch, any device) against the numpy oracle, the oracle against the
unmodified reference classes (extracted from pose_related.py by name — the module itself imports half of the repository), the
GCN head + loss glue against the oracle, and the checkpoint writer / runner."""
import ast
import os

import numpy as np
import pytest
import torch

import dsgcn_b200
from dsgcn_b200 import pipeline as P
from oracle import pipeline_oracle as PO

REF = "/root/reference/pyskl/datasets/pipelines/pose_related.py"
CASES = [("nturgb+d", 25, 3, False), ("coco", 17, 2, True), ("coco", 17, 3, False), ("openpose", 18, 2, True)]
FEATS = [("j",), ("b",), ("jm",), ("bm",), ("j", "b", "jm", "bm"), ("b", "j")]


def _data(V, C, score, seed=0):
    rng = np.random.RandomState(seed)
    kp = rng.randn(2, 9, V, C).astype(np.float32)
    sc = rng.rand(2, 9, V).astype(np.float32) if score else None
    return kp, sc


def _reference_classes():
    """JointToBone / ToMotion / MergeSkeFeat / GenSkeFeat / FormatGCNInput compiled from the reference source, nothing else."""
    src = open(REF).read()
    want = {"JointToBone", "ToMotion", "MergeSkeFeat", "GenSkeFeat", "FormatGCNInput"}
    tree = ast.parse(src)
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in want]
    for n in body:
        n.decorator_list = []

    class Rename:
        def __init__(self, mapping):
            self.mapping = mapping

        def __call__(self, results):
            for k, v in self.mapping.items():
                if k in results:
                    results[v] = results.pop(k)
            return results

    class Compose:
        def __init__(self, ops):
            self.ops = ops

        def __call__(self, results):
            for op in self.ops:
                results = op(results)
            return results
    ns = dict(np=np, Rename=Rename, Compose=Compose)
    exec(compile(ast.Module(body=body, type_ignores=[]), REF, "exec"), ns)
    return ns


@pytest.mark.parametrize("dataset,V,C,score", CASES)
@pytest.mark.parametrize("feats", FEATS)
def test_gen_ske_feat_vs_oracle(dataset, V, C, score, feats):
    kp, sc = _data(V, C, score)
    ref = PO.gen_ske_feat(kp.copy(), None if sc is None else sc.copy(), dataset, list(feats))
    got = P.gen_ske_feat(torch.from_numpy(kp), None if sc is None else torch.from_numpy(sc), dataset, feats)
    assert got.shape == ref.shape
    assert np.array_equal(got.numpy(), ref)                  # differences / halves of fp32 values: bit-exact
    # batched (leading clip dimension), as the device-side stage uses it
    gb = P.gen_ske_feat(torch.from_numpy(np.stack([kp, kp * 2])), None if sc is None else torch.from_numpy(np.stack([sc, sc])), dataset, feats)
    assert np.array_equal(gb[0].numpy(), ref)


@pytest.mark.skipif(not os.path.isfile(REF), reason="reference tree not present")
@pytest.mark.parametrize("dataset,V,C,score", CASES)
def test_oracle_vs_live_reference_pipeline(dataset, V, C, score):
    ns = _reference_classes()
    kp, sc = _data(V, C, score, seed=1)
    for feats in FEATS:
        results = dict(keypoint=kp.copy())
        if sc is not None:
            results["keypoint_score"] = sc.copy()
        out = ns["GenSkeFeat"](dataset=dataset, feats=list(feats))(results)["keypoint"]
        assert np.array_equal(out, PO.gen_ske_feat(kp.copy(), None if sc is None else sc.copy(), dataset, list(feats))), (dataset, feats)
    for M, num_person, mode, nc in [(1, 2, "zero", 1), (1, 2, "loop", 3), (2, 2, "zero", 3), (3, 2, "zero", 1)]:
        k = np.random.RandomState(M).randn(M, 9, V, 3).astype(np.float32)
        out = ns["FormatGCNInput"](num_person=num_person, mode=mode)(dict(keypoint=k.copy(), num_clips=nc))["keypoint"]
        assert np.array_equal(out, PO.format_gcn_input(k.copy(), num_person, mode, nc))
        assert np.array_equal(out, P.format_gcn_input(torch.from_numpy(k), num_person, mode, nc).numpy())


@pytest.mark.gpu
def test_gen_ske_feat_on_device():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    for dataset, V, C, score in CASES:
        kp, sc = _data(V, C, score, seed=2)
        ref = PO.gen_ske_feat(kp.copy(), None if sc is None else sc.copy(), dataset, ["j", "b", "jm", "bm"])
        got = P.gen_ske_feat(torch.from_numpy(kp).cuda(), None if sc is None else torch.from_numpy(sc).cuda(), dataset, ("j", "b", "jm", "bm"))
        assert got.is_cuda and np.array_equal(got.cpu().numpy(), ref)


def test_checkpoint_layout_and_runner(tmp_path):
    """epoch_N.pth = {'meta', 'state_dict' (CPU tensors, no 'module.' prefix), 'optimizer'}; the runner calls the hook points in the
    reference's order and follows the cosine schedule per iteration."""
    from dsgcn_b200 import train as TR, parallel

    class M(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.fc = torch.nn.Linear(4, 3)

        def train_step(self, data_batch, optimizer, **kw):
            loss = self.fc(data_batch["x"]).square().mean()
            return dict(loss=loss, log_vars=dict(loss=float(loss.detach())), num_samples=len(data_batch["x"]))
    m = M()
    opt = torch.optim.SGD(m.parameters(), lr=0.1, momentum=0.9)
    calls, lrs = [], []

    class Hook:
        def before_run(self, r): calls.append("before_run")
        def before_train_epoch(self, r): calls.append("before_train_epoch")
        def before_train_iter(self, r): calls.append("before_train_iter"); lrs.append(r.optimizer.param_groups[0]["lr"])
        def after_train_iter(self, r): calls.append("after_train_iter")
        def after_train_epoch(self, r): calls.append("after_train_epoch")
        def after_run(self, r): calls.append("after_run")
    loader = [dict(x=torch.randn(5, 4)) for _ in range(3)]
    r = TR.Runner(m, opt, work_dir=str(tmp_path), max_epochs=2, hooks=[Hook()])
    r.run(loader)
    assert calls[:4] == ["before_run", "before_train_epoch", "before_train_iter", "after_train_iter"] and calls[-1] == "after_run"
    assert calls.count("before_train_iter") == 6
    assert lrs == pytest.approx([parallel.cosine_lr(0.1, i, 6) for i in range(6)])
    ck = torch.load(os.path.join(tmp_path, "epoch_2.pth"))
    assert set(ck) == {"meta", "state_dict", "optimizer"} and ck["meta"]["epoch"] == 2 and ck["meta"]["iter"] == 6
    assert list(ck["state_dict"]) == ["fc.weight", "fc.bias"] and not ck["state_dict"]["fc.weight"].is_cuda
    m2 = M()
    TR.load_checkpoint(m2, os.path.join(tmp_path, "epoch_2.pth"))
    assert torch.equal(m2.fc.weight, m.fc.weight)


def test_head_and_loss_vs_oracle():
    """GCNHead (mean over T,V then M, Linear), CrossEntropyLoss, top-k: recognizergcn.py:20-51, simple_head.py:83-97, base.py:50-84."""
    from oracle import dsgcn_oracle as O
    torch.manual_seed(0)
    head = dsgcn_b200.GCNHead(num_classes=60, in_channels=32)
    head.fc_cls.weight.data.normal_(0, 0.1)
    feat = torch.randn(6, 2, 32, 5, 25)
    label = torch.randint(0, 60, (6,))
    ref = O.gcn_head_forward(feat, {"fc_cls.weight": head.fc_cls.weight.detach(), "fc_cls.bias": head.fc_cls.bias.detach()})
    got = head(feat)
    assert torch.allclose(got, ref, atol=1e-5)
    losses = head.loss(got, label)
    assert torch.allclose(losses["loss_cls"], torch.nn.functional.cross_entropy(ref, label), atol=1e-6)
    top1 = (ref.argmax(1) == label).float().mean()
    assert torch.allclose(losses["top1_acc"], top1)
    top5 = (ref.topk(5, 1).indices == label[:, None]).any(1).float().mean()
    assert torch.allclose(losses["top5_acc"], top5)


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_head_pooling_kernel(dtype):
    """GCNHead on the device pools with the temporal-mean kernel (one pass over the channels-last backbone output, fp32 result);
    forward and backward against the reference formulation mean(T,V) -> mean(M) in fp32."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    N, M_, C, T, V = 5, 2, 256, 25, 25
    rows = torch.randn(N * M_ * T * V, C, device=dev).to(dtype)
    feat = rows.view(N * M_, T, V, C).permute(0, 3, 1, 2).reshape(N, M_, C, T, V).requires_grad_()     # the backbone's output layout
    head = dsgcn_b200.GCNHead(num_classes=60, in_channels=C).to(dev)
    out = head(feat)
    ref_in = feat.detach().float().requires_grad_()
    ref = head.fc_cls(ref_in.mean((3, 4)).mean(1))
    assert torch.allclose(out, ref, atol=2e-5, rtol=1e-4)
    g = torch.randn_like(out)
    out.backward(g)
    ref.backward(g)
    err = (feat.grad.float() - ref_in.grad).norm() / ref_in.grad.norm()
    assert err < (1e-5 if dtype == torch.float32 else 5e-3), float(err)


@pytest.mark.gpu
def test_graphed_train_step_matches_eager():
    """train.GraphedTrainStep (the whole iteration as one CUDA graph on static buffers) gives the same losses and parameters as the
    eager OptimizerHook sequence zero_grad / train_step / backward / step."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dsgcn_b200 import parallel, train as TR
    NORTH_STAR = dict(gcn_type="dgphgcn1", gcn_ratio=0.125, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True,
                      gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn",
                      graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
                      tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    sgd = dict(lr=0.01, momentum=0.9, weight_decay=5e-4, nesterov=True)
    models, opts = [], []
    for _ in range(2):
        torch.manual_seed(0); np.random.seed(0)
        m = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **{**NORTH_STAR, "base_channels": 16, "gcn_ratio": 0.5}),
                                     cls_head=dict(type="GCNHead", num_classes=10, in_channels=64)).to(dev).train()
        with torch.no_grad():
            for n_, p_ in m.named_parameters():
                if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                    p_.normal_(0, 0.1)
        models.append(m)
        opts.append(parallel.FlatSGD(parallel.GradBuckets(m, n_buckets=3), **sgd))
    models[1].load_state_dict(models[0].state_dict())
    g = torch.Generator().manual_seed(1)
    xs = [torch.randn(4, 1, 2, 16, 25, 3, generator=g).pin_memory() for _ in range(3)]
    ys = [torch.randint(0, 10, (4, 1), generator=g).pin_memory() for _ in range(3)]
    sd0 = {k: v.clone() for k, v in models[0].state_dict().items()}
    step = TR.GraphedTrainStep(models[1], opts[1], xs[0].to(dev), ys[0].to(dev), warmup=2)
    models[1].load_state_dict(sd0)                           # warm-up + capture advanced the weights: start both from the same point
    for bk in opts[1].gb.buckets:
        bk.flat_m.zero_()
    for i in range(3):
        opts[0].zero_grad()
        out = models[0].train_step(dict(keypoint=xs[i].to(dev), label=ys[i].to(dev)), opts[0])
        out["loss"].backward()
        opts[0].step()
        got = step(xs[i], ys[i])
        assert got["log_vars"]["loss"] == pytest.approx(out["log_vars"]["loss"], rel=2e-2), i
        assert set(got["log_vars"]) == set(out["log_vars"])
    p0, p1 = dict(models[0].named_parameters()), dict(models[1].named_parameters())
    num = sum(float((p0[k] - p1[k]).double().pow(2).sum()) for k in p0)
    den = sum(float(p0[k].double().pow(2).sum()) for k in p0)
    assert (num / den) ** 0.5 < 1e-2
    step.release()


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_fused_head_loss_matches_general_path(dtype):
    """RecognizerGCN.forward_train through the fused head kernels (dsg_head_ce_fwd/_bwd) = the general path (nn.Linear +
    F.cross_entropy + top-k, recognizergcn.py:20-51 / heads/base.py:50-84): same logged scalars, same gradients."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from dsgcn_b200 import modules as M
    dsgcn_b200._lib._testing_use_library(None)
    dev = torch.device("cuda:0")
    torch.manual_seed(0); np.random.seed(0)
    cfg = dict(gcn_type="dgphgcn1", gcn_ratio=0.25, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True,
               gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn", base_channels=16,
               graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
               tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])
    m = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **cfg), cls_head=dict(type="GCNHead", num_classes=12, in_channels=64)).to(dev).train()
    with torch.no_grad():
        m.cls_head.fc_cls.weight.normal_(0, 0.3)
    x = torch.randn(6, 1, 2, 16, 25, 3, device=dev)
    y = torch.randint(0, 12, (6, 1), device=dev)
    M.set_compute_dtype(dtype)
    try:
        res = {}
        for fused in (True, False):
            m.cls_head.use_fused = fused
            m.zero_grad(set_to_none=True)
            sd = {k: v.clone() for k, v in m.state_dict().items()}
            out = m.train_step(dict(keypoint=x, label=y), None)
            out["loss"].backward()
            res[fused] = (out["log_vars"], m.cls_head.fc_cls.weight.grad.clone(), m.cls_head.fc_cls.bias.grad.clone(),
                          m.backbone.gcn[0].gcn.pre[0].weight.grad.clone())
            m.load_state_dict(sd)                       # BatchNorm buffers back: both passes see the same model
        for k in res[True][0]:
            assert res[True][0][k] == pytest.approx(res[False][0][k], rel=1e-5, abs=1e-6), k
        for a, b in zip(res[True][1:], res[False][1:]):
            assert float((a - b).norm() / (b.norm() + 1e-12)) < (1e-4 if dtype == torch.float32 else 2e-2)
    finally:
        m.cls_head.use_fused = True
        M.set_compute_dtype(torch.bfloat16)
