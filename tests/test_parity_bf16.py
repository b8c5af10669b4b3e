"""bf16 parity in the mode that is benchmarked (north_star: "bf16 logits top-1 agreement and rel 1e-2").

RecognizerGCN (north-star DGSTGCN + GCNHead, bf16 kernels) against the CPU fp32 oracle of the reference
(recognizers/recognizergcn.py:20-51 -> gcns/dgstgcn.py:156-170 -> heads/simple_head.py:83-97) on the same seeded
weights and inputs, at BASELINE config 1 (N=16) and at the benchmark batch (128 clips/GPU), eval and train mode:

  * rel-L2 of the pooled 256-d feature and of the logits <= 1e-2  (FIXED bound; PyTorch's own bf16-autocast error on the
    oracle is printed as a diagnostic only, it is not part of the criterion);
  * logits top-1 agreement >= 99 % on the eval set (512 clips, so one flipped near-tie is 0.2 %); at 16 / 128 clips a
    single near-tie is 6 % / 0.8 %, so there the assertion is: every clip whose fp32 top-2 margin is decisive
    (> 4 x the largest logit error) agrees, and at most 1 / 2 clips differ overall.
Random-init heads make near-ties common (60 logits of std ~ 0.01 |feat|); a trained head only widens the margins.
"""
import numpy as np
import pytest
import torch

import dsgcn_b200
from dsgcn_b200 import modules as M
from oracle import dsgcn_oracle as O

pytestmark = pytest.mark.gpu

NORTH_STAR = dict(gcn_type="dgphgcn1", gcn_ratio=0.125, gcn_node_attention=True, gcn_edge_attention=True, gcn_decompose=True,
                  gcn_subset_wise=True, gcn_ctr="T", gcn_ada="T", tcn_type="dgmstcn",
                  graph_cfg=dict(layout="nturgb+d", mode="random", num_filter=3, init_off=.04, init_std=.02),
                  tcn_ms_cfg=[(3, 1), (3, 2), (3, 3), (3, 4), ("max", 3), "1x1"])


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-12))


def _model():
    torch.manual_seed(0)
    np.random.seed(0)
    m = dsgcn_b200.RecognizerGCN(backbone=dict(type="DGSTGCN", **NORTH_STAR), cls_head=dict(type="GCNHead", num_classes=60, in_channels=256))
    with torch.no_grad():           # live dynamic branches (zero at init: SURVEY.md §0 item 6), default BN affine/buffers as at init
        for n_, p in m.named_parameters():
            if n_.rsplit(".", 1)[-1] in ("alpha", "beta", "add_coeff"):
                p.normal_(0, 0.1)
    return m


def _oracle(m, x, training, chunk=64):
    sd = {k: v.detach().clone().cpu() for k, v in m.backbone.state_dict().items()}
    hd = {k: v.detach().clone().cpu() for k, v in m.cls_head.state_dict().items()}
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        if training:                 # batch statistics couple the clips: one pass
            feat = O.dgstgcn_forward(x, sd, training=True)
        else:
            feat = torch.cat([O.dgstgcn_forward(x[i:i + chunk], sd, training=False) for i in range(0, x.shape[0], chunk)])
        pooled = feat.mean((3, 4)).mean(1)
        return pooled, O.gcn_head_forward(feat, hd)


def _ours(m, x, training):
    dev = torch.device("cuda:0")
    dsgcn_b200._lib._testing_use_library(None)
    m.to(dev).train(training)
    M.set_compute_dtype(torch.bfloat16)
    with torch.no_grad():
        feat = m.extract_feat(x.to(dev))
        pooled = feat.float().mean((3, 4)).mean(1)
        return pooled.cpu(), m.cls_head(feat).float().cpu()


def _check(pooled, logits, p_ref, l_ref, max_flips, tag):
    e_p, e_l = rel(pooled, p_ref), rel(logits, l_ref)
    agree = (logits.argmax(1) == l_ref.argmax(1))
    top2 = l_ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    err_inf = float((logits - l_ref).abs().max())
    decisive = margin > 4 * err_inf
    print(f"[bf16 parity] {tag}: pooled rel-L2 {e_p:.3e}, logits rel-L2 {e_l:.3e}, top-1 agreement {float(agree.float().mean()):.4f} "
          f"({int((~agree).sum())} of {len(agree)} differ; {int(decisive.sum())} decisive, all agree: {bool(agree[decisive].all())})")
    assert e_p <= 1e-2, f"{tag}: pooled feature rel-L2 {e_p:.3e} > 1e-2"
    assert e_l <= 1e-2, f"{tag}: logits rel-L2 {e_l:.3e} > 1e-2"
    assert bool(agree[decisive].all()), f"{tag}: a clip with a decisive fp32 margin changed its top-1 class"
    assert int((~agree).sum()) <= max_flips, f"{tag}: {int((~agree).sum())} top-1 flips"
    return float(agree.float().mean())


@pytest.mark.parametrize("clips,training", [(16, False), (16, True), (128, True)])
def test_logits_parity_bf16(clips, training):
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m = _model()
    x = torch.randn(clips, 2, 100, 25, 3, generator=torch.Generator().manual_seed(clips))
    p_ref, l_ref = _oracle(m, x, training)
    pooled, logits = _ours(m, x, training)
    _check(pooled, logits, p_ref, l_ref, max_flips=1 if clips == 16 else 2, tag=f"{clips} clips, {'train' if training else 'eval'}")


def test_logits_top1_agreement_eval_512():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    m = _model()
    x = torch.randn(512, 2, 100, 25, 3, generator=torch.Generator().manual_seed(512))
    p_ref, l_ref = _oracle(m, x, False)
    pooled, logits = torch.cat([_ours(m, x[i:i + 128], False)[0] for i in range(0, 512, 128)]), None
    logits = torch.cat([_ours(m, x[i:i + 128], False)[1] for i in range(0, 512, 128)])
    a = _check(pooled, logits, p_ref, l_ref, max_flips=5, tag="512 clips, eval")
    assert a >= 0.99, f"top-1 agreement {a:.4f} < 0.99"
