"""Model glue with the reference's names and call signatures (callers of the hot path):
registry + build_model (pyskl/models/builder.py:5-39), GCNHead (heads/simple_head.py:12-140),
CrossEntropyLoss (losses/cross_entropy_loss.py:11-84), RecognizerGCN (recognizers/recognizergcn.py:16-128,
recognizers/base.py:21-205).  The backbone is the kernel path; the head is a 256->num_classes linear layer on
pooled features (torch ops on the device: negligible work).  Differences kept deliberately small and stated:
top-k accuracy is computed on the device (no per-iteration .cpu() sync, heads/base.py:66-72) and the logged scalars are
reduced / copied to the host together (one collective + one D2H per iteration instead of four of each, recognizers/base.py:151-156).
"""
from collections import OrderedDict

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

from . import modules, ops


class Registry:
    def __init__(self, name):
        self.name = name
        self.module_dict = {}

    def register_module(self, name=None, module=None):
        def deco(cls):
            self.module_dict[name or cls.__name__] = cls
            return cls
        if module is not None:
            return deco(module)
        return deco

    def get(self, key):
        return self.module_dict.get(key)

    def build(self, cfg):
        if not isinstance(cfg, dict) or "type" not in cfg:
            raise TypeError(f"cfg must be a dict with a 'type' key, got {cfg}")
        args = dict(cfg)
        t = args.pop("type")
        cls = self.module_dict.get(t) if isinstance(t, str) else t
        if cls is None:
            raise KeyError(f"{t} is not in the {self.name} registry (built on this path: {sorted(self.module_dict)})")
        return cls(**args)


MODELS = Registry("models")
BACKBONES = HEADS = RECOGNIZERS = LOSSES = NECKS = MODELS     # one registry under five names (builder.py:5-10)

MODELS.register_module(module=modules.DGSTGCN)
MODELS.register_module(module=modules.STGCN)
MODELS.register_module(module=modules.CTRGCN)


def build_backbone(cfg):
    return BACKBONES.build(cfg)


def build_head(cfg):
    return HEADS.build(cfg)


def build_recognizer(cfg):
    return RECOGNIZERS.build(cfg)


def build_loss(cfg):
    return LOSSES.build(cfg)


def build_model(cfg):
    """builder.py:33-39"""
    args = dict(cfg)
    obj_type = args.pop("type")
    if MODELS.get(obj_type) is None:
        raise ValueError(f"{obj_type} is not registered")
    return MODELS.build(cfg)


def top_k_accuracy(scores, labels, topk=(1,)):
    """core/evaluation.py top_k_accuracy, on the device."""
    res = []
    maxk = max(topk)
    pred = scores.topk(maxk, dim=1).indices
    hit = pred.eq(labels.view(-1, 1))
    for k in topk:
        res.append(hit[:, :k].any(dim=1).float().mean())
    return res


@MODELS.register_module()
class CrossEntropyLoss(nn.Module):
    def __init__(self, loss_weight=1.0, class_weight=None):
        super().__init__()
        self.loss_weight = loss_weight
        self.class_weight = None if class_weight is None else torch.Tensor(class_weight)

    def forward(self, cls_score, label, **kwargs):
        if cls_score.size() == label.size():
            assert cls_score.dim() == 2 and len(kwargs) == 0
            lsm = F.log_softmax(cls_score, 1)
            if self.class_weight is not None:
                cw = self.class_weight.to(cls_score.device)
                loss = -(label * lsm * cw.unsqueeze(0)).sum(1).sum() / torch.sum(cw.unsqueeze(0) * label)
            else:
                loss = -(label * lsm).sum(1).mean()
        else:
            if self.class_weight is not None:
                assert "weight" not in kwargs
                kwargs["weight"] = self.class_weight.to(cls_score.device)
            loss = F.cross_entropy(cls_score, label, **kwargs)
        return loss * self.loss_weight


class _PoolFn(torch.autograd.Function):
    """SimpleHead(mode='GCN') pooling (heads/simple_head.py:83-90): AdaptiveAvgPool2d(1) over (T, V), then the mean over the M
    persons = one mean over the M*T*V rows of a clip.  The backbone output is channels-last rows already, so this is the temporal-
    mean kernel with V = 1 (fp32 result, no intermediate [N*M, C] tensor, no dtype-conversion pass); backward broadcasts g / count."""

    @staticmethod
    def forward(ctx, x):
        N, M_, C, T, V = x.shape
        rows = modules.to_rows(x.detach().reshape(N * M_, C, T, V), x.dtype)          # a view for a backbone output
        ctx.shape, ctx.dtype = x.shape, x.dtype
        return ops.tmean(rows, N, M_ * T * V, 1).view(N, C)

    @staticmethod
    def backward(ctx, g):
        N, M_, C, T, V = ctx.shape
        d = (g / float(M_ * T * V)).to(ctx.dtype)
        return d.view(N, 1, C, 1, 1).expand(N, M_, C, T, V)


class _HeadCEFn(torch.autograd.Function):
    """fc_cls + cross-entropy + top-1 / top-5 (heads/simple_head.py:93-96, heads/base.py:50-84, losses/cross_entropy_loss.py:77-80) as
    two kernels forward (dsg_head_ce_fwd, the batch means through dsg_tmean) and two backward (dsg_head_ce_bwd): no logits /
    log-softmax / one-hot / top-k intermediates from library kernels."""
    # fc_cls gradients go back through autograd (AccumulateGrad adds them into the flat-buffer views): with EVERY parameter of a
    # packed model written in place no AccumulateGrad node runs, and capturing the iteration in a CUDA graph then fails with
    # cudaErrorStreamCaptureIsolation in the engine's end-of-backward stream sync (tools/probes/capture_head.py, torch 2.11)
    direct_sink = False

    @staticmethod
    def forward(ctx, pooled, weight, bias, label):
        x = pooled.detach().contiguous()
        logits, stats = ops.head_ce_fwd(x, weight.detach(), bias.detach(), label)
        means = ops.tmean(stats, 1, stats.shape[0], 1).view(3)          # [mean CE, top-1, top-5]
        ctx.save_for_backward(x, logits, label)
        ctx.params = (weight, bias)
        ctx.need_x = ctx.needs_input_grad[0]
        ctx.mark_non_differentiable(logits)
        return means[0], means[1], means[2], logits

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, gloss, _g1, _g5, _gl):
        x, logits, label = ctx.saved_tensors
        weight, bias = ctx.params
        N = x.shape[0]
        gscale = (gloss.detach().float() / N).reshape(1)
        sinks = []
        for p in (weight, bias):
            direct = _HeadCEFn.direct_sink and getattr(p, "_dsg_sink", False) and p.grad is not None
            sinks.append(p.grad if direct else torch.zeros_like(p))
        _, dx = ops.head_ce_bwd(logits, label, x, weight.detach(), gscale, dW=sinks[0], db=sinks[1], need_dpooled=ctx.need_x)
        out = []
        for p, g in zip((weight, bias), sinks):
            if p.grad is not None and g.data_ptr() == p.grad.data_ptr():     # written in place into the flat bucket: tell it directly
                cb = getattr(p, "_dsg_ready", None)
                if cb is not None:
                    cb(p)
                out.append(None)
            else:
                out.append(g)
        return dx, out[0], out[1], None


@MODELS.register_module()
class GCNHead(nn.Module):
    """SimpleHead(mode='GCN'): mean over (T,V), mean over M, dropout(p) , Linear."""
    use_fused = True          # training: fc_cls + cross-entropy + top-k in the dsg_head_ce_* kernels where the configuration allows

    def __init__(self, num_classes, in_channels, loss_cls=dict(type="CrossEntropyLoss"), dropout=0., init_std=0.01,
                 multi_class=False, label_smooth_eps=0.0, **kwargs):
        super().__init__()
        self.num_classes, self.in_channels, self.in_c = num_classes, in_channels, in_channels
        self.loss_cls = build_loss(loss_cls)
        self.multi_class, self.label_smooth_eps = multi_class, label_smooth_eps
        self.dropout_ratio, self.init_std = dropout, init_std
        self.dropout = nn.Dropout(p=dropout) if dropout != 0 else None
        self.mode = "GCN"
        self.fc_cls = nn.Linear(in_channels, num_classes)

    def init_weights(self):
        nn.init.normal_(self.fc_cls.weight, 0, self.init_std)
        nn.init.constant_(self.fc_cls.bias, 0)

    def forward(self, x):
        if x.dim() != 2:
            N, M_, C, T, V = x.shape
            if x.is_cuda and x.dtype in (torch.float32, torch.bfloat16) and ops.L.is_device_build():
                x = _PoolFn.apply(x)                 # one kernel: mean over the M*T*V rows of every clip
            else:
                x = x.float().mean((3, 4)).mean(1)
        assert x.shape[1] == self.in_c
        if self.dropout is not None:
            x = self.dropout(x)
        return self.fc_cls(x)

    def fused_loss(self, x, label):
        """forward + loss of training in the fused kernels; None when the configuration needs the general path (soft / multi-class
        labels, class weights, dropout, label smoothing, CPU tensors)."""
        lc = self.loss_cls
        if (not self.use_fused or x.dim() != 5 or not x.is_cuda or x.dtype not in (torch.float32, torch.bfloat16) or not ops.L.is_device_build()
                or self.dropout is not None or self.multi_class or type(lc) is not CrossEntropyLoss or lc.class_weight is not None
                or label.dtype != torch.int64 or label.dim() != 1 or label.shape[0] != x.shape[0] or self.num_classes > 1024):
            return None
        pooled = _PoolFn.apply(x)
        loss, top1, top5, _ = _HeadCEFn.apply(pooled, self.fc_cls.weight, self.fc_cls.bias, label.contiguous())
        losses = dict(top1_acc=top1, top5_acc=top5)
        losses["loss_cls"] = loss * lc.loss_weight
        return losses

    def loss(self, cls_score, label, **kwargs):
        losses = dict()
        if label.shape == torch.Size([]):
            label = label.unsqueeze(0)
        if not self.multi_class and cls_score.size() != label.size():
            top1, top5 = top_k_accuracy(cls_score.detach(), label.detach(), (1, min(5, cls_score.shape[1])))
            losses["top1_acc"], losses["top5_acc"] = top1, top5
        elif self.multi_class and self.label_smooth_eps != 0:
            label = (1 - self.label_smooth_eps) * label + self.label_smooth_eps / self.num_classes
        losses["loss_cls"] = self.loss_cls(cls_score, label, **kwargs)
        return losses


@MODELS.register_module()
class RecognizerGCN(nn.Module):

    def __init__(self, backbone, neck=None, cls_head=None, train_cfg=dict(), test_cfg=dict()):
        super().__init__()
        if neck:
            raise NotImplementedError("necks are outside the DS-GCN path (neck=None in configs/dsstgcn)")
        self.backbone = build_backbone(backbone)
        self.neck = None
        self.cls_head = build_head(cls_head) if cls_head else None
        self.train_cfg = train_cfg or dict()
        self.test_cfg = test_cfg or dict()
        self.max_testing_views = self.test_cfg.get("max_testing_views", None)
        self.init_weights()

    @property
    def with_cls_head(self):
        return self.cls_head is not None

    @property
    def with_neck(self):
        return False

    def init_weights(self):
        self.backbone.init_weights()
        if self.with_cls_head:
            self.cls_head.init_weights()

    def extract_feat(self, keypoint):
        return self.backbone(keypoint)

    def average_clip(self, cls_score):
        assert len(cls_score.shape) == 3
        average_clips = self.test_cfg.get("average_clips", "prob")
        if average_clips not in ["score", "prob", None]:
            raise ValueError(f'{average_clips} is not supported. Supported: ["score", "prob", None]')
        if average_clips is None:
            return cls_score
        if average_clips == "prob":
            return F.softmax(cls_score, dim=2).mean(dim=1)
        return cls_score.mean(dim=1)

    def forward_train(self, keypoint, label, **kwargs):
        assert self.with_cls_head
        assert keypoint.shape[1] == 1
        if keypoint.dtype != torch.float:
            keypoint = keypoint.float()
        x = self.extract_feat(keypoint[:, 0])
        lab = label.squeeze(-1)
        fused = self.cls_head.fused_loss(x, lab) if hasattr(self.cls_head, "fused_loss") else None
        if fused is not None:
            return fused
        cls_score = self.cls_head(x)
        losses = dict()
        losses.update(self.cls_head.loss(cls_score, lab))
        return losses

    def forward_test(self, keypoint, **kwargs):
        assert self.with_cls_head
        bs, nc = keypoint.shape[:2]
        keypoint = keypoint.reshape((bs * nc,) + keypoint.shape[2:])
        x = self.extract_feat(keypoint)
        cls_score = self.cls_head(x)
        cls_score = cls_score.reshape(bs, nc, cls_score.shape[-1])
        if "average_clips" not in self.test_cfg:
            self.test_cfg["average_clips"] = "prob"
        return self.average_clip(cls_score).data.cpu().numpy()

    def forward(self, keypoint, label=None, return_loss=True, **kwargs):
        if return_loss:
            if label is None:
                raise ValueError("Label should not be None.")
            return self.forward_train(keypoint, label, **kwargs)
        return self.forward_test(keypoint, **kwargs)

    def _parse_losses(self, losses):
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f"{name} is not a tensor or list of tensors")
        loss = sum(v for k, v in log_vars.items() if "loss" in k)
        log_vars["loss"] = loss
        # recognizers/base.py:151-156 all-reduces and .item()s every scalar on its own (4 collectives + 4 host syncs per iteration);
        # the same numbers with ONE collective and ONE device-to-host copy
        names = list(log_vars)
        packed = torch.stack([log_vars[n].detach().float().reshape(()) for n in names])
        if dist.is_available() and dist.is_initialized():
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        for name, value in zip(names, packed.tolist()):
            log_vars[name] = value
        return loss, log_vars, losses

    def train_step(self, data_batch, optimizer, **kwargs):
        losses = self(**data_batch, return_loss=True)
        loss, log_vars, losses = self._parse_losses(losses)
        return dict(loss=loss, losses=losses, log_vars=log_vars, num_samples=len(next(iter(data_batch.values()))))
