"""Data-parallel plumbing for the DS-GCN path and the step that follows it (SGD, schedule).

The reference trains under DDP with plain (per-rank) BatchNorm statistics and `broadcast_buffers=False`
(pyskl/apis/train.py:93-104): every op on the hot path is independent across clips, so the only collective is the
gradient all-reduce (mean).  DDP's reducer buckets the gradients and launches each bucket's all-reduce as soon as its
last gradient has been produced, so the collective overlaps the rest of backward; `GradBuckets` does the same with

  * pre-registered FLAT buffers: the parameters of a bucket and their gradients are views of one fp32 buffer each
    (`p.data`, `p.grad`), so there is no flatten / unflatten copy around the collective and the optimizer update is ONE
    kernel launch per bucket (`dsg_sgd_step_dev`, learning rate read from device memory);
  * buckets in gradient-ready order (head, blocks 9..0, data_bn): blocks 7-9 hold ~74 % of the parameters but only the
    first quarter of the backward time, so with three buckets only the small last one (shallow blocks, ~0.1 MB) is exposed;
  * one NCCL all-reduce (`ReduceOp.AVG`) per bucket, launched from a post-accumulate-grad hook with `async_op=True` (NCCL's own
    stream; the current stream only waits in `synchronize()`, right before the update).  All of it is CUDA-graph capturable:
    bench.py captures forward + backward + collectives + update in one graph per rank at every N.

Parameters that never receive a gradient (`conv2_se.*`, gcn.py:2253-2254) stay outside the buckets, which keeps SGD from touching
them — the same as the reference, whose optimizer skips `grad is None`.
"""
import math

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors

from . import ops


def trainable_parameters(model):
    """Parameters the reference's optimizer actually updates: everything except the never-used conv2_se convolutions."""
    return [p for n, p in model.named_parameters() if "conv2_se" not in n and p.requires_grad]


def _dist_on():
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def broadcast_parameters(model, src=0):
    """Replicate rank `src`'s parameters and buffers (DDP does this once at construction)."""
    if not _dist_on():
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def allreduce_gradients(params, world_size=None):
    """Mean of the per-rank gradients, in place, with a single collective over one flat buffer (the simple, un-overlapped
    form; `GradBuckets` is the one the benchmark uses)."""
    if not _dist_on():
        return
    world_size = world_size or dist.get_world_size()
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = _flatten_dense_tensors(grads)
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # the mean is taken inside the collective
    else:
        dist.all_reduce(flat)
        flat.div_(world_size)
    torch._foreach_copy_(grads, list(_unflatten_dense_tensors(flat, grads)))


def _group_key(name):
    """Bucket granularity: 'backbone.gcn.7.tcn.bn.weight' -> 'backbone.gcn.7' (one block); 'backbone.data_bn.weight' ->
    'backbone.data_bn'; 'cls_head.fc_cls.weight' -> 'cls_head'."""
    parts = name.split(".")
    for i in range(len(parts) - 1):
        if parts[i] == "gcn" and parts[i + 1].isdigit():
            return ".".join(parts[:i + 2])
    if parts[0] == "backbone" and len(parts) > 2:
        return ".".join(parts[:2])
    return parts[0]


class _Bucket:
    __slots__ = ("params", "names", "offsets", "flat_p", "flat_g", "flat_m", "pending", "work")


class GradBuckets:
    """Flat parameter / gradient buffers in gradient-ready order + overlapped per-bucket all-reduce.

        gb = GradBuckets(model, n_buckets=3)        # after model.to(device) and broadcast_parameters
        gb.zero_grad(); loss.backward(); gb.synchronize(); <update reading gb.buckets[i].flat_p / flat_g>

    `direct=True` (default): the kernels accumulate parameter gradients straight into the flat gradient buffer (functional.py
    `grad_like` / `grad_cat`; exactly one backward per zero_grad), and parameters a unit always uses concatenated
    (`module._flat_groups()`) are laid out adjacently so the concatenation is a view (`module._dsg_flat`).
    """
    ALIGN = 64          # floats: every parameter (group) starts on a 256-byte boundary (vectorised reductions, TMA-able slices)

    def __init__(self, model, n_buckets=3, overlap=True, direct=True):
        named = [(n, p) for n, p in model.named_parameters() if "conv2_se" not in n and p.requires_grad]
        if not named:
            raise ValueError("no trainable parameters")
        dev = named[0][1].device
        for n, p in named:
            if p.dtype != torch.float32 or p.device != dev:
                raise ValueError(f"{n}: parameters must be fp32 on one device")
        trainable = {id(p) for _, p in named}
        group_of = {}                                 # id(param) -> (module, key, [params])
        if direct:
            for mod in model.modules():
                fg = getattr(mod, "_flat_groups", None)
                if fg is None:
                    continue
                for key, plist in fg().items():
                    if len(plist) > 1 and all(id(q) in trainable for q in plist):
                        for q in plist:
                            group_of[id(q)] = (mod, key, plist)
        ready = list(reversed(named))                 # autograd produces gradients from the head back to data_bn
        name_of = {id(p): n for n, p in named}
        groups, cur = [], None
        for n, p in ready:
            k = _group_key(n)
            if k != cur:
                groups.append([])
                cur = k
            groups[-1].append((n, p))
        n_buckets = max(1, min(n_buckets, len(groups)))
        per = math.ceil(len(groups) / n_buckets)
        pad = lambda nel: (nel + self.ALIGN - 1) // self.ALIGN * self.ALIGN
        self.buckets = []
        for b in range(n_buckets):
            items = [it for g in groups[b * per:(b + 1) * per] for it in g]
            if not items:
                continue
            in_bucket = {id(p) for _, p in items}
            bk = _Bucket()
            bk.names, bk.params, bk.offsets = [], [], []
            placed, tot, cat_views = set(), 0, []
            for n, p in items:
                if id(p) in placed:
                    continue
                members = [p]
                if id(p) in group_of:
                    mod, key, members = group_of[id(p)]
                    if not all(id(q) in in_bucket for q in members):
                        raise ValueError(f"concatenated parameter group {key} of {type(mod).__name__} spans two buckets")
                    cat_views.append((mod, key, members, tot))
                for q in members:
                    bk.names.append(name_of[id(q)])
                    bk.params.append(q)
                    bk.offsets.append(tot)
                    placed.add(id(q))
                    tot += q.numel()
                tot = pad(tot)
            bk.flat_p = torch.zeros(tot, dtype=torch.float32, device=dev)
            bk.flat_g = torch.zeros(tot, dtype=torch.float32, device=dev)
            bk.flat_m = None
            with torch.no_grad():
                for p, o in zip(bk.params, bk.offsets):
                    view = bk.flat_p[o:o + p.numel()].view(p.shape)
                    view.copy_(p.data)
                    p.data = view                                   # the module's parameter now lives in the flat buffer
                    p.grad = bk.flat_g[o:o + p.numel()].view(p.shape)
                    if direct:
                        p._dsg_sink = True
            for mod, key, members, o in cat_views:
                nel = sum(q.numel() for q in members)
                rows = sum(q.shape[0] for q in members)
                shape = (rows,) if members[0].dim() == 1 else (rows, nel // rows)
                if not hasattr(mod, "_dsg_flat"):
                    mod._dsg_flat = {}
                mod._dsg_flat[key] = (bk.flat_p[o:o + nel].view(shape), bk.flat_g[o:o + nel].view(shape))
            bk.pending, bk.work = len(bk.params), None
            self.buckets.append(bk)
        self.params = [p for bk in self.buckets for p in bk.params]
        self.world = dist.get_world_size() if _dist_on() else 1
        self.overlap = overlap
        self._avg_in_collective = _dist_on() and dist.get_backend() == "nccl"
        self._hooks = []
        if self.world > 1 and overlap:
            for bk in self.buckets:
                hook = self._make_hook(bk)
                for p in bk.params:
                    self._hooks.append(p.register_post_accumulate_grad_hook(hook))    # gradients that arrive through autograd
                    p._dsg_ready = hook                                               # gradients the kernels wrote in place

    def _make_hook(self, bk):
        def hook(_p):
            bk.pending -= 1
            if bk.pending == 0:
                self._launch(bk)
        return hook

    def _launch(self, bk):
        if self._avg_in_collective:
            bk.work = dist.all_reduce(bk.flat_g, op=dist.ReduceOp.AVG, async_op=True)
        else:
            bk.work = dist.all_reduce(bk.flat_g, async_op=True)

    def zero_grad(self):
        """One memset per bucket; gradients stay views of the flat buffers (never set to None)."""
        for bk in self.buckets:
            bk.flat_g.zero_()
            bk.pending, bk.work = len(bk.params), None
            for p in bk.params:                                      # a foreign zero_grad(set_to_none=True) would detach the views
                if p.grad is None or p.grad.data_ptr() < bk.flat_g.data_ptr() or p.grad.data_ptr() >= bk.flat_g.data_ptr() + bk.flat_g.numel() * 4:
                    self._reattach(bk)
                    break

    def _reattach(self, bk):
        for p, o in zip(bk.params, bk.offsets):
            p.grad = bk.flat_g[o:o + p.numel()].view(p.shape)

    def synchronize(self):
        """After backward: make the current stream wait for every bucket's collective (launching the ones whose hooks did
        not fire, e.g. overlap=False or a parameter that received no gradient this step)."""
        if self.world == 1:
            return
        for bk in self.buckets:
            if bk.work is None:
                self._launch(bk)
        for bk in self.buckets:
            bk.work.wait()
            if not self._avg_in_collective:
                bk.flat_g.div_(self.world)
            bk.work = None
            bk.pending = len(bk.params)

    def grad_bytes(self):
        return sum(bk.flat_g.numel() * 4 for bk in self.buckets)


class FlatSGD:
    """SGD with momentum / weight decay / Nesterov (torch.optim.SGD semantics, dampening 0) over GradBuckets: one
    `dsg_sgd_step_dev` launch per bucket.  Reference: optimizer = dict(type='SGD', lr=0.1, momentum=0.9, weight_decay=5e-4,
    nesterov=True) (configs/_init_/lr_schedual.py:11), driven by mmcv's OptimizerHook as zero_grad() / backward / step().
    `param_groups[0]['lr']` is what a scheduler sets; the kernel reads it from device memory (`set_lr`), so a captured step
    follows the schedule without re-capture."""

    def __init__(self, buckets, lr=0.1, momentum=0.9, weight_decay=0.0, nesterov=False, dampening=0):
        if dampening != 0:
            raise NotImplementedError("dampening != 0")
        if nesterov and momentum <= 0:
            raise ValueError("Nesterov momentum requires a momentum")
        self.gb = buckets
        self.param_groups = [dict(params=buckets.params, lr=lr, initial_lr=lr, momentum=momentum, weight_decay=weight_decay,
                                  nesterov=nesterov, dampening=0)]
        dev = buckets.buckets[0].flat_p.device
        self._lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=dev)
        self._lr_host = float(lr)
        for bk in buckets.buckets:
            bk.flat_m = torch.zeros_like(bk.flat_p)

    def set_lr(self, lr):
        """Host -> device refresh of the learning rate (call outside a captured region)."""
        self.param_groups[0]["lr"] = float(lr)
        if float(lr) != self._lr_host:
            self._lr_dev.fill_(float(lr))
            self._lr_host = float(lr)

    def zero_grad(self, set_to_none=False):
        self.gb.zero_grad()

    def step(self):
        g = self.param_groups[0]
        if g["lr"] != self._lr_host and not (self._lr_dev.is_cuda and torch.cuda.is_current_stream_capturing()):
            self.set_lr(g["lr"])
        self.gb.synchronize()
        for bk in self.gb.buckets:
            ops.sgd_step(bk.flat_p, bk.flat_g, bk.flat_m, self._lr_dev, g["momentum"], g["weight_decay"], g["nesterov"])

    def state_dict(self):
        return dict(param_groups=[{k: v for k, v in self.param_groups[0].items() if k != "params"}],
                    momentum=[bk.flat_m.clone() for bk in self.gb.buckets])

    def load_state_dict(self, sd):
        self.param_groups[0].update(sd["param_groups"][0])
        for bk, m in zip(self.gb.buckets, sd["momentum"]):
            bk.flat_m.copy_(m)
        self.set_lr(self.param_groups[0]["lr"])


def cosine_lr(base_lr, it, total_iters, min_lr=0.0):
    """mmcv CosineAnnealingLrUpdaterHook with by_epoch=False (configs/_init_/lr_schedual.py:12): annealing_cos(base, min, it/total)."""
    f = min(max(it / max(total_iters, 1), 0.0), 1.0)
    return min_lr + 0.5 * (base_lr - min_lr) * (1.0 + math.cos(math.pi * f))
