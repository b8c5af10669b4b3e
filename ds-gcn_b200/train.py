"""The step after the hot path: a minimal, mmcv-free training loop with the reference runner's hook points
and a checkpoint writer producing the reference's `epoch_N.pth` layout.

  runner loop      pyskl/core/local_runner/epoch_based_sparse_runner.py:44-63 (train: before_train_epoch / before_train_iter /
                   run_iter -> model.train_step / after_train_iter / after_train_epoch), mmcv OptimizerHook = zero_grad, backward, step
  schedule         configs/_init_/lr_schedual.py:11-13  SGD(lr=0.1, momentum=0.9, wd=5e-4, nesterov) + CosineAnnealing(min_lr=0, by_epoch=False)
  checkpoint       mmcv.runner.save_checkpoint: dict(meta=..., state_dict=OrderedDict of CPU tensors, optimizer=...)
"""
import os
import time
from collections import OrderedDict

import torch

from . import parallel


def save_checkpoint(model, filename, optimizer=None, meta=None):
    """`{'meta', 'state_dict', 'optimizer'}` with CPU tensors and no 'module.' prefix — what tools/test.py loads."""
    model = getattr(model, "module", model)
    meta = dict(meta or {})
    meta.setdefault("time", time.asctime())
    sd = OrderedDict((k, v.detach().cpu().clone()) for k, v in model.state_dict().items())
    ckpt = dict(meta=meta, state_dict=sd)
    if optimizer is not None:
        ckpt["optimizer"] = optimizer.state_dict()
    os.makedirs(os.path.dirname(os.path.abspath(filename)) or ".", exist_ok=True)
    torch.save(ckpt, filename)
    return filename


def load_checkpoint(model, filename, map_location="cpu", strict=False):
    """mmcv.runner.load_checkpoint semantics for this path: accepts a bare state dict or {'state_dict': ...}, strips 'module.'."""
    ckpt = torch.load(filename, map_location=map_location)
    sd = ckpt.get("state_dict", ckpt)
    sd = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in sd.items())
    getattr(model, "module", model).load_state_dict(sd, strict=strict)
    return ckpt


class GraphedTrainStep:
    """One whole training iteration — forward, loss, backward (with the bucketed all-reduce), optimizer update — captured in a CUDA
    graph and replayed on static buffers: what the OptimizerHook sequence `zero_grad(); train_step(); loss.backward(); step()` costs
    in ~430 eager launches becomes one graph launch.  Call it with the batch (host-pinned or device tensors of the captured shapes);
    it copies the batch into the static inputs, replays, and returns the same dict `RecognizerGCN.train_step` returns (`log_vars`
    read back with ONE 16-byte device-to-host copy).  The learning rate is read from device memory by `FlatSGD`, so schedules
    work across replays (`optimizer.set_lr`).  Shapes are fixed at construction; BatchNorm buffers and parameters update in place.
    """

    def __init__(self, model, optimizer, keypoint, label, warmup=3):
        if not keypoint.is_cuda:
            raise ValueError("construct GraphedTrainStep with device tensors of the batch shape")
        self.model, self.optimizer = model, optimizer
        self.kp, self.lb = keypoint.clone(), label.clone()
        self.names = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                 # eager warm-up on a side stream (autograd nodes must not be tied to the default stream)
            for _ in range(max(warmup, 1)):
                self._iteration()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._iteration()
        torch.cuda.synchronize()

    def _iteration(self):
        import torch.distributed as dist
        self.optimizer.zero_grad()
        losses = self.model(self.kp, self.lb, return_loss=True)
        log = {k: v.mean() for k, v in losses.items()}
        loss = sum(v for k, v in log.items() if "loss" in k)
        log["loss"] = loss
        self.names = list(log)
        packed = torch.stack([log[k].detach().float().reshape(()) for k in self.names])
        if dist.is_available() and dist.is_initialized():
            packed = packed / dist.get_world_size()
            dist.all_reduce(packed)
        self.packed, self.loss = packed, loss
        loss.backward()
        self.optimizer.step()

    def __call__(self, keypoint, label):
        self.kp.copy_(keypoint, non_blocking=True)
        self.lb.copy_(label, non_blocking=True)
        self.graph.replay()
        vals = self.packed.tolist()                   # the one host synchronisation of the iteration
        return dict(loss=self.loss, log_vars=OrderedDict(zip(self.names, vals)), num_samples=len(keypoint))

    def release(self):
        """Drop the captured graph (before destroying a process group whose collectives it holds)."""
        self.graph.reset()


class Runner:
    """Epoch-based loop with the hook points of EpochBasedSparseRunner.  `hooks`: objects with any of before_run, before_train_epoch,
    before_train_iter, after_train_iter, after_train_epoch, after_run (called with the runner)."""

    def __init__(self, model, optimizer, work_dir=None, max_epochs=1, base_lr=None, min_lr=0.0, hooks=()):
        self.model, self.optimizer, self.work_dir, self.max_epochs = model, optimizer, work_dir, max_epochs
        self.base_lr = base_lr if base_lr is not None else optimizer.param_groups[0]["lr"]
        self.min_lr, self.hooks = min_lr, list(hooks)
        self.epoch = self.iter = self.inner_iter = 0
        self.outputs, self.log_buffer = None, []

    def call_hook(self, name):
        for h in self.hooks:
            fn = getattr(h, name, None)
            if fn is not None:
                fn(self)

    def _set_lr(self, lr):
        if hasattr(self.optimizer, "set_lr"):
            self.optimizer.set_lr(lr)
        else:
            for g in self.optimizer.param_groups:
                g["lr"] = lr

    def train(self, data_loader):
        self.model.train()
        self.max_iters = self.max_epochs * len(data_loader)
        self.call_hook("before_train_epoch")
        for i, data_batch in enumerate(data_loader):
            self.inner_iter = i
            self._set_lr(parallel.cosine_lr(self.base_lr, self.iter, self.max_iters, self.min_lr))      # CosineAnnealing, by_epoch=False
            self.call_hook("before_train_iter")
            self.optimizer.zero_grad()
            self.outputs = self.model.train_step(data_batch, self.optimizer)
            self.outputs["loss"].backward()                                                            # OptimizerHook
            self.optimizer.step()
            self.log_buffer.append(self.outputs["log_vars"])
            self.call_hook("after_train_iter")
            self.iter += 1
        self.call_hook("after_train_epoch")
        self.epoch += 1

    def run(self, data_loader, checkpoint_interval=1):
        self.call_hook("before_run")
        while self.epoch < self.max_epochs:
            self.train(data_loader)
            if self.work_dir and checkpoint_interval and self.epoch % checkpoint_interval == 0:
                save_checkpoint(self.model, os.path.join(self.work_dir, f"epoch_{self.epoch}.pth"), self.optimizer,
                                meta=dict(epoch=self.epoch, iter=self.iter))
        self.call_hook("after_run")
