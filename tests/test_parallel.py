"""N>1 path on CPU: world_size-2 gloo processes exercise the gradient all-reduce / parameter broadcast used by
bench.py --gpus N (per-rank BatchNorm statistics, mean of gradients, conv2_se skipped)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import dsgcn_b200
from dsgcn_b200 import parallel


class Tiny(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.fc = torch.nn.Linear(4, 3)
        self.conv2_se = torch.nn.Linear(4, 3)      # never used, like dgphgcn1.conv2_se
        self.bn = torch.nn.BatchNorm1d(3)

    def forward(self, x):
        return self.bn(self.fc(x)).sum()


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)                   # different init per rank: broadcast must fix it
    m = Tiny()
    parallel.broadcast_parameters(m)
    torch.manual_seed(7)
    x = torch.randn(world * 6, 4)[rank * 6:(rank + 1) * 6]      # the batch is sharded across ranks
    m(x).backward()
    params = parallel.trainable_parameters(m)
    assert all("conv2_se" not in n for n, p in m.named_parameters() if any(p is q for q in params))
    local = [p.grad.clone() for p in params]
    parallel.allreduce_gradients(params)
    torch.save(dict(w=m.fc.weight.detach(), local=local, synced=[p.grad.clone() for p in params],
                    rm=m.bn.running_mean.clone(), se_grad=m.conv2_se.weight.grad), os.path.join(out, f"r{rank}.pt"))
    dist.destroy_process_group()


def test_gradient_allreduce_two_ranks_gloo(tmp_path):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"r{i}.pt")) for i in range(2))
    assert torch.equal(r0["w"], r1["w"])                                    # broadcast made the replicas identical
    for a, b, l0, l1 in zip(r0["synced"], r1["synced"], r0["local"], r1["local"]):
        assert torch.allclose(a, b) and torch.allclose(a, (l0 + l1) / 2, atol=1e-6)   # mean of the per-rank gradients
    assert not torch.allclose(r0["rm"], r1["rm"])                           # BatchNorm statistics stay per rank (no SyncBN)
    assert r0["se_grad"] is None and r1["se_grad"] is None                  # conv2_se never gets a gradient


class TinyNet(torch.nn.Module):
    """Three 'blocks' named like the backbone so the bucketing rule (one group per gcn.N) is exercised."""

    def __init__(self):
        super().__init__()
        self.backbone = torch.nn.Module()
        self.backbone.data_bn = torch.nn.BatchNorm1d(4)
        self.backbone.gcn = torch.nn.ModuleList([torch.nn.Linear(4, 4) for _ in range(4)])
        self.backbone.gcn[1].conv2_se = torch.nn.Linear(4, 4)       # never used
        self.cls_head = torch.nn.Linear(4, 3)

    def forward(self, x):
        h = self.backbone.data_bn(x)
        for blk in self.backbone.gcn:
            h = torch.relu(blk(h))
        return self.cls_head(h).square().mean()


def _bucket_worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "emu"))
    import build_emu
    dsgcn_b200._lib._testing_use_library(build_emu.build())        # the fused SGD kernel on the host-side simulator
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(100 + rank)
    m = TinyNet()
    parallel.broadcast_parameters(m)
    ref = TinyNet()
    ref.load_state_dict(m.state_dict())
    gb = parallel.GradBuckets(m, n_buckets=3)
    opt = parallel.FlatSGD(gb, lr=0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
    ropt = torch.optim.SGD(parallel.trainable_parameters(ref), lr=0.1, momentum=0.9, weight_decay=5e-4, nesterov=True)
    torch.manual_seed(7)
    data = torch.randn(3, world * 6, 4)
    launched = []
    for it in range(3):
        lr = parallel.cosine_lr(0.1, it, 3)
        opt.set_lr(lr)
        for g in ropt.param_groups:
            g["lr"] = lr
        x = data[it, rank * 6:(rank + 1) * 6]
        opt.zero_grad()
        m(x).backward()
        launched.append([bk.work is not None for bk in gb.buckets])  # hooks launched every bucket's collective during backward
        opt.step()
        ropt.zero_grad(set_to_none=True)
        ref(x).backward()
        parallel.allreduce_gradients(parallel.trainable_parameters(ref))
        ropt.step()
    torch.save(dict(sd={k: v.clone() for k, v in m.state_dict().items()}, ref={k: v.clone() for k, v in ref.state_dict().items()},
                    launched=launched, names=[bk.names for bk in gb.buckets],
                    flat=all(p.data_ptr() >= bk.flat_p.data_ptr() and p.data_ptr() < bk.flat_p.data_ptr() + bk.flat_p.numel() * 4
                             for bk in gb.buckets for p in bk.params)), os.path.join(out, f"b{rank}.pt"))
    dist.destroy_process_group()


def test_bucketed_overlapped_allreduce_and_flat_sgd_two_ranks_gloo(tmp_path):
    """GradBuckets + FlatSGD (flat parameter/gradient buffers, per-bucket async all-reduce launched from autograd hooks, one fused
    update per bucket, cosine schedule through the device-side learning rate) == per-parameter all-reduce + torch.optim.SGD."""
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_bucket_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (torch.load(os.path.join(tmp_path, f"b{i}.pt")) for i in range(2))
    assert r0["flat"] and r1["flat"]
    assert all(all(row) for row in r0["launched"])
    names = r0["names"]
    assert len(names) == 3 and names[0][0].startswith("cls_head") and names[-1][-1].startswith("backbone.data_bn")
    assert not any("conv2_se" in n for b in names for n in b)
    for k in r0["sd"]:
        if "conv2_se" in k or "running" in k or "num_batches" in k:
            continue
        assert torch.allclose(r0["sd"][k], r1["sd"][k], atol=1e-7), k                 # replicas stay identical
        assert torch.allclose(r0["sd"][k], r0["ref"][k], atol=1e-6), k                # and equal the un-bucketed reference update
    assert not torch.allclose(r0["sd"]["backbone.data_bn.running_mean"], r1["sd"]["backbone.data_bn.running_mean"])
