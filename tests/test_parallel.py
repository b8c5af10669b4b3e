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
