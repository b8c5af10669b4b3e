"""Data-parallel plumbing for the DS-GCN path.

The reference trains under DDP with plain (per-rank) BatchNorm statistics and `broadcast_buffers=False`
(pyskl/apis/train.py:93-104): every op on the hot path is independent across clips, so the only collective is the
gradient all-reduce (mean) once per step.  Parameters that never receive a gradient (`conv2_se.*`,
gcn.py:2253-2254) are skipped, which keeps SGD from touching them — the same as the reference, whose optimizer skips
`grad is None`.  One flat buffer, one NCCL launch (5.5 MB: latency-bound, not bandwidth-bound)."""
import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors


def trainable_parameters(model):
    """Parameters the reference's optimizer actually updates: everything except the never-used conv2_se convolutions."""
    return [p for n, p in model.named_parameters() if "conv2_se" not in n and p.requires_grad]


def broadcast_parameters(model, src=0):
    """Replicate rank `src`'s parameters and buffers (DDP does this once at construction)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src)


def allreduce_gradients(params, world_size=None):
    """Mean of the per-rank gradients, in place, with a single collective over one flat buffer."""
    if not (dist.is_available() and dist.is_initialized()):
        return
    world_size = world_size or dist.get_world_size()
    if world_size == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    flat = _flatten_dense_tensors(grads)
    if dist.get_backend() == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)          # the mean is taken inside the collective
    else:
        dist.all_reduce(flat)
        flat.div_(world_size)
    # scatter back with multi-tensor copies (a handful of launches instead of one per parameter)
    torch._foreach_copy_(grads, list(_unflatten_dense_tensors(flat, grads)))
