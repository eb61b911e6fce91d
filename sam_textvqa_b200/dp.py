"""Data-parallel runtime: one process per GPU, parameters replicated, ONE gradient all-reduce per step.

Replaces the reference's single-process `nn.DataParallel` (/root/reference/train.py:111-112), which
re-broadcasts all 96.6 M parameters and reduces gradients to GPU 0 every step.  Samples are
independent (LayerNorm only), so the batch dimension is the only partition and the gradient sum
is the only exchange.

All `.grad` tensors are views into one flat fp32 buffer, so the exchange is a single NCCL
all-reduce over NVLink/NVSwitch with no flatten/unflatten copies, and `zero_grad` is one memset.
The loss normaliser `max(sum(loss_mask),1)` is per-replica here; `global_loss_scale` gives the
factor that makes the summed gradients equal to the reference's global normalisation
(sam/task_utils.py:28-29 on the gathered scores).
"""
import torch
import torch.distributed as dist


class FlatGradBuffer(object):
    """Owns one flat fp32 gradient buffer; every parameter's .grad is a view into it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in self.params:
            n = p.numel()
            p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def zero(self):
        self.flat.zero_()
        # autograd accumulates in place into existing .grad tensors; re-attach in case a caller
        # replaced them (e.g. optimizer.zero_grad(set_to_none=True)).
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + n].view_as(p)
            off += n

    def all_reduce(self, group=None, average=True, async_op=False):
        """Sum (or average) gradients over the data-parallel group with one collective."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if average:
            self.flat.div_(dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


def shard_batch(batch, rank, world):
    """Contiguous shard of every per-sample tensor (dict values may be nested dicts)."""
    def cut(v):
        if torch.is_tensor(v) and v.dim() > 0:
            n = v.shape[0]
            per = (n + world - 1) // world
            return v[rank * per:min(n, (rank + 1) * per)]
        if isinstance(v, dict):
            return {k: cut(x) for k, x in v.items()}
        return v
    return {k: cut(v) for k, v in batch.items()}


def global_loss_scale(local_mask_sum, group=None):
    """local_count / global_count: multiply the local loss by this (with average=False reduction of
    gradients) to reproduce the reference's global `sum(losses) / max(sum(mask),1)`."""
    t = local_mask_sum.detach().clone().float().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        tot = t.clone()
        dist.all_reduce(tot, group=group)
        return (t.clamp(min=1.0) / tot.clamp(min=1.0)).item()
    return 1.0
