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
        if self.flat.is_cuda:
            from . import ops
            ops.zero_(self.flat)          # stream-ordered memset (a memset node in a captured step)
        else:
            self.flat.zero_()
        # autograd accumulates in place into existing .grad tensors; re-attach in case a caller
        # replaced them (e.g. optimizer.zero_grad(set_to_none=True)).
        off = 0
        for p in self.params:
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat.data_ptr() + 4 * off:
                p.grad = self.flat[off:off + n].view_as(p)
            off += n

    # ---- overlapped exchange -------------------------------------------------------------------------
    # The kernels accumulate straight into the flat buffer, so autograd never sees parameter gradients and
    # parameter hooks do not fire; instead every backward op of `ops` reports the parameters it has just
    # finished (`ops.grad_ready_hook`).  A parameter fed by several ops (classifier.weight: output head and
    # PrevPredEmbeddings) is ready after its last report; the number of reports per parameter is learned in
    # the first step.  Maximal runs of ready, not yet reduced parameters are all-reduced on a side stream as
    # soon as they reach `bucket_bytes`, so the NVLink exchange runs under the rest of the backward pass.
    def enable_overlap(self, group=None, average=False, bucket_bytes=48 << 20):
        from . import ops
        self._ov = {"group": group, "average": average, "bucket": int(bucket_bytes), "expected": None,
                    "seen": {}, "stream": None, "works": []}
        self._index = {id(p): i for i, p in enumerate(self.params)}
        offs, off = [], 0
        for p in self.params:
            offs.append(off)
            off += p.numel()
        self._offs = offs + [off]
        ops.grad_ready_hook = self._notify

    def begin_step(self):
        ov = self._ov
        if ov["expected"] is None:
            ov["seen"] = {}
        else:
            ov["left"] = list(ov["expected"])
            ov["ready"] = [False] * len(self.params)
            ov["sent"] = [False] * len(self.params)
        ov["works"] = []

    def _notify(self, params):
        ov = getattr(self, "_ov", None)
        if ov is None:
            return
        if ov["expected"] is None:                     # learning step
            for p in params:
                i = self._index.get(id(p))
                if i is not None:
                    ov["seen"][i] = ov["seen"].get(i, 0) + 1
            return
        changed = False
        for p in params:
            i = self._index.get(id(p))
            if i is None or ov["ready"][i]:
                continue
            ov["left"][i] -= 1
            if ov["left"][i] <= 0:
                ov["ready"][i] = True
                changed = True
        if changed:
            self._flush(final=False)

    def _flush(self, final):
        ov = self._ov
        n = len(self.params)
        i = 0
        while i < n:
            if ov["sent"][i] or not (ov["ready"][i] or final):
                i += 1
                continue
            j = i
            while j < n and not ov["sent"][j] and (ov["ready"][j] or final):
                j += 1
            lo, hi = self._offs[i], self._offs[j]
            if final or (hi - lo) * 4 >= ov["bucket"]:
                self._launch(lo, hi)
                for k in range(i, j):
                    ov["sent"][k] = True
            i = j

    def _launch(self, lo, hi):
        ov = self._ov
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(ov["group"]) == 1:
            return
        view = self.flat[lo:hi]
        if view.is_cuda:
            if ov["stream"] is None:
                ov["stream"] = torch.cuda.Stream()
            ev = torch.cuda.Event()
            ev.record()                                  # everything enqueued so far produced this slice
            with torch.cuda.stream(ov["stream"]):
                ov["stream"].wait_event(ev)
                if ov["average"]:
                    view.div_(dist.get_world_size(ov["group"]))
                ov["works"].append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=ov["group"], async_op=True))
        else:
            if ov["average"]:
                view.div_(dist.get_world_size(ov["group"]))
            ov["works"].append(dist.all_reduce(view, op=dist.ReduceOp.SUM, group=ov["group"], async_op=True))

    def finish_step(self):
        """Call after backward(): exchanges whatever is left and makes the current stream wait for all of it."""
        ov = self._ov
        if ov["expected"] is None:                     # learning step: one plain all-reduce
            ov["expected"] = [ov["seen"].get(i, 0) for i in range(len(self.params))]
            self.all_reduce(group=ov["group"], average=ov["average"])
            return
        self._flush(final=True)
        for w in ov["works"]:
            w.wait()
        if ov["stream"] is not None:
            torch.cuda.current_stream().wait_stream(ov["stream"])

    def all_reduce(self, group=None, average=False, async_op=False):
        """Sum (or average) gradients over the data-parallel group with one collective.  Prefer the sum with
        `global_loss_scale` on the loss: averaging costs a pass over the buffer."""
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return None
        if average:
            self.flat.div_(dist.get_world_size(group))
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)


class GradExchange(object):
    """The one collective of a data-parallel step: SUM of the flat gradient buffer over the ranks.

    The 1/N of an average never touches the buffer: scale the local loss instead (`global_loss_scale`, which also
    reproduces the reference's GLOBAL `max(sum(loss_mask), 1)` normaliser, sam/task_utils.py:28-29 on gathered scores).
    wire_dtype torch.float32: one NCCL all-reduce in place (387 MB for the shipped model).
    wire_dtype torch.bfloat16 (SAMK_DP_WIRE=bf16): the buffer is packed to bf16 (one pass), all-reduced (193 MB) and
    unpacked; the sum then carries bf16 rounding (2^-9 relative per addend) -- off by default."""

    def __init__(self, grads, world, group=None, wire_dtype=None, overlapped=False):
        import os
        self.grads, self.world, self.group = grads, world, group
        self.overlapped = overlapped      # the bucketed exchange runs inside the step (FlatGradBuffer.enable_overlap)
        if wire_dtype is None:
            wire_dtype = torch.bfloat16 if os.environ.get("SAMK_DP_WIRE", "f32") == "bf16" else torch.float32
        self.wire_dtype = wire_dtype
        self.wire = torch.empty_like(grads.flat, dtype=wire_dtype) if wire_dtype != torch.float32 else None
        self.chunks = int(os.environ.get("SAMK_DP_CHUNKS", "1"))    # measured at N=2: 4 chunks 0.78 ms exposed, 1 chunk 0.52 ms
        self.kernels_per_call = 0 if self.wire is None else 2 * self.chunks

    def all_reduce(self):
        if self.overlapped or self.world <= 1 or not (dist.is_available() and dist.is_initialized()):
            return
        flat = self.grads.flat
        if self.wire is None:
            dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
            return
        from . import ops
        # pack | all-reduce | unpack, optionally as a pipeline over SAMK_DP_CHUNKS chunks (the collective of chunk c on
        # NCCL's stream while chunk c+1 is packed and chunk c-1 unpacked); one chunk measured fastest
        n = flat.numel()
        C = max(1, min(self.chunks, n // (1 << 20)))
        step = (n + C - 1) // C
        step = (step + 1023) // 1024 * 1024                 # 16-byte aligned chunk starts in both formats
        bounds = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
        works = []
        for lo, hi in bounds:
            ops.cast_flat(flat[lo:hi], self.wire[lo:hi])
            works.append(dist.all_reduce(self.wire[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        for (lo, hi), w in zip(bounds, works):
            w.wait()                                        # stream-side wait: the caller's stream waits for that chunk
            ops.cast_flat(self.wire[lo:hi], flat[lo:hi])

    def describe(self):
        import os
        n = self.grads.flat.numel()
        if self.overlapped:
            ov = self.grads._ov
            return {"collective": "bucketed ncclAllReduce(sum) on a side stream under the backward pass (captured in the step's "
                                  "CUDA graph), buckets >= %d MB of finished gradients, NCCL_MAX_CTAS=%s, %s SMs left out of "
                                  "the persistent kernels' grids" % (ov["bucket"] >> 20, os.environ.get("NCCL_MAX_CTAS", "default"),
                                                                     os.environ.get("SAMK_DP_RESERVE_SMS", "8")),
                    "wire_dtype": "float32", "bytes": n * 4, "average": "folded into the loss scale (no pass over the buffer)"}
        return {"collective": "ncclAllReduce(sum) over the flat gradient buffer, after the step" +
                              ("" if self.wire is None else ", pack | all-reduce | unpack pipelined over %d chunks" % self.chunks),
                "wire_dtype": str(self.wire_dtype).replace("torch.", ""),
                "bytes": n * (4 if self.wire is None else 2), "average": "folded into the loss scale (no pass over the buffer)"}


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


def global_loss_scale(local_mask, group=None):
    """local_count / global_count: multiply the local loss (normalised by the local count) by this and SUM the
    gradients over the ranks to reproduce the reference's global `sum(losses) / max(sum(mask), 1)`.
    local_mask: the rank's `train_loss_mask` (or its sum)."""
    t = local_mask.detach().float().sum().reshape(1)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        tot = t.clone()
        dist.all_reduce(tot, group=group)
        return (t.clamp(min=1.0) / tot.clamp(min=1.0)).item()
    return 1.0
