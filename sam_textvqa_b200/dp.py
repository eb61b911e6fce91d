"""Data-parallel runtime: one process per GPU, parameters replicated, the gradient SUM is the only exchange.

Replaces the reference's single-process `nn.DataParallel` (/root/reference/train.py:111-112), which
re-broadcasts all 96.6 M parameters and reduces gradients to GPU 0 every step.  Samples are
independent (LayerNorm only), so the batch dimension is the only partition.

All `.grad` tensors are views into one flat fp32 buffer (`FlatGradBuffer`): no flatten / unflatten copies, `zero_grad`
is one memset, and the exchange works on ranges of one index space.  Two transports:
  * `PeerWire` / `FlatGradBuffer.enable_overlap(transport="peer")` -- `samk_exchange_sum` (csrc/exchange.cu): our kernels
    over NVLink peer memory / NVSwitch multicast, bucket by bucket under the backward pass (the default of bench.py);
  * NCCL: one all-reduce over the buffer after the step (`GradExchange`), or bucketed on a side stream.
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
        self._storage = torch.zeros((total + 63) // 64 * 64, dtype=torch.float32, device=dev)   # (padding: whole wire vectors)
        self.flat = self._storage[:total]
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
    def enable_overlap(self, group=None, average=False, bucket_bytes=48 << 20, transport="nccl", wire_dtype=torch.bfloat16):
        """transport "nccl": bucketed ncclAllReduce on a side stream; "peer": samk_exchange_sum over NVLink peer memory
        (PeerWire; collective call -- every rank enables it at the same point)."""
        from . import ops
        self._ov = {"group": group, "average": average, "bucket": int(bucket_bytes), "expected": None,
                    "seen": {}, "stream": None, "works": [], "peer": None}
        if transport == "peer" and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            assert not average, "the peer exchange sums; fold 1/N into the loss scale (global_loss_scale)"
            self._ov["peer"] = PeerWire(self.flat.numel(), wire_dtype, group)
        elif transport not in ("nccl", "peer"):
            self._ov["peer"] = transport                   # an object with .vec and .exchange_sum(storage, lo, hi) (tests)
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
            nxt = j
            if ov["peer"] is not None and not final:
                # bucket edges on wire-vector boundaries (parameters whose sizes are not multiples of 8 wait for
                # their neighbours; whatever is left goes out with the final flush)
                vec = ov["peer"].vec
                while i < j and self._offs[i] % vec:
                    i += 1
                while j > i and self._offs[j] % vec:
                    j -= 1
                if j <= i:
                    i = nxt
                    continue
            lo, hi = self._offs[i], self._offs[j]
            if final or (hi - lo) * 4 >= ov["bucket"]:
                self._launch(lo, hi)
                for k in range(i, j):
                    ov["sent"][k] = True
            i = nxt

    def _launch(self, lo, hi):
        ov = self._ov
        if ov["peer"] is None and (not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(ov["group"]) == 1):
            return
        view = self.flat[lo:hi]
        peer = ov["peer"]
        if peer is not None:
            vec = peer.vec
            plo, phi = lo, hi
            if phi == self.flat.numel():
                phi = (phi + vec - 1) // vec * vec       # (the storage behind `flat` is padded, see __init__)
            if plo % vec == 0 and phi % vec == 0:
                if view.is_cuda:
                    if ov["stream"] is None:
                        ov["stream"] = torch.cuda.Stream()
                    ev = torch.cuda.Event()
                    ev.record()                          # everything enqueued so far produced this slice
                    with torch.cuda.stream(ov["stream"]):
                        ov["stream"].wait_event(ev)
                        peer.exchange_sum(self._storage, plo, phi)
                else:
                    peer.exchange_sum(self._storage, plo, phi)
                return
            # an unaligned leftover (never with the shipped model): the library collective below
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


class PeerWire(object):
    """Symmetric wire buffer + flag pads for `samk_exchange_sum` (csrc/exchange.cu): the SUM of a range of the flat
    gradient buffer over the ranks through NVLink peer memory -- NVSwitch multicast (`multimem.ld_reduce` /
    `multimem.st`) when the allocation has a multicast mapping, peer loads / stores otherwise.  The kernels are small
    enough to share SMs with the persistent GEMMs, so the exchange of a finished bucket runs UNDER the rest of the
    backward pass (FlatGradBuffer.enable_overlap(transport="peer")).  torch's symmetric-memory allocator provides the
    mappings (plumbing); the data path is ours.  Collective: every rank constructs it at the same point."""

    def __init__(self, n_elems, wire_dtype=torch.bfloat16, group=None, use_multicast=None, timeout_clocks=0):
        import ctypes
        import os
        import torch.distributed._symmetric_memory as symm
        from . import _lib
        group = group if group is not None else dist.group.WORLD
        dev = torch.device("cuda", torch.cuda.current_device())
        if os.environ.get("SAMK_XCHG_FORCE_FAIL", "0") == "1":          # (exercises the callers' fallback to the library path)
            raise RuntimeError("peer exchange disabled by SAMK_XCHG_FORCE_FAIL")
        self.group, self.wire_dtype = group, wire_dtype
        self.vec = 8 if wire_dtype == torch.bfloat16 else 4
        self.n = (int(n_elems) + 63) // 64 * 64
        self.wire = symm.empty(self.n, dtype=wire_dtype, device=dev)
        self.hdl = symm.rendezvous(self.wire, group)
        self.flags = symm.empty(64, dtype=torch.int32, device=dev)
        self.flags.zero_()
        self.fhdl = symm.rendezvous(self.flags, group)
        self.state = torch.zeros(2, dtype=torch.int32, device=dev)          # [epoch, error]
        self.rank, self.world = int(self.hdl.rank), int(self.hdl.world_size)
        mc = int(getattr(self.hdl, "multicast_ptr", 0) or 0)
        if use_multicast is None:
            # measured (profiles/r02_dp_experiments.txt): two ranks are faster with plain peer loads / stores (each rank
            # reads half of one peer's buffer), four and eight with the switch doing the additions
            env = os.environ.get("SAMK_XCHG_MULTICAST")
            use_multicast = (env != "0") if env is not None else self.world > 2
        self.multicast = bool(mc) and use_multicast
        self._wire_ptrs = (ctypes.c_void_p * self.world)(*[int(x) for x in self.hdl.buffer_ptrs])
        self._flag_ptrs = (ctypes.c_void_p * self.world)(*[int(x) for x in self.fhdl.buffer_ptrs])
        pw = _lib.PeerWire()
        pw.rank, pw.world = self.rank, self.world
        pw.wire_dtype = _lib.DT_BF16 if wire_dtype == torch.bfloat16 else _lib.DT_F32
        pw.wire_peers = ctypes.cast(self._wire_ptrs, ctypes.POINTER(ctypes.c_void_p))
        pw.wire_mc = mc if self.multicast else None
        pw.flag_peers = ctypes.cast(self._flag_ptrs, ctypes.POINTER(ctypes.c_void_p))
        pw.epoch = self.state.data_ptr()
        pw.error = self.state.data_ptr() + 4
        pw.timeout_clocks = int(timeout_clocks)
        self._pw = pw
        self.launches_per_call = 5
        self.log = []                                                        # (lo, hi) of every call (tests, describe)
        torch.cuda.synchronize()
        dist.barrier(group)                                                  # every pad is zeroed before anyone signals

    def exchange_sum(self, flat, lo, hi):
        """flat[lo:hi] (fp32, local) <- sum over the ranks, on the current stream.  lo, hi: multiples of self.vec."""
        import ctypes
        from . import _lib
        from ._lib import check, lib, stream_ptr
        assert flat.dtype == torch.float32 and flat.is_contiguous() and hi <= self.n
        if len(self.log) < 4096:
            self.log.append((int(lo), int(hi)))
        check(lib().samk_exchange_sum(ctypes.byref(self._pw), flat.data_ptr(), int(lo), int(hi), stream_ptr()), "exchange_sum")
        from . import ops
        ops._count(self.launches_per_call)

    def check(self):
        """raises when a rank did not arrive at a barrier in time (call after a synchronize, outside the timed region)"""
        err = int(self.state[1].item())
        if err:
            raise RuntimeError("samk_exchange_sum: rank %d did not arrive in time (seen from rank %d)" % (err - 1, self.rank))

    def describe(self):
        return ("NVSwitch multicast (multimem.ld_reduce / multimem.st)" if self.multicast else "NVLink peer loads / stores") + \
            ", %s on the wire" % str(self.wire_dtype).replace("torch.", "")


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
        if self.overlapped and self.grads._ov.get("peer") is not None:
            ov = self.grads._ov
            peer = ov["peer"]
            return {"collective": "samk_exchange_sum per bucket (>= %d MB of finished gradients) on a side branch of the captured "
                                  "step, under the backward pass: pack | barrier | reduce own shard | barrier | unpack; %s; "
                                  "no SMs reserved" % (ov["bucket"] >> 20, peer.describe()),
                    "wire_dtype": str(peer.wire_dtype).replace("torch.", ""), "bytes": n * (2 if peer.vec == 8 else 4),
                    "average": "folded into the loss scale (no pass over the buffer)"}
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
