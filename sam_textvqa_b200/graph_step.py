"""Whole training step (forward + loss + backward) captured once into a CUDA graph and replayed.

One step of the shipped model is ~310 kernel launches through ctypes plus the autograd tape; on the host that is
about as long as the ~11.5 ms the GPU needs, so the eager loop is launch-bound as soon as anything else (input
upload, loss read-back) also needs the CPU.  Replaying the captured graph costs the host one launch.

What makes this safe for training:
  * inputs live in static device buffers (`load` copies a batch into them; host tensors may be pinned);
  * gradients go to the static `dp.FlatGradBuffer` (zeroed inside the graph);
  * the bf16 operand copies of the weights are re-made inside the graph (the weight cache is cleared before capture),
    so an optimizer step between replays is honoured;
  * dropout masks change per replay: the captured step starts with `samk_advance_dropout_salt` (a device-side counter
    is hashed into a salt that is XORed into every Philox key; the captured kernels keep their seed / offset
    arguments).  No host work per replay besides the graph launch.
The gradient all-reduce and the optimizer stay outside the graph.
"""
import torch

from . import ops
from ._lib import check, lib, stream_ptr


class GraphedTrainStep(object):
    def __init__(self, model, grads, example_batch, loss_fn=None, warmup=3, allreduce=None):
        """example_batch: dict of device tensors (shapes and dtypes of every later batch); may contain the
        'spatial_adj_matrices' dict.  loss_fn(scores, batch) -> scalar; default = masked BCE on targets."""
        # allreduce: None = exchange outside (caller), "overlap" = capture the bucketed NCCL all-reduce of
        # dp.FlatGradBuffer.enable_overlap() inside the graph as a parallel branch (first replay-able after the
        # learning step that the eager warm-up provides)
        self.model, self.grads, self.allreduce = model, grads, allreduce
        self.loss_fn = loss_fn or (lambda scores, b: ops.bce_with_mask_loss(scores, b["targets"], b["train_loss_mask"]))
        self.device = next(model.parameters()).device
        self.static = self._clone(example_batch)
        self.replays = 0
        self.graph = None
        self._zero_stream = None
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._eager_step()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        ops.clear_weight_cache()            # weight operand casts must be part of the captured work
        self.graph = torch.cuda.CUDAGraph()
        l0 = ops.launch_count
        with torch.cuda.graph(self.graph):
            self.loss = self._eager_step()
        self.kernels_per_replay = ops.launch_count - l0       # samk kernels captured (torch's own fills / cats not counted)
        torch.cuda.synchronize()

    def _clone(self, b):
        """Static device copies of every tensor of the example batch (host tensors are uploaded)."""
        dev = self.device
        out = {}
        for k, v in b.items():
            if torch.is_tensor(v):
                out[k] = v.to(dev, copy=True)
            elif isinstance(v, dict):
                out[k] = {kk: (vv.to(dev, copy=True) if torch.is_tensor(vv) else vv) for kk, vv in v.items()}
            else:
                out[k] = v
        return out

    def _eager_step(self):
        # first node of the captured step: new dropout salt for this replay, computed and installed on the device
        check(lib().samk_advance_dropout_salt(stream_ptr()), "advance salt")
        # the 387 MB gradient buffer is cleared on its own stream beside the forward pass (a parallel branch of the
        # captured graph) and joined before the backward pass starts accumulating into it
        main = torch.cuda.current_stream()
        if self._zero_stream is None:
            self._zero_stream = torch.cuda.Stream()
        self._zero_stream.wait_stream(main)
        with torch.cuda.stream(self._zero_stream):
            self.grads.zero()
        if self.allreduce == "overlap":
            self.grads.begin_step()
        bd = dict(self.static)
        if isinstance(bd.get("spatial_adj_matrices"), dict):
            bd["spatial_adj_matrices"] = dict(bd["spatial_adj_matrices"])
        scores = self.model(bd)["textvqa_scores"]
        loss = self.loss_fn(scores, bd)
        main.wait_stream(self._zero_stream)
        loss.backward()
        if self.allreduce == "overlap":
            self.grads.finish_step()
        return loss

    def load(self, batch, non_blocking=True):
        """Copy a batch (host pinned or device tensors) into the static input buffers on the current stream."""
        for k, v in batch.items():
            dst = self.static.get(k)
            if torch.is_tensor(v) and torch.is_tensor(dst):
                dst.copy_(v, non_blocking=non_blocking)
            elif isinstance(v, dict) and isinstance(dst, dict):
                for kk, vv in v.items():
                    if kk in dst:
                        dst[kk].copy_(vv, non_blocking=non_blocking)

    def run(self):
        """Replay on the current stream; returns the (static) loss tensor.  Gradients are in grads.flat."""
        self.replays += 1
        self.graph.replay()
        return self.loss
