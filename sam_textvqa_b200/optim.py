"""Clip + Adam on flat buffers: the optimizer side of the training step.

The reference's loop (train.py:139-143) is `clip_gradients` (sam/task_utils.py:33-34: `clip_grad_norm_` over every
parameter) followed by `torch.optim.Adam(optimizer_grouped_parameters, lr=base_lr)` (task_utils.py:39-42, torch
defaults) and the warm-up `LambdaLR` (:43-57).  With all gradients in one `dp.FlatGradBuffer` both become a handful of
launches: one sum of squares over the flat gradient, then one fused clip + Adam update per parameter group (each group
is a contiguous range of the flat layout, so it has a single learning rate).  Parameters, `exp_avg` and `exp_avg_sq`
live in flat fp32 buffers of the same layout; every `p.data` is a view, so `state_dict()` / `load_state_dict()` keep
working.  Construct it BEFORE `graph_step.GraphedTrainStep`: it moves the parameters into the flat buffer, and a
captured step reads them at the addresses they had during capture.  The arithmetic follows torch.optim.Adam operation
by operation (GPU test against it).
"""
import torch

from ._lib import check, lib, ptr, stream_ptr


def flat_grad_buffer_for(param_groups):
    """A `dp.FlatGradBuffer` whose layout keeps every param group of `get_optimizer_parameters` contiguous."""
    from .dp import FlatGradBuffer
    seen, ordered = set(), []
    for g in param_groups:
        for p in g["params"]:
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                ordered.append(p)
    return FlatGradBuffer(ordered)


class FlatAdam(object):
    """param_groups: list of {"params": [...], ["lr": float]} as returned by `get_optimizer_parameters`
    (sa_m4c.py:349-371; its first group carries no "lr" and takes the optimizer default `lr`, like torch.optim.Adam);
    grads: the FlatGradBuffer that holds their gradients in the same order (see `flat_grad_buffer_for`)."""

    def __init__(self, param_groups, grads, lr=None, betas=(0.9, 0.999), eps=1e-8, max_grad_norm=None):
        self.grads = grads
        self.betas, self.eps = (float(betas[0]), float(betas[1])), float(eps)
        self.max_grad_norm = None if max_grad_norm is None else float(max_grad_norm)
        self.step_count = 0
        index = {id(p): i for i, p in enumerate(grads.params)}
        offs, off = [], 0
        for p in grads.params:
            offs.append(off)
            off += p.numel()
        self.param_groups = []
        cursor = 0
        for g in param_groups:
            ps = [p for p in g["params"] if p.requires_grad]
            if not ps:
                continue
            ids = [index[id(p)] for p in ps]
            if ids != list(range(cursor, cursor + len(ids))):
                raise ValueError("param groups must be contiguous runs of the FlatGradBuffer layout "
                                 "(build it with optim.flat_grad_buffer_for)")
            begin = offs[ids[0]]
            end = offs[ids[-1]] + ps[-1].numel()
            group_lr = g.get("lr", lr)
            if group_lr is None:
                raise ValueError("a param group without 'lr' needs the optimizer default: FlatAdam(..., lr=base_lr)")
            self.param_groups.append({"params": ps, "lr": float(group_lr), "initial_lr": float(group_lr),
                                      "range": (begin, end)})
            cursor += len(ids)
        if cursor != len(grads.params):
            raise ValueError("every parameter of the gradient buffer must belong to a param group")
        dev = grads.flat.device
        if dev.type != "cuda":
            raise RuntimeError("FlatAdam runs on the GPU (samk_adam_step); got %s" % dev)
        self.flat_params = torch.empty_like(grads.flat)
        for p, o in zip(grads.params, offs):
            view = self.flat_params[o:o + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
        self.exp_avg = torch.zeros_like(grads.flat)
        self.exp_avg_sq = torch.zeros_like(grads.flat)
        self._sumsq = torch.zeros(1, dtype=torch.float64, device=dev)

    def grad_norm(self):
        """Global L2 norm of the gradients as a 0-d device tensor (what clip_grad_norm_ returns)."""
        self._sumsq.zero_()
        check(lib().samk_sumsq(ptr(self.grads.flat), self.grads.flat.numel(), ptr(self._sumsq), stream_ptr()), "sumsq")
        return self._sumsq.sqrt().float()[0]

    def _check_grad_views(self):
        """`model.zero_grad()` of torch 2.x sets .grad to None (train.py:144): the next backward then allocates fresh
        gradients while this optimizer reads the flat buffer.  Re-attach the views (a detached gradient that already
        holds values is copied in) so the reference loop keeps working."""
        off = 0
        flat = self.grads.flat
        for p in self.grads.params:
            n = p.numel()
            g = p.grad
            if g is None:
                p.grad = flat[off:off + n].view_as(p)
            elif g.data_ptr() != flat.data_ptr() + 4 * off:
                flat[off:off + n].view_as(p).copy_(g)
                p.grad = flat[off:off + n].view_as(p)
            off += n

    def state_dict(self):
        """torch.optim.Adam-compatible: per-parameter `step`, `exp_avg`, `exp_avg_sq` (views of the flat buffers) and the
        param groups (train.py:179-181 saves optimizer_state_dict)."""
        state, groups, idx, off = {}, [], 0, 0
        offs = {}
        for p in self.grads.params:
            offs[id(p)] = off
            off += p.numel()
        for g in self.param_groups:
            ids = []
            for p in g["params"]:
                o, n = offs[id(p)], p.numel()
                state[idx] = {"step": torch.tensor(float(self.step_count)), "exp_avg": self.exp_avg[o:o + n].view_as(p),
                              "exp_avg_sq": self.exp_avg_sq[o:o + n].view_as(p)}
                ids.append(idx)
                idx += 1
            groups.append({"lr": g["lr"], "initial_lr": g["initial_lr"], "betas": self.betas, "eps": self.eps, "weight_decay": 0,
                           "amsgrad": False, "params": ids})
        return {"state": state, "param_groups": groups, "max_grad_norm": self.max_grad_norm}

    def load_state_dict(self, sd):
        idx, off = 0, 0
        offs = {}
        for p in self.grads.params:
            offs[id(p)] = off
            off += p.numel()
        steps = []
        for g, sg in zip(self.param_groups, sd["param_groups"]):
            g["lr"] = float(sg["lr"])
            g["initial_lr"] = float(sg.get("initial_lr", sg["lr"]))
            for p in g["params"]:
                st = sd["state"].get(idx, sd["state"].get(str(idx)))
                if st is not None:
                    o, n = offs[id(p)], p.numel()
                    self.exp_avg[o:o + n].view_as(p).copy_(st["exp_avg"])
                    self.exp_avg_sq[o:o + n].view_as(p).copy_(st["exp_avg_sq"])
                    steps.append(int(float(st["step"])))
                idx += 1
        if steps:
            if len(set(steps)) != 1:
                raise ValueError("FlatAdam keeps one step count for all parameters; the state has %s" % sorted(set(steps)))
            self.step_count = steps[0]

    def step(self):
        """clip_gradients + optimizer.step() of train.py:139-142.  The gradients are left unscaled."""
        self._check_grad_views()
        self.step_count += 1
        sumsq = None
        if self.max_grad_norm is not None:
            self._sumsq.zero_()
            check(lib().samk_sumsq(ptr(self.grads.flat), self.grads.flat.numel(), ptr(self._sumsq), stream_ptr()), "sumsq")
            sumsq = self._sumsq
        for g in self.param_groups:
            b, e = g["range"]
            check(lib().samk_adam_step(ptr(self.flat_params[b:e]), ptr(self.grads.flat[b:e]), ptr(self.exp_avg[b:e]),
                                       ptr(self.exp_avg_sq[b:e]), e - b, g["lr"], self.betas[0], self.betas[1], self.eps,
                                       self.step_count, ptr(sumsq) if sumsq is not None else None,
                                       self.max_grad_norm if self.max_grad_norm is not None else 0.0, stream_ptr()),
                  "adam_step")
        # the kernels wrote through raw pointers: tell autograd-side caches (ops.weight_operand keys its bf16 operand
        # copies on the parameters' version counters) that the parameters changed
        torch.autograd.graph.increment_version(self.grads.params)

    def zero_grad(self):
        self.grads.zero()


def warmup_lr_lambda(warmup_iters, warmup_factor, lr_decay_iters=(), lr_decay=1.0):
    """The multiplier of sam/task_utils.py:46-55: linear warm-up from `warmup_factor` to 1 over `warmup_iters`
    updates, then `lr_decay ** (number of decay milestones passed)`."""
    def fn(it):
        if it <= warmup_iters:
            alpha = float(it) / float(warmup_iters)
            return warmup_factor * (1.0 - alpha) + alpha
        idx = 0
        for m in sorted(lr_decay_iters):
            if it >= m:
                idx += 1
        return pow(lr_decay, idx)
    return fn


def set_lr_multiplier(optimizer, mult):
    for g in optimizer.param_groups:
        g["lr"] = g["initial_lr"] * mult
