#!/usr/bin/env python
"""SA-M4C fwd+bwd samples/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl samk|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one synthetic batch: SAM4C forward in train mode (all
dropouts at the yml's 0.1) -> masked BCE loss -> backward (with the gradient exchange of N > 1 ranks:
samk_exchange_sum over NVLink peer memory, bucket by bucket under the backward pass; SAMK_DP_TRANSPORT=nccl = one
ncclAllReduce after the step) -> zero_grad; by default the relation graph is built on the device inside the step.  Workload at every N: BASELINE config[1], the shipped c3 experiment
(layers n,n,s,s,s,s; 20 text + 100 obj + 50 OCR + 12 dec tokens; d=768; V=5000), 128 samples per GPU
(weak scaling).  `value` is timed with the inputs resident in HBM; `e2e` runs the same step from
pinned host buffers with the host->device copies and a device->host read of the loss inside the
timed region.

The line also carries
  parity    the precision mode that was TIMED, checked on this box against the reference-generated golden of the
            shipped c3 stack (tests/golden/sam4c_c3.npz): logits rel err and argmax identity, plus the other mode;
  extras    (N = 1) throughput of the other precision mode, a full optimizer step, greedy decoding, BASELINE
            configs 0 / 2, the spatial-graph kernel (pairs/s, GB/s);
  roofline  all tcgen05 GEMM launches of one step against the sustained bf16 peak, timed with CUDA events recorded inside a
            captured step (ops.TimingEvent), per-shape detail with the DRAM bytes ncu measured (profiles/ncu_traffic.json,
            written by tools/ncu_summary.py from one ncu pass over a step); `attention`: the same for the north-star kernel.

`--impl reference` times the UNMODIFIED reference (oracle/_ref: a verbatim, git-ignored copy of /root/reference/sam
made by `python -m oracle.build_ref`, run through the oracle/shim stand-ins for its two un-vendored dependencies) on the
host cores: same model, same batch recipe with the relation graph from the reference's own builder, train mode, the
reference's own loss; every step processes a bounded sample (--ref-batch rows of the 128-row batch).  Without the copy
it falls back to the oracle port and says so (`kind`).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_SAMPLE = 52.356e9      # fwd+bwd, cfg2, torch FlopCounter on the reference
METRIC = "SA-M4C fwd+bwd samples/sec"
UNIT = "samples/s"
CFG = dict(T=20, O=100, R=50, D=12, V=5000)
WORKLOAD = "c3 yml SA-M4C (n,n,s,s,s,s), 20+100+50+12 tokens, d=768, V=5000, train fwd+bwd, dropout 0.1"   # both arms


def config_dict(world, B, launch=None):
    cfg = {"workload": WORKLOAD, "batch_per_gpu": B, "global_batch": world * B, "parallelism": "dp%d" % world}
    if launch:
        cfg["launch"] = launch
        cfg["l2"] = "per-step working set (~6 GB of activations) >> 126 MB L2; no explicit flush"
    return cfg


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1397.0), d.get("hbm_gbs", 6514.2), "measured"
    return 1400.0, 6650.0, "fallback"


class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------
# CPU arm: the unmodified reference (oracle/_ref) or, without it, the oracle port
# ---------------------------------------------------------------------------------------------------
def _ref_graph_one(args):
    """(worker) relation types of one sample from the reference's own builder, sam/spatial_utils.py:92-218"""
    root, shim, boxes = args
    import warnings
    warnings.filterwarnings("ignore")
    for p in (root, shim):
        if p not in sys.path:
            sys.path.insert(0, p)
    import sam.spatial_utils as S
    return S.build_graph_using_normalized_boxes(boxes)["1"]


def cpu_reference_step_fn(B, seed, threads):
    """Returns (step, kind, description): one zero_grad + fwd + loss + bwd of the reference on `B` samples."""
    import numpy as np
    import torch
    from oracle import ref_loader
    from sam_textvqa_b200 import synth
    from sam_textvqa_b200.config import c3_config
    torch.set_num_threads(threads)
    mmt, tb = c3_config()
    if not ref_loader.reference_available():
        return cpu_port_step_fn(B, seed, threads)
    import warnings
    warnings.filterwarnings("ignore")
    M, S, registry = ref_loader.load_reference(CFG["V"])

    def graph_fn(boxes):            # the dataset side of the reference builds these in a process pool as well
        import multiprocessing as mp
        jobs = [(ref_loader.REF_ROOT, ref_loader.SHIM_DIR, b) for b in np.asarray(boxes)]
        with mp.get_context("fork").Pool(max(1, min(threads, len(jobs)))) as pool:
            return np.stack(pool.map(_ref_graph_one, jobs))

    batch = synth.make_batch(B, seed=seed, contexts=(1, 3), graph_fn=graph_fn, **CFG)
    torch.manual_seed(0)
    model = M.SAM4C(M.BertConfig.from_dict(mmt), M.BertConfig.from_dict(tb)).train()
    try:
        from sam.task_utils import M4CDecodingBCEWithMaskLoss       # sam/task_utils.py:19-30, unmodified
        loss_fn, loss_src = M4CDecodingBCEWithMaskLoss(), "sam.task_utils.M4CDecodingBCEWithMaskLoss"
    except Exception as exc:                                          # a dataset-side import missing on this host
        from oracle import sam4c_oracle
        loss_fn, loss_src = sam4c_oracle.bce_with_mask_loss, "oracle restatement of the loss (%s)" % type(exc).__name__
    keys = [k for k, v in batch.items() if torch.is_tensor(v) and k not in ("boxes", "spatial_types")]

    def step():
        model.zero_grad()
        bd = {k: batch[k] for k in keys}
        bd["spatial_adj_matrices"] = dict(batch["spatial_adj_matrices"])
        scores = model(bd)["textvqa_scores"]
        loss = loss_fn(scores, batch["targets"], batch["train_loss_mask"])
        loss.backward()
        return float(loss.detach())

    desc = ("unmodified reference SAM4C (%s via oracle/shim), torch CPU fp32, train mode, %s, relation graph from the "
            "reference builder, %d of the 128 samples of the batch per step" % (ref_loader.REF_ROOT, loss_src, B))
    return step, "reference", desc


def cpu_port_step_fn(B, seed, threads):
    """Fallback when oracle/_ref is absent: the oracle port (same op sequence as the reference)."""
    import numpy as np
    import torch
    from oracle import sam4c_oracle
    from sam_textvqa_b200 import synth
    from sam_textvqa_b200.config import c3_config
    from tests._util import sam4c_state_shapes
    torch.set_num_threads(threads)
    mmt, tb = c3_config()
    rs = np.random.RandomState(seed)

    def rand_types(boxes):   # random types are allowed for the CPU timing leg only
        return rs.randint(0, 13, (boxes.shape[0], boxes.shape[1], boxes.shape[1])).astype(np.int8)

    batch = synth.make_batch(B, seed=seed, contexts=(3,), graph_fn=rand_types, **CFG)
    P = synth.seeded_state(sam4c_state_shapes(mmt, tb, CFG["V"]), 0)
    P = {k: v.requires_grad_(True) for k, v in P.items()}

    def step():
        scores, _, _ = sam4c_oracle.forward(P, batch, mmt, tb, train=True)
        loss = sam4c_oracle.bce_with_mask_loss(scores, batch["targets"], batch["train_loss_mask"])
        loss.backward()
        for v in P.values():
            v.grad = None
        return float(loss.detach())

    return step, "port", ("oracle port (oracle/_ref missing): torch CPU fp32, reference op sequence incl. dense masks, "
                          "unique-check, dropout; random relation types; B=%d" % B)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.ref_batch
    step, kind, desc = cpu_reference_step_fn(B, 0, threads)
    for _ in range(args.warmup):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / max(args.steps, 1)
    val = B / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(args.gpus, args.batch),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def parity_check(dev):
    """Logits of the shipped c3 stack against the golden the UNMODIFIED reference produced for this very model and
    batch (tests/golden/sam4c_c3.npz; generator oracle/make_golden.py), in the current precision mode."""
    import numpy as np
    import torch
    from sam_textvqa_b200 import ops, spatial_utils, synth
    from sam_textvqa_b200.config import c3_config
    from sam_textvqa_b200.registry import registry
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig
    from tests._util import rel_err, sam4c_state_shapes
    gold = np.load(os.path.join(ROOT, "tests", "golden", "sam4c_c3.npz"))
    V = 500
    saved = registry.get("answer_vocab")
    registry.answer_vocab = ["w%d" % i for i in range(V)]
    try:
        mmt, tb = c3_config(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, obj_drop=0.0, ocr_drop=0.0)
        tb = dict(tb, hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
        model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb))
        model.load_state_dict(synth.seeded_state(sam4c_state_shapes(mmt, tb, V), 1), strict=True)
        model = model.to(dev).train()
        graph_fn = lambda boxes: spatial_utils.build_graph_batch(boxes, 0.5)[0]
        batch = synth.make_batch(3, V=V, seed=5, contexts=(1, 3), graph_fn=graph_fn)
        types_ok = bool(np.array_equal(batch["spatial_types"].numpy(), gold["types"]))    # CUDA graph kernel == reference builder
        ref = torch.from_numpy(gold["tf/scores"])
        live = ref > -5000
        out = {}
        cur = ops.get_precision()
        for mode in (cur, "bf16x3" if cur != "bf16x3" else "f16"):
            ops.set_precision(mode)
            ops.clear_weight_cache()
            bd = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
            bd["spatial_adj_matrices"] = {k: v.to(dev) for k, v in batch["spatial_adj_matrices"].items()}
            with torch.no_grad():
                scores = model(bd)["textvqa_scores"].float().cpu()
            out[mode] = {"logits_rel_err": rel_err(scores, ref, live),
                         "argmax_match": bool(torch.equal(scores.argmax(-1), ref.argmax(-1)))}
        ops.set_precision(cur)
        ops.clear_weight_cache()
        del model
        res = {"mode": cur, "logits_rel_err": out[cur]["logits_rel_err"], "argmax_match": out[cur]["argmax_match"],
               "tolerance": 1e-3, "golden": "tests/golden/sam4c_c3.npz (unmodified reference, c3 stack, L=182, B=3)",
               "graph_types_bit_exact": types_ok}
        other = [m for m in out if m != cur][0]
        res["other_mode"] = dict(out[other], mode=other)
        return res
    finally:
        registry.answer_vocab = saved


def bind_to_gpu_numa_node(local):
    """One process per GPU: run on the CPUs of the GPU's NUMA node, so that the pinned staging buffers (first touch) and
    the host->device copies stay on the socket the GPU hangs off.  Eight ranks uploading 212 MB per step each otherwise
    share whatever node the processes happened to start on.  Returns a description, or None when the topology is unknown."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(local)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open(path + "numa_node").read().strip())
        cpus = set()
        for part in open(path + "local_cpulist").read().strip().split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if node < 0 or not cpus or cpus == os.sched_getaffinity(0):
            return None
        os.sched_setaffinity(0, cpus)
        return "numa node %d (%d cpus)" % (node, len(cpus))
    except Exception:
        return None


def run_samk(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    host_affinity = bind_to_gpu_numa_node(local) if world > 1 and os.environ.get("SAMK_BENCH_NUMA", "1") == "1" else None
    # SAMK_DP_OVERLAP=1: the gradient exchange runs bucket by bucket under the backward pass (captured in the graph).
    # NCCL is capped at a few CTAs and the persistent kernels leave that many SMs out of their grids.
    # Measured at N=2 (profiles/r02_dp_experiments.txt): 13.23 ms/step against 12.64 ms for the plain exchange after the
    # step -- the 8 reserved SMs cost 0.25 ms and the last bucket (TextBERT + its 94 MB embedding table, final only when
    # the backward pass ends) stays exposed at the capped NCCL bandwidth -- so it is opt-in.
    # SAMK_DP_TRANSPORT=peer: the same bucketed exchange, but through samk_exchange_sum (csrc/exchange.cu: NVSwitch
    # multicast / NVLink peer memory, blocks small enough to share SMs with the backward GEMMs) -- no SM reservation.
    transport = os.environ.get("SAMK_DP_TRANSPORT", "peer")
    overlap = world > 1 and (os.environ.get("SAMK_DP_OVERLAP", "0") == "1" or transport == "peer")
    if world > 1:
        os.environ.setdefault("SAMK_DP_WIRE", "bf16")     # 193 MB on the wire instead of 387 MB (SAMK_DP_WIRE=f32: exact sum)
    reserve = int(os.environ.get("SAMK_DP_RESERVE_SMS", "8")) if transport != "peer" else 0
    if overlap and transport != "peer":
        os.environ.setdefault("NCCL_MAX_CTAS", str(reserve))
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from sam_textvqa_b200 import build as samk_build
    samk_build.build()
    from sam_textvqa_b200 import dp, ops, optim, spatial_utils, synth
    from sam_textvqa_b200.config import c3_config
    from sam_textvqa_b200.registry import registry
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig

    ops.set_precision(args.precision)
    parity = parity_check(dev) if rank == 0 and not args.profile_only else None
    registry.answer_vocab = ["w%d" % i for i in range(CFG["V"])]
    registry.BOS_IDX = 1
    mmt, tb = c3_config()
    torch.manual_seed(0)
    model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb)).to(dev).train()
    # gradients and parameters in flat buffers laid out by learning-rate group (get_optimizer_parameters,
    # sa_m4c.py:349-371); the fused clip + Adam of the `extras` leg works on them (built before the graph capture:
    # it moves the parameters)
    groups = model.get_optimizer_parameters(1e-4)
    grads = optim.flat_grad_buffer_for(groups)
    opt = optim.FlatAdam(groups, grads, lr=1e-4, max_grad_norm=0.25)
    if overlap:
        wire_dt = torch.bfloat16 if os.environ.get("SAMK_DP_WIRE", "bf16") == "bf16" else torch.float32
        try:
            grads.enable_overlap(average=False, bucket_bytes=int(os.environ.get("SAMK_DP_BUCKET_MB", "32")) << 20,
                                 transport=transport, wire_dtype=wire_dt)
            ok = torch.ones(1, device=dev)
        except Exception as exc:               # e.g. no peer access / symmetric memory on this box: say so, use the library path
            print("bench: peer exchange unavailable on rank %d (%r)" % (rank, exc), file=sys.stderr, flush=True)
            ok = torch.zeros(1, device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok) == 0.0:                   # all ranks take the same path
            ops.grad_ready_hook = None
            if hasattr(grads, "_ov"):
                del grads._ov
            overlap, transport, reserve = False, "nccl", 0
        if reserve:
            from sam_textvqa_b200._lib import lib as _samk_lib
            _samk_lib().samk_reserve_sms(reserve)
    exchange = dp.GradExchange(grads, world, overlapped=overlap) if world > 1 else None
    B = args.batch

    def graph_fn(boxes):
        return spatial_utils.build_graph_batch(boxes.astype("float32"), 0.5)[0]

    host = synth.make_batch(B, seed=rank, contexts=(3,), graph_fn=graph_fn, **CFG)
    host.pop("boxes"); host.pop("spatial_types")
    adj = host.pop("spatial_adj_matrices")["3"]
    names = sorted(k for k, v in host.items() if torch.is_tensor(v))
    pinned = {k: host[k].pin_memory() for k in names}
    pinned_adj = adj.pin_memory()
    # On-device batch preparation (sa_m4c._relation_bits): the step builds the relation graph and the head
    # bits on the GPU from the padded boxes the batch carries anyway (bit-identical to the reference builder: parity
    # object, tests/test_gpu_graph.py), so the int8 [B,150,150,12] matrix (34.6 MB per step) is neither uploaded nor
    # kept resident.  SAMK_BENCH_DEVICE_GRAPH=0: the reference contract (matrix prepared by the caller, uploaded).
    dev_graph = os.environ.get("SAMK_BENCH_DEVICE_GRAPH", "1") == "1"
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values()) + (0 if dev_graph else pinned_adj.numel())
    resident = {k: v.to(dev) for k, v in pinned.items()}
    resident_adj = pinned_adj.to(dev)
    # Data-parallel loss normalisation: the reference divides by the GLOBAL number of valid steps (DataParallel gathers
    # the scores, task_utils.py:118-129).  Every rank scales its loss by local_count / global_count and the exchange
    # SUMS the gradients: no divide pass over the 387 MB buffer.
    loss_scale = dp.global_loss_scale(resident["train_loss_mask"]) if world > 1 else None

    def loss_of(scores, b):
        loss = ops.bce_with_mask_loss(scores, b["targets"], b["train_loss_mask"])
        return loss * loss_scale if loss_scale is not None else loss

    def fwd_bwd(inputs, adj_dev, exchange_inside=True):
        grads.zero()
        if overlap and exchange_inside:
            grads.begin_step()
        bd = dict(inputs)
        if not dev_graph:
            bd["spatial_adj_matrices"] = {"3": adj_dev}
        loss = loss_of(model(bd)["textvqa_scores"], inputs)
        loss.backward()
        if overlap and exchange_inside:
            grads.finish_step()
        return loss

    # The step as a user runs it: captured once into a CUDA graph (sam_textvqa_b200/graph_step.py) and replayed --
    # ~340 launches through ctypes + autograd cost the host about as long as the GPU needs for the step.
    # SAMK_BENCH_EAGER=1 keeps the eager loop.
    graphed = None
    if os.environ.get("SAMK_BENCH_EAGER", "0") != "1":
        try:
            from sam_textvqa_b200.graph_step import GraphedTrainStep
            ex = dict(resident)
            if not dev_graph:
                ex["spatial_adj_matrices"] = {"3": resident_adj}
            graphed = GraphedTrainStep(model, grads, ex, loss_fn=loss_of, allreduce="overlap" if overlap else None)
        except Exception as exc:                      # fall back loudly, never silently
            print("bench: CUDA-graph capture failed (%r); running the eager step" % (exc,), file=sys.stderr, flush=True)
            graphed = None

    def step(inputs, adj_dev, do_exchange=True):
        if graphed is None:
            loss = fwd_bwd(inputs, adj_dev)
        else:
            if inputs is not resident:                # fresh upload: device-to-device copy into the graph's input buffers
                graphed.load(inputs)
                if not dev_graph:
                    graphed.load({"spatial_adj_matrices": {"3": adj_dev}})
            loss = graphed.run()
        if exchange is not None and do_exchange:
            exchange.all_reduce()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item() / steps

    if args.profile_only:      # for `ncu` launch lists: 1 warm-up + 2 steps, nothing else
        log_path = os.environ.get("SAMK_LAUNCH_LOG")
        for i in range(3):
            if log_path and graphed is None and i == 2:      # eager run: shapes of the last step's launches, in order
                ops.launch_log = []
            step(resident, resident_adj)
        torch.cuda.synchronize()
        if log_path and ops.launch_log is not None:
            json.dump(ops.launch_log, open(log_path, "w"))
        return
    warm = max(args.warmup, 3)
    for _ in range(warm):
        step(resident, resident_adj)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.launch_count
    ms = timed(lambda: step(resident, resident_adj), args.steps)
    launches = ops.launch_count - l0
    if graphed is not None:                    # replays do not pass through the Python counters
        launches = (graphed.kernels_per_replay + (exchange.kernels_per_call if exchange else 0)) * args.steps
    if overlap and grads._ov.get("peer") is not None:
        torch.cuda.synchronize()
        grads._ov["peer"].check()              # a rank that missed a barrier would have produced garbage, not a number

    # ---- end to end: every step's inputs come from pinned host memory; the copy of step i+1 runs on a
    # side stream while step i computes (double-buffered device staging), and every step's loss is read
    # back to the host.  All K uploads and K loss reads are inside the timed region.
    # With the captured step the uploads land DIRECTLY in the step's static input buffers: the step is captured twice
    # (two GraphedTrainStep objects, each with its own input buffers) and the replays alternate, so step i+1's upload
    # runs while step i computes without a device-to-device hop in between (SAMK_BENCH_E2E_GRAPHS=1: one captured step fed
    # through double-buffered staging + a device-to-device load, the round-1 form).
    copy_stream = torch.cuda.Stream()
    two_graphs = graphed is not None and os.environ.get("SAMK_BENCH_E2E_GRAPHS", "2") == "2"
    graphs = [graphed]
    if two_graphs:
        graphs.append(GraphedTrainStep(model, grads, ex, loss_fn=loss_of, allreduce="overlap" if overlap else None))
        staged = None
    else:
        staged = [({k: torch.empty_like(v) for k, v in resident.items()}, torch.empty_like(resident_adj)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in consumed:
        ev.record()

    def stage(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            if two_graphs:
                graphs[slot].load(pinned)                        # host (pinned) -> the static inputs of that captured step
                if not dev_graph:
                    graphs[slot].load({"spatial_adj_matrices": {"3": pinned_adj}})
            else:
                bufs, abuf = staged[slot]
                for k in names:
                    bufs[k].copy_(pinned[k], non_blocking=True)
                if not dev_graph:
                    abuf.copy_(pinned_adj, non_blocking=True)
            ready[slot].record(copy_stream)

    host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    loss_values = []

    def e2e_run(steps):
        stage(0)
        for i in range(steps):
            torch.cuda.current_stream().wait_event(ready[i % 2])
            if i + 1 < steps:
                stage((i + 1) % 2)
            if two_graphs:
                loss = graphs[i % 2].run()
                consumed[i % 2].record()              # (the replay reads its inputs until it ends)
            elif graphed is not None:                 # inputs are taken by the device-to-device load in front of the replay
                up, a = staged[i % 2]
                graphed.load(up)
                if not dev_graph:
                    graphed.load({"spatial_adj_matrices": {"3": a}})
                consumed[i % 2].record()
                loss = graphed.run()
            else:
                up, a = staged[i % 2]
                loss = fwd_bwd(up, a)
                consumed[i % 2].record()
            if exchange is not None:
                exchange.all_reduce()
            # device->host read of EVERY step's loss: an async copy into pinned memory plus an event right behind the
            # step; the host reads step i-1's value after it has enqueued step i, waiting on that event only
            host_loss[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_done[i % 2].record()
            if i >= 1:
                loss_done[(i - 1) % 2].synchronize()
                loss_values.append(float(host_loss[(i - 1) % 2][0]))
        if steps >= 1:
            loss_done[(steps - 1) % 2].synchronize()
            loss_values.append(float(host_loss[(steps - 1) % 2][0]))

    e2e_run(max(args.warmup, 5))      # the host->device path (pinned pages, PCIe link state) needs its own warm-up
    barrier()
    s_ev, e_ev = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s_ev.record()
    e2e_run(args.steps)
    e_ev.record()
    barrier()
    ms_t = torch.tensor([s_ev.elapsed_time(e_ev)], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms_e2e = ms_t.item() / args.steps
    clocks = sampler.stop() if sampler else None

    # exposed part of the gradient exchange: the same K steps without it (after the headline measurements)
    allreduce_exposed_ms = None
    if exchange is not None and graphed is not None:
        if overlap:          # the exchange is part of the captured step: capture the step once more without it
            from sam_textvqa_b200 import ops as _ops
            hook, _ops.grad_ready_hook = _ops.grad_ready_hook, None
            plain = GraphedTrainStep(model, grads, ex, loss_fn=loss_of, allreduce=None)
            for _ in range(3):
                plain.run()
            ms_noex = timed(plain.run, args.steps)
            _ops.grad_ready_hook = hook
            del plain
        else:
            ms_noex = timed(lambda: step(resident, resident_adj, do_exchange=False), args.steps)
        allreduce_exposed_ms = ms - ms_noex

    # ---- roofline of the dominant kernel family (tcgen05 GEMM): per-launch CUDA events INSIDE a captured step ----
    # The step is captured once more with an event pair around every GEMM and attention launch (ops.TimingEvent:
    # cudaEventRecordExternal nodes of the graph, on the stream the kernels are launched on) and replayed: the durations
    # are those of the kernels as they run in the replayed step, side branches and all.  Fallback (capture unavailable):
    # two eager steps behind a spin kernel, as in round 1.
    saved_hook, ops.grad_ready_hook = ops.grad_ready_hook, None      # no exchange in the instrumented steps
    prof = aprof = None
    n_inst = 1
    roofline_how = "CUDA events recorded inside a captured step (event-record nodes), one replay"
    if graphed is not None and os.environ.get("SAMK_BENCH_EVENTS", "graph") == "graph":
        try:
            ops.profile_in_graph = True
            ops.gemm_profile, ops.attn_profile = [], []
            inst = GraphedTrainStep(model, grads, ex, loss_fn=loss_of, allreduce=None, warmup=1)
            for _ in range(3):
                inst.run()
            torch.cuda.synchronize()
            per_g, per_a = len(ops.gemm_profile) // 2, len(ops.attn_profile) // 2      # (1 eager warm-up step + the capture)
            prof, aprof = ops.gemm_profile[-per_g:], ops.attn_profile[-per_a:]
            _ = sum(s_.elapsed_time(e_) for s_, e_, _, _ in prof)                      # raises if the events are unreadable
            del inst
        except Exception as exc:
            print("bench: in-graph event timing unavailable (%r); instrumented eager steps instead" % (exc,), file=sys.stderr, flush=True)
            prof = aprof = None
        finally:
            ops.profile_in_graph = False
            ops.gemm_profile = ops.attn_profile = None
    if prof is None:
        n_inst = 2
        roofline_how = "CUDA events around every launch of two eager steps queued behind a spin kernel"
        ops.gemm_profile = []
        ops.attn_profile = []
        for _ in range(n_inst):
            # the eager loop is launch-bound: an event pair around a kernel the GPU is waiting for also times the host's
            # launch latency.  A spin kernel in front keeps the GPU busy while the host enqueues the whole step.
            try:
                torch.cuda._sleep(int(4e7))        # ~20 ms at 1.9 GHz; the host needs ~10-15 ms to enqueue a step
            except Exception:                      # private torch API: without it the figures are only more pessimistic
                pass
            fwd_bwd(resident, resident_adj, exchange_inside=False)
            torch.cuda.synchronize()
        prof, ops.gemm_profile = ops.gemm_profile, None
        aprof, ops.attn_profile = ops.attn_profile, None
    ops.grad_ready_hook = saved_hook
    g_ms = sum(s.elapsed_time(e) for s, e, _, _ in prof) / n_inst
    g_flop = sum(f for _, _, f, _ in prof) / n_inst
    n_gemm = len(prof) // n_inst
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # DRAM bytes per launch measured by `ncu --set full` for the big shapes (tools/ncu_summary.py --json)
    ncu_traffic = {}
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            ncu_traffic = json.load(open(tpath))
        except Exception:
            ncu_traffic = {}
    groups = {}
    for s_, e_, f_, shape in prof:
        g = groups.setdefault(shape, [0.0, 0, f_])
        g[0] += s_.elapsed_time(e_)
        g[1] += 1
    by_shape = []
    tr_sum, tr_launches, tr_ms = 0.0, 0, 0.0
    for shape, (tot_ms, cnt, f_) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        us = tot_ms / cnt * 1e3
        tf = f_ / (us * 1e-6) / 1e12
        Mg, Ng, Kg, amn, bmn = shape
        item = {"M": Mg, "N": Ng, "K": Kg, "a_mn": amn, "b_mn": bmn, "launches_per_step": cnt // n_inst,
                "us_per_launch": us, "achieved": tf, "frac": tf / peak_tf, "share_of_step": tot_ms / n_inst / ms if ms > 0 else None,
                "operand_MB": (Mg * Kg + Ng * Kg) * 2 / 1e6}
        key = "gemm_%dx%dx%d_%d%d" % (Mg, Ng, Kg, int(amn), int(bmn))
        if key in ncu_traffic and B == 128:
            item["traffic"] = ncu_traffic[key]["dram_bytes"]
            item["traffic_src"] = ncu_traffic[key].get("src")
            tr_sum += item["traffic"] * (cnt // n_inst)
            tr_launches += cnt // n_inst
            tr_ms += tot_ms / n_inst
        if len(by_shape) < 6:
            by_shape.append(item)
    # the north-star kernel: fused masked attention of the MMT layers (L = 182), HBM-bound at this length
    Lm = CFG["T"] + CFG["O"] + CFG["R"] + CFG["D"]
    attention = {}
    for kind in ("fwd", "bwd"):
        rows = [(s.elapsed_time(e), nb, fl) for k, L_, s, e, nb, fl in aprof if k == kind and L_ == Lm]
        if rows:
            ms_k = sum(r[0] for r in rows) / len(rows)
            attention[kind] = {"us_per_launch": ms_k * 1e3, "launches_per_step": len(rows) // n_inst,
                               "algorithmic_MB": rows[0][1] / 1e6, "achieved_GBs": rows[0][1] / (ms_k * 1e-3) / 1e9,
                               "hbm_frac": rows[0][1] / (ms_k * 1e-3) / 1e9 / peak_gbs,
                               "dense_equiv_TFLOPs": rows[0][2] / (ms_k * 1e-3) / 1e12}
            key = "attn_%s_L%d" % (kind, Lm)
            if key in ncu_traffic and B == 128:
                attention[kind]["traffic"] = ncu_traffic[key]["dram_bytes"]
                attention[kind]["tensor_pipe_pct"] = ncu_traffic[key].get("tensor_pipe_pct")
                attention[kind]["traffic_src"] = ncu_traffic[key].get("src")
    attention["peak_GBs"] = peak_gbs
    attention["note"] = "L=182 is HBM/ALU-bound (SURVEY 8d); the bwd launch includes the dO preparation kernel"

    if rank != 0:
        if world > 1:
            del graphed
            torch.cuda.synchronize()
            dist.destroy_process_group()
        return
    value = world * B / (ms * 1e-3)
    e2e_val = world * B / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": warm, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16" if args.precision == "f16" else "bf16x3", "data": "synthetic",
        "config": config_dict(world, B, ("cuda-graph replay of the captured step" if graphed is not None else "eager") +
                              ("; relation graph + head bits built on the device inside the step from the padded boxes"
                               if dev_graph else "; relation matrix prepared by the caller")),
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "parity": parity,
        "roofline": {"bound": "tensor", "kernel": "samk::gemm_tc_kernel (all %d GEMM launches of one step)" % n_gemm,
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": (tr_sum / tr_launches) if tr_launches else None,
                     "traffic_note": ("mean DRAM bytes per launch over the %d launches (%.0f %% of the GEMM time) whose shape has "
                                      "an ncu --set full capture in profiles/ncu_traffic.json" % (tr_launches, 100 * tr_ms / g_ms))
                     if tr_launches else "no profiles/ncu_traffic.json",
                     "peak_source": peak_src + " (sustained bf16)", "timing": roofline_how,
                     "gemm_ms_per_step": g_ms, "gemm_share_of_step": g_ms / ms if ms > 0 else None,
                     "step_flop_frac_of_peak": value / world * FLOP_PER_SAMPLE / (peak_tf * 1e12),
                     "by_shape": by_shape},
        "attention": attention,
    }
    if host_affinity:
        line["host_affinity"] = host_affinity
    if exchange is not None:
        line["allreduce_exposed_ms"] = allreduce_exposed_ms
        line["exchange"] = exchange.describe()
    if world == 1 and not args.no_extras:
        line["extras"] = run_extras(args, dev, model, grads, opt, graphed, resident, resident_adj, B)
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cstep, kind, desc = cpu_reference_step_fn(args.ref_batch, 0, threads)
        cstep()
        t0 = time.time()
        n = 2
        for _ in range(n):
            cstep()
        dt = (time.time() - t0) / n
        line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": UNIT, "cores": threads, "kind": kind,
                                "sample": desc + "; 1 warm-up + %d timed steps" % n}
    print(json.dumps(line), flush=True)
    if world > 1:
        del graphed
        torch.cuda.synchronize()
        dist.destroy_process_group()


def run_extras(args, dev, model, grads, opt, graphed, resident, resident_adj, B):
    """Secondary measurements on the same box (N = 1): short, each reported with its own timing."""
    import numpy as np
    import torch
    from sam_textvqa_b200 import dp, ops, spatial_utils, synth
    from sam_textvqa_b200.config import c3_config
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig
    out = {}

    def timeit(fn, steps, warm=2):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / steps

    def fwd_bwd(mdl, gr, res, adjd):
        gr.zero()
        bd = dict(res)
        bd["spatial_adj_matrices"] = adjd
        ops.bce_with_mask_loss(mdl(bd)["textvqa_scores"], res["targets"], res["train_loss_mask"]).backward()

    try:
        # 1. the same step in the other precision mode (strict: 3-term splits, exact fp32 attention), eager
        cur = ops.get_precision()
        other = "bf16x3" if cur == "f16" else "f16"
        ops.set_precision(other)
        ops.clear_weight_cache()
        ms_o = timeit(lambda: fwd_bwd(model, grads, resident, {"3": resident_adj}), 3, 1)
        out["other_precision_mode"] = {"mode": other, "ms_per_step": ms_o, "samples_per_s": B / ms_o * 1e3, "launch": "eager"}
        ops.set_precision(cur)
        ops.clear_weight_cache()
        # 2. full training step: graph-replayed fwd+bwd + fused clip + Adam on the flat buffers (train.py:139-143)
        if graphed is not None:
            try:
                def train_step():
                    graphed.run()
                    opt.step()
                ms_t = timeit(train_step, 5, 2)
                out["train_step_with_clip_adam"] = {"ms_per_step": ms_t, "samples_per_s": B / ms_t * 1e3}
            except Exception as exc:
                out["train_step_with_clip_adam"] = {"error": repr(exc)[:200]}
        # 3. greedy 12-step decoding (eval), KV-cached
        model.eval()

        def decode():
            bd = dict(resident)
            bd["spatial_adj_matrices"] = {"3": resident_adj}
            with torch.no_grad():
                model(bd)
        ms_d = timeit(decode, 3, 1)
        out["greedy_decode_12_steps"] = {"ms_per_batch": ms_d, "samples_per_s": B / ms_d * 1e3}
        model.train()
        # 4. spatial-graph kernel (BASELINE: the reference builds 19-56 k pairs/s per core in Python)
        N = CFG["O"] + CFG["R"]
        boxes = torch.from_numpy(synth.make_boxes(np.random.RandomState(0), 1024, N)[..., :4].astype("float32")).to(dev)
        ms_g = timeit(lambda: spatial_utils.build_graph_batch(boxes, 0.5, context=3), 10, 2)
        pairs = 1024 * N * N
        gbytes = 1024 * (16 * N + N * N + 2 * N * N)            # float4 boxes in, int8 types + uint16 head bits out
        out["graph_kernel"] = {"B": 1024, "N": N, "us_per_launch": ms_g * 1e3, "pairs_per_s": pairs / (ms_g * 1e-3),
                               "algorithmic_GBs": gbytes / (ms_g * 1e-3) / 1e9,
                               "bound": "latency / fp64 classification (24.9 KB in+out per sample)"}
        # 5. BASELINE configs 0 and 2 (fwd+bwd, eager)
        graph_fn = lambda b: spatial_utils.build_graph_batch(b.astype("float32"), 0.5)[0]
        for name, over, Bc, O, R, ctx in (("cfg0: 1 spatial layer, 20+36+50 tokens, B=4", dict(layer_type_list=["s"], mix_list=["share3"]), 4, 36, 50, 3),
                                          ("cfg2: c5, 100 obj + 100 OCR tokens, B=256",
                                           dict(mix_list=["none", "none", "share5", "share5", "share5", "share5"]), 256, 100, 100, 5)):
            mmt, tb = c3_config(**over)
            torch.manual_seed(0)
            m2 = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb)).to(dev).train()
            g2 = dp.FlatGradBuffer(m2.parameters())
            b2 = synth.make_batch(Bc, O=O, R=R, V=CFG["V"], seed=0, contexts=(ctx,), graph_fn=graph_fn)
            adjd = {str(ctx): b2.pop("spatial_adj_matrices")[str(ctx)].to(dev)}
            if ctx != 1:
                adjd["1"] = adjd[str(ctx)]
            r2 = {k: v.to(dev) for k, v in b2.items() if torch.is_tensor(v)}
            ms_c = timeit(lambda: fwd_bwd(m2, g2, r2, adjd), 3 if Bc > 8 else 10, 2)
            out[name] = {"ms_per_step": ms_c, "samples_per_s": Bc / ms_c * 1e3, "launch": "eager"}
            del m2, g2, r2, adjd
            torch.cuda.empty_cache()
    except Exception as exc:      # extras never take the headline down
        out["error"] = repr(exc)[:300]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="samk", choices=["samk", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="samples per GPU")
    ap.add_argument("--ref-batch", type=int, default=32, help="bounded CPU sample: rows of the batch per reference step")
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="run 3 bare steps and exit (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_samk(args)


if __name__ == "__main__":
    main()
