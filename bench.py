#!/usr/bin/env python
"""SA-M4C fwd+bwd samples/s on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl samk|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one pass of the hot path over one synthetic batch: SAM4C forward in train mode (all
dropouts at the yml's 0.1) -> masked BCE loss -> backward (-> one NCCL gradient all-reduce when
N > 1) -> zero_grad.  Workload at every N: BASELINE config[1], the shipped c3 experiment
(layers n,n,s,s,s,s; 20 text + 100 obj + 50 OCR + 12 dec tokens; d=768; V=5000), 128 samples per GPU
(weak scaling).  `value` is timed with the inputs resident in HBM; `e2e` runs the same step from
pinned host buffers with the host->device copies and a device->host read of the loss inside the
timed region.  `--impl reference` times the CPU restatement of the reference algorithm
(oracle/sam4c_oracle.py, the reference's own torch CPU ops and mask algebra; the reference tree
itself does not exist on the GPU box) on the host cores, on a bounded sample of the same workload.
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
def cpu_port_step_fn(B, seed, threads):
    """Returns (step, description): one fwd+loss+bwd+zero_grad of the oracle port on `B` samples."""
    import numpy as np
    import torch
    from oracle import graph_oracle, sam4c_oracle
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

    return step, "oracle port (torch CPU fp32, reference op sequence incl. dense masks, unique-check, dropout), B=%d" % B


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    B = args.ref_batch
    step, desc = cpu_port_step_fn(B, 0, threads)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = (time.time() - t0) / args.steps
    val = B / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": 1, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample_batch": B},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
def run_samk(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from sam_textvqa_b200 import build as samk_build
    samk_build.build()
    from sam_textvqa_b200 import dp, ops, spatial_utils, synth
    from sam_textvqa_b200.config import c3_config
    from sam_textvqa_b200.registry import registry
    from sam_textvqa_b200.sa_m4c import SAM4C, BertConfig

    ops.set_precision(args.precision)
    registry.answer_vocab = ["w%d" % i for i in range(CFG["V"])]
    registry.BOS_IDX = 1
    mmt, tb = c3_config()
    torch.manual_seed(0)
    model = SAM4C(BertConfig.from_dict(mmt), BertConfig.from_dict(tb)).to(dev).train()
    grads = dp.FlatGradBuffer(model.parameters())
    # SAMK_DP_OVERLAP=1: bucketed NCCL all-reduce on a side stream under the backward pass.  Measured at N=2 it does
    # not pay (13.97 vs 13.85 ms/step): NCCL's CTAs displace CTAs of the persistent 148-CTA GEMMs, which then run a
    # second partial wave.  Default: one all-reduce of the flat buffer after backward (~0.9 ms exposed).
    overlap = world > 1 and os.environ.get("SAMK_DP_OVERLAP", "0") == "1"
    if overlap:
        # one early bucket: everything that is final when the MMT stack's backward ends (~half of the bytes) goes
        # out while the TextBERT backward runs (its GEMMs have fewer tiles than SMs, so NCCL's CTAs cost nothing)
        grads.enable_overlap(bucket_bytes=int(os.environ.get("SAMK_DP_BUCKET_MB", "160")) << 20)
    B = args.batch

    def graph_fn(boxes):
        return spatial_utils.build_graph_batch(boxes.astype("float32"), 0.5)[0]

    host = synth.make_batch(B, seed=rank, contexts=(3,), graph_fn=graph_fn, **CFG)
    host.pop("boxes"); host.pop("spatial_types")
    adj = host.pop("spatial_adj_matrices")["3"]
    names = sorted(k for k, v in host.items() if torch.is_tensor(v))
    pinned = {k: host[k].pin_memory() for k in names}
    pinned_adj = adj.pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in pinned.values()) + pinned_adj.numel()
    resident = {k: v.to(dev) for k, v in pinned.items()}
    resident_adj = pinned_adj.to(dev)

    def step(inputs, adj_dev):
        grads.zero()
        if overlap:
            grads.begin_step()
        bd = dict(inputs)
        bd["spatial_adj_matrices"] = {"3": adj_dev}
        scores = model(bd)["textvqa_scores"]
        loss = ops.bce_with_mask_loss(scores, inputs["targets"], inputs["train_loss_mask"])
        loss.backward()
        if overlap:
            grads.finish_step()
        elif world > 1:
            grads.all_reduce()
        return loss

    def upload():
        up = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        return up, pinned_adj.to(dev, non_blocking=True)

    # The step as a user runs it: captured once into a CUDA graph (sam_textvqa_b200/graph_step.py) and replayed --
    # ~310 launches through ctypes + autograd cost the host about as long as the GPU needs for the step.
    # SAMK_BENCH_EAGER=1 keeps the eager loop.
    graphed = None
    if os.environ.get("SAMK_BENCH_EAGER", "0") != "1":
        try:
            from sam_textvqa_b200.graph_step import GraphedTrainStep
            ex = dict(resident)
            ex["spatial_adj_matrices"] = {"3": resident_adj}
            graphed = GraphedTrainStep(model, grads, ex, allreduce="overlap" if overlap else None)
        except Exception as exc:                      # fall back loudly, never silently
            print("bench: CUDA-graph capture failed (%r); running the eager step" % (exc,), file=sys.stderr, flush=True)
            graphed = None
    eager_step = step

    def step(inputs, adj_dev):                       # noqa: F811  (same contract as the eager step above)
        if graphed is None:
            return eager_step(inputs, adj_dev)
        if inputs is not resident:                    # fresh upload: device-to-device copy into the graph's input buffers
            graphed.load(inputs)
            graphed.load({"spatial_adj_matrices": {"3": adj_dev}})
        loss = graphed.run()
        if world > 1 and not overlap:
            grads.all_reduce()
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
        for _ in range(3):
            step(resident, resident_adj)
        torch.cuda.synchronize()
        return
    for _ in range(max(args.warmup, 3)):
        step(resident, resident_adj)
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.launch_count
    ms = timed(lambda: step(resident, resident_adj), args.steps)
    launches = ops.launch_count - l0
    if graphed is not None:                    # replays do not pass through the Python counters
        launches = graphed.kernels_per_replay * args.steps

    # ---- end to end: every step's inputs come from pinned host memory; the copy of step i+1 runs on a
    # side stream while step i computes (double-buffered device staging), and every step's loss is read
    # back to the host.  All K uploads and K loss reads are inside the timed region.
    copy_stream = torch.cuda.Stream()
    # two persistent device staging sets; `consumed[slot]` (compute stream) says the step that read slot has taken its
    # inputs, `ready[slot]` (copy stream) says the next upload into slot has landed
    staged = [({k: torch.empty_like(v) for k, v in resident.items()}, torch.empty_like(resident_adj)) for _ in range(2)]
    ready = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    for ev in consumed:
        ev.record()

    def stage(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            bufs, abuf = staged[slot]
            for k in names:
                bufs[k].copy_(pinned[k], non_blocking=True)
            abuf.copy_(pinned_adj, non_blocking=True)
            ready[slot].record(copy_stream)

    diag = os.environ.get("SAMK_E2E_DIAG", "")
    host_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_done = [torch.cuda.Event(), torch.cuda.Event()]
    loss_values = []

    def e2e_run(steps):
        if diag == "noupload":
            for i in range(steps):
                step(resident, resident_adj).item()
            return
        stage(0)
        for i in range(steps):
            torch.cuda.current_stream().wait_event(ready[i % 2])
            if i + 1 < steps:
                stage((i + 1) % 2)
            up, a = staged[i % 2]
            if graphed is not None:                   # inputs are taken by the device-to-device load in front of the replay
                graphed.load(up)
                graphed.load({"spatial_adj_matrices": {"3": a}})
                consumed[i % 2].record()
                loss = graphed.run()
                if world > 1 and not overlap:
                    grads.all_reduce()
            else:
                loss = eager_step(up, a)
                consumed[i % 2].record()
            # device->host read of EVERY step's loss: an async copy into pinned memory plus an event right behind the
            # step; the host reads step i-1's value after it has enqueued step i, waiting on that event only
            # (`.item()` would synchronise the whole stream, i.e. also wait for the step just enqueued, and the GPU
            # would then idle while the host prepares the next one)
            if diag != "noitem":
                host_loss[i % 2].copy_(loss.detach().reshape(1), non_blocking=True)
                loss_done[i % 2].record()
                if i >= 1:
                    loss_done[(i - 1) % 2].synchronize()
                    loss_values.append(float(host_loss[(i - 1) % 2][0]))
        if diag != "noitem" and steps >= 1:
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

    # ---- roofline of the dominant kernel family (tcgen05 GEMM), instrumented extra steps ----
    ops.gemm_profile = []
    ops.attn_profile = []
    for _ in range(2):
        # per-launch events cannot be recorded inside a graph replay, and the eager loop is launch-bound: an event pair
        # around a kernel the GPU is waiting for also times the host's launch latency.  A spin kernel in front keeps the
        # GPU busy while the host enqueues the whole step, so the events bracket back-to-back kernel executions.
        try:
            torch.cuda._sleep(int(4e7))        # ~20 ms at 1.9 GHz; the host needs ~10-15 ms to enqueue a step
        except Exception:                      # private torch API: without it the figures are only more pessimistic
            pass
        eager_step(resident, resident_adj)
        torch.cuda.synchronize()
    prof, ops.gemm_profile = ops.gemm_profile, None
    aprof, ops.attn_profile = ops.attn_profile, None
    g_ms = sum(s.elapsed_time(e) for s, e, _, _ in prof) / 2
    g_flop = sum(f for _, _, f, _ in prof) / 2
    n_gemm = len(prof) // 2
    peak_tf, peak_gbs, peak_src = peaks()
    achieved = g_flop / (g_ms * 1e-3) / 1e12 if g_ms > 0 else 0.0
    # the same live events per GEMM shape: the five shapes that take the most time in a step, each with its algorithmic
    # operand bytes (bf16 A and B read once, output written once) and, where profiles/ holds an `ncu --set full`
    # capture of that shape, the measured DRAM bytes per launch
    ncu_dram = {   # (M, N, K, a_mn, b_mn) -> (read MB, write MB, file)  -- profiles/r01j_ncu_summary.txt
        (768, 3072, 23296, True, True): (342.2, 10.0, "r01j_ncu_gemm_ffn2_wgrad_raw.csv"),
        (23296, 3072, 768, False, False): (40.6, 234.0, "r01j_ncu_gemm_ffn1_fwd_raw.csv"),
        (23296, 3072, 768, False, True): (183.7, 110.0, "r01j_ncu_gemm_ffn2_dgrad_raw.csv"),
        (23296, 2304, 768, False, False): (39.4, 54.9, "r01j_ncu_gemm_qkv_fwd_raw.csv"),
    }
    groups = {}
    for s_, e_, f_, shape in prof:
        g = groups.setdefault(shape, [0.0, 0, f_])
        g[0] += s_.elapsed_time(e_)
        g[1] += 1
    by_shape = []
    for shape, (tot_ms, cnt, f_) in sorted(groups.items(), key=lambda kv: -kv[1][0])[:5]:
        us = tot_ms / cnt * 1e3
        tf = f_ / (us * 1e-6) / 1e12
        Mg, Ng, Kg, amn, bmn = shape
        item = {"M": Mg, "N": Ng, "K": Kg, "a_mn": amn, "b_mn": bmn, "launches_per_step": cnt // 2,
                "us_per_launch": us, "achieved": tf, "frac": tf / peak_tf, "share_of_step": tot_ms / 2 / ms if ms > 0 else None,
                "operand_MB": (Mg * Kg + Ng * Kg) * 2 / 1e6}
        if shape in ncu_dram and B == 128:
            rd, wr, src = ncu_dram[shape]
            item["traffic"] = (rd + wr) * 1e6
            item["traffic_src"] = "profiles/" + src
        by_shape.append(item)
    # the north-star kernel: fused masked attention of the MMT layers (L = 182), HBM-bound at this length
    Lm = CFG["T"] + CFG["O"] + CFG["R"] + CFG["D"]
    attention = {}
    for kind in ("fwd", "bwd"):
        rows = [(s.elapsed_time(e), nb, fl) for k, L_, s, e, nb, fl in aprof if k == kind and L_ == Lm]
        if rows:
            ms_k = sum(r[0] for r in rows) / len(rows)
            attention[kind] = {"us_per_launch": ms_k * 1e3, "launches_per_step": len(rows) // 2,
                               "algorithmic_MB": rows[0][1] / 1e6, "achieved_GBs": rows[0][1] / (ms_k * 1e-3) / 1e9,
                               "hbm_frac": rows[0][1] / (ms_k * 1e-3) / 1e9 / peak_gbs,
                               "dense_equiv_TFLOPs": rows[0][2] / (ms_k * 1e-3) / 1e12}
    attention["peak_GBs"] = peak_gbs
    attention["ncu"] = ("profiles/r01j_ncu_summary.txt: fwd (attn_fwd3) 80.6 us, DRAM 185+29 MB per launch, tensor pipe "
                        "12.1 %, issue slots 50 %; bwd (attn_bwd2) 190 us, DRAM 153+73 MB, tensor pipe 14.4 % "
                        "(L=182 is HBM/ALU-bound, SURVEY 8d)")

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = world * B / (ms * 1e-3)
    e2e_val = world * B / (ms_e2e * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f16" if args.precision == "f16" else "bf16x3", "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "batch_per_gpu": B, "global_batch": world * B, "parallelism": "dp%d" % world,
                   "launch": "cuda-graph replay of the captured step" if graphed is not None else "eager",
                   "l2": "per-step working set (~6 GB of activations) >> 126 MB L2; no explicit flush"},
        "clocks": clocks,
        "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": int(h2d_bytes),
                "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "samk::gemm_tc_kernel (all %d GEMM launches of one step)" % n_gemm,
                     "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved / peak_tf,
                     "traffic": None, "peak_source": peak_src + " (sustained bf16)",
                     "gemm_ms_per_step": g_ms, "gemm_share_of_step": g_ms / ms if ms > 0 else None,
                     "step_flop_frac_of_peak": value / world * FLOP_PER_SAMPLE / (peak_tf * 1e12),
                     "by_shape": by_shape,
                     "traffic_note": "aggregate over all GEMM shapes of the step, so no single per-launch figure; per-shape "
                                     "DRAM bytes from ncu --set full are in profiles/r01i_ncu_summary.txt "
                                     "(e.g. FFN2 wgrad 342+9 MB vs 322 MB algorithmic)"},
        "attention": attention,
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cstep, desc = cpu_port_step_fn(args.ref_batch, 0, threads)
        cstep()
        t0 = time.time()
        n = 2
        for _ in range(n):
            cstep()
        dt = (time.time() - t0) / n
        line["cpu_baseline"] = {"value": args.ref_batch / dt, "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": desc + ", 1 warm-up + %d timed iterations" % n}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="samk", choices=["samk", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="samples per GPU")
    ap.add_argument("--ref-batch", type=int, default=16, help="bounded CPU sample size")
    ap.add_argument("--precision", default="f16", choices=["f16", "bf16x3"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-only", action="store_true", help="run 3 bare steps and exit (for ncu launch lists)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_samk(args)


if __name__ == "__main__":
    main()
