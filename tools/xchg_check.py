#!/usr/bin/env python
"""samk_exchange_sum (csrc/exchange.cu) against ncclAllReduce on real GPUs (torchrun, one process per GPU):
equality of the sums (eager, back to back with changing data, and replayed from a CUDA graph), both data paths
(NVSwitch multicast / peer loads and stores), both wire formats, and the time of one exchange of the shipped model's
gradient buffer next to the library collective.  Prints one JSON line on rank 0.  TEST / EVIDENCE TOOL."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from sam_textvqa_b200 import dp
    out = {"world": world, "cases": []}
    n = int(os.environ.get("XCHG_N", str(96_600_000 // 64 * 64)))
    gen = torch.Generator(device=dev).manual_seed(100 + rank)

    def reference(x, wire_dtype):
        y = x.to(wire_dtype).float().clone()
        dist.all_reduce(y)
        return y.to(wire_dtype).float()

    for wire_dtype in (torch.bfloat16, torch.float32):
        for mc in (True, False):
            pw = dp.PeerWire(n, wire_dtype, use_multicast=mc)
            if mc and not pw.multicast:
                out["cases"].append({"wire": str(wire_dtype), "multicast": "unavailable"})
                continue
            storage = torch.zeros(pw.n, device=dev)
            vec = pw.vec
            ulp = 2.0 ** -7 if wire_dtype == torch.bfloat16 else 2.0 ** -19     # (fp32: the order of the N addends differs)
            worst = torch.zeros(1, device=dev)
            ranges = [(0, pw.n), (vec * 3, vec * 3 + 1024 * vec), (pw.n // 2 // vec * vec, pw.n), (vec * 5, vec * 6)]
            for it in range(6):                                   # back to back, new data each time, no host sync between
                lo, hi = ranges[it % len(ranges)]
                x = torch.randn(pw.n, device=dev, generator=gen) * (1.0 + it)
                want = reference(x[lo:hi], wire_dtype)
                storage.copy_(x)
                pw.exchange_sum(storage, lo, hi)
                err = ((storage[lo:hi] - want).abs() / (want.abs() + 1.0)).max()
                worst = torch.maximum(worst, err.reshape(1))
                if lo > 0:                                        # outside the range nothing moves
                    worst = torch.maximum(worst, (storage[:lo] - x[:lo]).abs().max().reshape(1))
            # replayed from a CUDA graph, on a side stream beside a compute kernel of the capture stream
            x = torch.randn(pw.n, device=dev, generator=gen)
            a = torch.randn(4096, 4096, device=dev)
            b = a @ a                                             # (library handle created outside the capture)
            want = reference(x, wire_dtype)
            side = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                storage.copy_(x)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):
                    pw.exchange_sum(storage, 0, pw.n)
                b = a @ a
                torch.cuda.current_stream().wait_stream(side)
            for _ in range(3):
                g.replay()
            gerr = ((storage - want).abs() / (want.abs() + 1.0)).max()
            # timing of one whole-buffer exchange next to the library collective on the same bytes
            def timeit(fn, reps=10):
                for _ in range(2):
                    fn()
                dist.barrier(); torch.cuda.synchronize()
                s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                s.record()
                for _ in range(reps):
                    fn()
                e.record(); torch.cuda.synchronize()
                t = torch.tensor([s.elapsed_time(e) / reps], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                return float(t)
            wire_nccl = torch.zeros(pw.n, dtype=wire_dtype, device=dev)
            t_peer = timeit(lambda: pw.exchange_sum(storage, 0, pw.n))
            t_nccl = timeit(lambda: dist.all_reduce(wire_nccl))
            torch.cuda.synchronize()
            pw.check()
            out["cases"].append({"wire": str(wire_dtype).replace("torch.", ""), "multicast": pw.multicast, "n": pw.n,
                                 "max_rel_err": float(worst), "graph_rel_err": float(gerr), "tol": 2 * ulp,
                                 "ms_exchange_sum": t_peer, "ms_nccl_allreduce_same_bytes": t_nccl})
            assert float(worst) <= 2 * ulp and float(gerr) <= 2 * ulp, out["cases"][-1]
            del pw
    dist.barrier()
    if rank == 0:
        print(json.dumps(out), flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
