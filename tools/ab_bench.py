#!/usr/bin/env python
"""Interleaved A/B runs of bench.py on ONE box: box-to-box and within-box drift of the step time is about +-1 %, as
large as most single-kernel changes, so each variant is run `--rounds` times in ABAB order and the medians are compared.

  python tools/ab_bench.py --rounds 3 base: pdl1:SAMK_PDL=1 "side0:SAMK_SIDE_BRANCH=0"

Each variant is `label:ENV=VALUE,ENV2=VALUE2` (empty after the colon = default environment).  Prints one line per run and
a summary table (median ms/step, spread, ratio to the first variant).  Run it under gpurun; keep --steps small.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_once(env_pairs, steps, warmup, extra):
    env = dict(os.environ)
    env.update(env_pairs)
    cmd = [sys.executable, os.path.join(ROOT, "bench.py"), "--steps", str(steps), "--warmup", str(warmup),
           "--no-cpu-baseline"] + extra
    out = subprocess.run(cmd, env=env, cwd=ROOT, capture_output=True, text=True, timeout=600)
    for line in reversed(out.stdout.splitlines()):
        if line.startswith("{"):
            return json.loads(line)
    raise RuntimeError("bench.py printed no JSON line:\n" + out.stderr[-2000:])


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("variants", nargs="+")
    ap.add_argument("--rounds", type=int, default=3)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--bench-args", default="", help="extra arguments passed to bench.py")
    args = ap.parse_args()
    variants = []
    for v in args.variants:
        label, _, envs = v.partition(":")
        pairs = dict(kv.split("=", 1) for kv in envs.split(",") if kv)
        variants.append((label, pairs))
    results = {label: [] for label, _ in variants}
    for r in range(args.rounds):
        for label, pairs in variants:
            d = run_once(pairs, args.steps, args.warmup, args.bench_args.split())
            results[label].append((d["ms_per_step"], d["e2e"]["ms_per_step"]))
            print("round %d %-16s %.3f ms/step  e2e %.3f ms  %s" % (r, label, d["ms_per_step"], d["e2e"]["ms_per_step"],
                                                                  d["clocks"]["reasons"] if d.get("clocks") else ""), flush=True)
    base = statistics.median(x[0] for x in results[variants[0][0]])
    print("\n%-16s %10s %10s %10s %8s" % ("variant", "median ms", "min", "max", "vs first"))
    for label, _ in variants:
        ms = [x[0] for x in results[label]]
        print("%-16s %10.3f %10.3f %10.3f %8.4f" % (label, statistics.median(ms), min(ms), max(ms), statistics.median(ms) / base))


if __name__ == "__main__":
    main()
