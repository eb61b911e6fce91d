#!/bin/bash
# round-2 GPU check: full GPU test suite, bench line (both arms), attention bench
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02a}
python -c "import __graft_entry__ as g; g.build()" || exit 1
nproc > $OUT/${TAG}_nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -s > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest_gpu.log; tail -12 $OUT/${TAG}_pytest_gpu.log
timeout 300 python tools/attn_bench.py > $OUT/${TAG}_attn_bench.log 2>&1; cat $OUT/${TAG}_attn_bench.log
timeout 600 python bench.py --steps 20 --warmup 5 ${BENCH_ARGS:-} > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -5 $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, "e2e", d["e2e"]["value"])
print("parity", d["parity"])
print("attn", {k:(round(v["us_per_launch"],1), round(v["hbm_frac"],3)) for k,v in d["attention"].items() if isinstance(v,dict)})
print("gemm", round(d["roofline"]["achieved"],1), round(d["roofline"]["frac"],3))
print("extras", json.dumps(d.get("extras"), indent=0)[:1500])
print("cpu", d.get("cpu_baseline"))
PY
if [ "${REF_ARM:-0}" = "1" ]; then
timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err; cat $OUT/${TAG}_bench_reference.json | cut -c1-600
fi
