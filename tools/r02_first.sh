#!/bin/bash
# round-2 first GPU check: mixed-format GEMM, attention, model parity, short bench
set -u
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r02a}
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q > $OUT/${TAG}_pytest_ops.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest_ops.log; tail -15 $OUT/${TAG}_pytest_ops.log
timeout 600 python -m pytest tests/test_gpu_model.py tests/test_gpu_graph.py -m gpu -x -q -s > $OUT/${TAG}_pytest_model.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest_model.log; tail -25 $OUT/${TAG}_pytest_model.log
timeout 300 python tools/attn_bench.py > $OUT/${TAG}_attn_bench.log 2>&1; cat $OUT/${TAG}_attn_bench.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; tail -5 $OUT/${TAG}_bench.err; python - <<PY
import json
d=json.load(open("$OUT/${TAG}_bench.json"))
print({k:d[k] for k in ("value","ms_per_step","gpu_launches")}, d["e2e"]["value"], d["attention"].get("fwd"), d["attention"].get("bwd"), d["roofline"]["achieved"])
PY
