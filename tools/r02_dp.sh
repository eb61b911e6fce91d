#!/bin/bash
# N-GPU check: NCCL equivalence test + scaling bench line
set -u
N=${1:-2}; TAG=${2:-r02dp}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" || exit 1
timeout 600 python -m pytest tests/test_gpu_dp.py -m gpu -x -q -s > $OUT/${TAG}_pytest_dp.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest_dp.log; tail -5 $OUT/${TAG}_pytest_dp.log
for cfg in ${CFGS:-default}; do
  case $cfg in
    default) ENVV="";;
    bf16wire) ENVV="SAMK_DP_WIRE=bf16";;
    overlap) ENVV="SAMK_DP_OVERLAP=1";;
    *) ENVV="$cfg";;
  esac
  env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n${N}_${cfg}.json 2> $OUT/${TAG}_bench_n${N}_${cfg}.err
  tail -3 $OUT/${TAG}_bench_n${N}_${cfg}.err | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_${cfg}.json") if l.startswith("{")][-1])
    print("$cfg", "N=", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "exposed_ms", d.get("allreduce_exposed_ms"), d.get("exchange"))
except Exception as e: print("no line", e)
PY
done
