#!/bin/bash
# weak-scaling line at N GPUs (bench.py under torchrun), default exchange settings
set -u
N=${1:-8}; TAG=${2:-r02s}
OUT=gpurun_out; mkdir -p $OUT
python -c "import __graft_entry__ as g; g.build()" || exit 1
for cfg in ${CFGS:-default}; do
  case $cfg in
    default) ENVV="A=1";;
    f32wire) ENVV="SAMK_DP_WIRE=f32";;
    *) ENVV="$cfg";;
  esac
  env $ENVV timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --steps 20 --warmup 5 > $OUT/${TAG}_bench_n${N}_${cfg}.json 2> $OUT/${TAG}_bench_n${N}_${cfg}.err
  grep -v "OMP_NUM\|\*\*\*" $OUT/${TAG}_bench_n${N}_${cfg}.err | tail -3 | cut -c1-300
  python - <<PY
import json
try:
    d=json.loads([l for l in open("$OUT/${TAG}_bench_n${N}_${cfg}.json") if l.startswith("{")][-1])
    print("$cfg", "N=", d["n_gpus"], "value", round(d["value"]), "ms", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"]), "exposed_ms", d.get("allreduce_exposed_ms"), d.get("exchange",{}).get("wire_dtype"))
except Exception as e: print("no line", e)
PY
done
