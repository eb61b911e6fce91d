#!/bin/bash
# One gpurun call that collects the evidence set for a round: GPU parity tests, the bench line (both arms),
# the ncu launch list of the bench step, `ncu --set full` captures of the dominant kernels, and the
# secondary configs.  Usage (from the repo root, on the GPU box): bash tools/gpu_evidence.sh <tag> [sections]
set -u
TAG=${1:-r01x}
WHAT=${2:-"tests bench launches ncu attn extra"}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
python -c "import __graft_entry__ as g; g.build()" || exit 1
nproc > $OUT/${TAG}_nproc.txt
for w in $WHAT; do
case $w in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest_gpu.log
  tail -3 $OUT/${TAG}_pytest_gpu.log ;;
bench)
  timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench.json
  timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
  cat $OUT/${TAG}_bench_reference.json ;;
launches)
  timeout 600 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --profile-only > $OUT/${TAG}_launches.out 2>&1
  python tools/summarize_launches.py $OUT/${TAG}_launches.csv 1 > $OUT/${TAG}_launches_summary.txt 2>&1
  head -30 $OUT/${TAG}_launches_summary.txt; gzip -f $OUT/${TAG}_launches.csv ;;
ncu)
  FULL="$NCU --set full --import-source on"
  for shape in ffn1_fwd ffn2_wgrad ffn2_dgrad qkv_fwd; do
    REPS=1 timeout 300 $FULL -k regex:gemm_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_ncu_gemm_$shape \
        python tools/gemm_bench.py $shape > $OUT/${TAG}_ncu_gemm_$shape.out 2>&1
  done
  timeout 400 $FULL -k regex:attn_fwd3_kernel -s 12 -c 3 -f -o $OUT/${TAG}_ncu_attn_fwd \
      python bench.py --profile-only > $OUT/${TAG}_ncu_attn_fwd.out 2>&1
  timeout 400 $FULL -k regex:attn_bwd2_kernel -s 9 -c 2 -f -o $OUT/${TAG}_ncu_attn_bwd \
      python bench.py --profile-only > $OUT/${TAG}_ncu_attn_bwd.out 2>&1
  timeout 400 $FULL -k regex:"layernorm_bwd_kernel|colsum_kernel|layernorm_fwd_kernel" -s 90 -c 6 -f -o $OUT/${TAG}_ncu_ln \
      python bench.py --profile-only > $OUT/${TAG}_ncu_ln.out 2>&1
  for f in $OUT/${TAG}_ncu_*.ncu-rep; do
    ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
    ncu -i $f --page source --csv 2>/dev/null | gzip > ${f%.ncu-rep}_source.csv.gz
    ncu -i $f --page details 2>/dev/null | gzip > ${f%.ncu-rep}_details.txt.gz
    rm -f $f
  done
  ls -la $OUT/ ;;
attn)
  timeout 300 python tools/attn_bench.py > $OUT/${TAG}_attn_bench.log 2>&1
  timeout 300 python tools/attn_bench.py --sweep >> $OUT/${TAG}_attn_bench.log 2>&1
  timeout 300 python tools/attn_bench.py --B 256 --context 5 >> $OUT/${TAG}_attn_bench.log 2>&1
  cat $OUT/${TAG}_attn_bench.log
  timeout 300 python tools/gemm_bench.py > $OUT/${TAG}_gemm_bench.log 2>&1; cat $OUT/${TAG}_gemm_bench.log ;;
extra)
  timeout 600 python tools/extra_bench.py > $OUT/${TAG}_extra_bench.log 2>&1; cat $OUT/${TAG}_extra_bench.log ;;
esac
done
