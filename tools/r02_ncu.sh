#!/bin/bash
# ncu evidence: per-launch DRAM bytes / tensor pipe of one eager step (-> profiles/ncu_traffic.json) and full-set
# captures of the attention kernels at L = 182 (bench geometry), 524 and 1036 (BASELINE config 4 sweep)
set -u
TAG=${1:-r02n}; WHAT=${2:-"traffic attn"}
OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --clock-control none"
python -c "import __graft_entry__ as g; g.build()" || exit 1
for w in $WHAT; do
case $w in
traffic)
  SAMK_BENCH_EAGER=1 SAMK_LAUNCH_LOG=$OUT/${TAG}_launch_shapes.json timeout 900 $NCU \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
    -c 1300 --csv --log-file $OUT/${TAG}_step_metrics.csv python bench.py --profile-only > $OUT/${TAG}_step_metrics.out 2>&1
  python tools/ncu_summary.py --traffic $OUT/${TAG}_step_metrics.csv $OUT/${TAG}_launch_shapes.json $OUT/${TAG}_ncu_traffic.json "profiles/${TAG}_step_metrics.csv.gz (ncu --metrics, eager step, cold caches)"
  gzip -f $OUT/${TAG}_step_metrics.csv ;;
attn)
  FULL="$NCU --set full --import-source on"
  timeout 400 $FULL -k regex:attn_fwd3_kernel -s 12 -c 2 -f -o $OUT/${TAG}_ncu_attn_fwd_L182 python bench.py --profile-only > $OUT/${TAG}_ncu_attn_fwd.out 2>&1
  timeout 400 $FULL -k regex:attn_bwd2_kernel -s 9 -c 2 -f -o $OUT/${TAG}_ncu_attn_bwd_L182 python bench.py --profile-only > $OUT/${TAG}_ncu_attn_bwd.out 2>&1
  timeout 400 $FULL -k regex:"attn_fwd2_kernel|attn_fwd3_kernel" -s 8 -c 1 -f -o $OUT/${TAG}_ncu_attn_fwd_L524 python tools/attn_bench.py --only 442 --reps 1 > $OUT/${TAG}_ncu_attn_fwd_L524.out 2>&1
  timeout 400 $FULL -k regex:attn_bwd_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_ncu_attn_bwd_L524 python tools/attn_bench.py --only 442 --reps 1 > $OUT/${TAG}_ncu_attn_bwd_L524.out 2>&1
  timeout 400 $FULL -k regex:"attn_fwd2_kernel|attn_fwd3_kernel" -s 4 -c 1 -f -o $OUT/${TAG}_ncu_attn_fwd_L1036 python tools/attn_bench.py --only 954 --reps 1 > $OUT/${TAG}_ncu_attn_fwd_L1036.out 2>&1
  timeout 400 $FULL -k regex:attn_bwd_tc_kernel -s 3 -c 1 -f -o $OUT/${TAG}_ncu_attn_bwd_L1036 python tools/attn_bench.py --only 954 --reps 1 > $OUT/${TAG}_ncu_attn_bwd_L1036.out 2>&1
  for f in $OUT/${TAG}_ncu_*.ncu-rep; do
    ncu -i $f --page raw --csv > ${f%.ncu-rep}_raw.csv 2>/dev/null
    ncu -i $f --page source --csv 2>/dev/null | gzip > ${f%.ncu-rep}_source.csv.gz
    rm -f $f
  done
  python tools/ncu_summary.py $OUT/${TAG}_ncu_*_raw.csv > $OUT/${TAG}_ncu_summary.txt 2>&1
  grep -E "^--|duration|dram read|dram write|tensor pipe|issue slots|occupancy|stall (long_sb|barrier|short|wait|math|not_sel)" $OUT/${TAG}_ncu_summary.txt | head -120 ;;
esac
done
ls -la $OUT | tail -20
