cd /root/repo
OUT=gpurun_out; TAG=r02e
timeout 600 ncu --clock-control none --metrics gpu__time_duration.sum -c 6000 --csv --log-file $OUT/${TAG}_launches.csv python bench.py --profile-only > $OUT/${TAG}_launches.out 2>&1
python tools/summarize_launches.py $OUT/${TAG}_launches.csv 1 > $OUT/${TAG}_launches_summary.txt 2>&1
head -60 $OUT/${TAG}_launches_summary.txt; gzip -f $OUT/${TAG}_launches.csv
