cd /root/repo
SAMK_DEBUG_CAPTURE=1 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | grep -v "^{" | tail -30
