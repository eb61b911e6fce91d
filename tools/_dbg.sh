cd /root/repo
OUT=gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -m gpu -x -q -k "attention" > $OUT/r02i_pytest.log 2>&1; echo "rc=$?" >> $OUT/r02i_pytest.log; tail -5 $OUT/r02i_pytest.log
timeout 300 python tools/attn_bench.py > $OUT/r02i_attn_bench.log 2>&1; cat $OUT/r02i_attn_bench.log
SAMK_LIB=/root/repo/sam_textvqa_b200/libsamk_tl.so BWD=1 timeout 120 python tools/attn_timeline.py 2>&1 | head -13
