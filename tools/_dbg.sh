cd /root/repo
OUT=gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "beam or argmax or greedy or cached" > $OUT/r02h_pytest.log 2>&1; echo "rc=$?" >> $OUT/r02h_pytest.log; tail -30 $OUT/r02h_pytest.log
