cd /root/repo
OUT=gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -s -k "slice_of_128 or long_sequence or flat_adam or beam" > $OUT/r02k_pytest.log 2>&1; echo "rc=$?" >> $OUT/r02k_pytest.log; tail -25 $OUT/r02k_pytest.log
