OUT=gpurun_out; TAG=${1:-pol}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_policy.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
