OUT=gpurun_out; TAG=${1:-pol}; mkdir -p $OUT
timeout 300 python -m pytest tests/test_gpu_policy.py -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "rc=$?" >> $OUT/${TAG}_pytest.log
tail -30 $OUT/${TAG}_pytest.log
timeout 300 python bench.py --workload intention --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['policy'])"
TMJX_POLICY_V1=2 timeout 300 python bench.py --workload intention --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('v2', d['value'], d['policy'])"
