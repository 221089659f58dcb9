# quick GPU A/B: parity tests + timing sweep with the tuning knobs
OUT=gpurun_out; TAG=${1:-ab}; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> $OUT/${TAG}_pytest.log
tail -15 $OUT/${TAG}_pytest.log
(
  timeout 300 python tools/gpu_perf_sweep.py 4096 8192
) > $OUT/${TAG}_sweep.log 2>&1
cat $OUT/${TAG}_sweep.log
timeout 600 python bench.py --workload contact --steps 10 --warmup 3 2>/dev/null | cut -c1-160
