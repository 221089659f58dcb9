OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > $OUT/final_bench.json 2> $OUT/final_bench.err; echo "bench rc=$?"; cut -c1-260 $OUT/final_bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > $OUT/final_bench_reference.json 2>> $OUT/final_bench.err; echo "ref rc=$?"; cut -c1-200 $OUT/final_bench_reference.json
