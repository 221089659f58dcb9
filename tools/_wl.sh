OUT=gpurun_out; TAG=${1:-wl}; mkdir -p $OUT
timeout 600 python bench.py --workload intention --steps 30 --warmup 3 > $OUT/${TAG}_bench_intention.json 2> $OUT/${TAG}_intention.err; echo "rc=$?"; tail -3 $OUT/${TAG}_intention.err; cut -c1-200 $OUT/${TAG}_bench_intention.json
timeout 600 python bench.py --workload contact --steps 10 --warmup 3 > $OUT/${TAG}_bench_contact.json 2> $OUT/${TAG}_contact.err; echo "rc=$?"; cut -c1-200 $OUT/${TAG}_bench_contact.json
