OUT=gpurun_out; TAG=${1:-pl}; mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_launches_intention.csv \
   python bench.py --workload intention --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu.log 2>&1
tail -2 $OUT/${TAG}_ncu.log | cut -c1-300
