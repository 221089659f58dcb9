OUT=gpurun_out; TAG=${1:-bench}
python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_reference.json 2> $OUT/${TAG}_reference.err; tail -c 400 $OUT/${TAG}_reference.json
python bench.py --steps 20 --warmup 3 > $OUT/${TAG}_bench.json 2> $OUT/${TAG}_bench.err; echo "bench rc=$?"; cat $OUT/${TAG}_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $OUT/${TAG}_launches.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_launch.log 2>&1
tail -5 $OUT/${TAG}_launches.csv
