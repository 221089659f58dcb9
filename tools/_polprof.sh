OUT=gpurun_out; TAG=${1:-pp}; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:linear_tf32_tma_kernel -s 15 -c 2 -f -o $OUT/${TAG}_linear_prof \
   python bench.py --workload intention --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_lin.log 2>&1
ls -la $OUT/${TAG}_linear_prof.ncu-rep
