OUT=gpurun_out; TAG=${1:-prof}
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tmjx_env_kernel -s 4 -c 1 -f -o $OUT/${TAG}_newton_prof \
      python bench.py --workload contact --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_ncu_newton.log 2>&1
ls -la $OUT/${TAG}_newton_prof.ncu-rep
