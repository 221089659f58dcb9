OUT=gpurun_out; TAG=${1:-sc}; mkdir -p $OUT
nvidia-smi -L | head -8
for N in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2950$N bench.py --gpus $N --steps 50 --warmup 3 > $OUT/${TAG}_bench_${N}gpu.json 2> $OUT/${TAG}_bench_${N}gpu.err
  echo "N=$N rc=$?"; cut -c1-200 $OUT/${TAG}_bench_${N}gpu.json
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --workload contact > $OUT/${TAG}_bench_contact_8gpu.json 2> $OUT/${TAG}_bench_contact_8gpu.err
echo "contact rc=$?"; cut -c1-200 $OUT/${TAG}_bench_contact_8gpu.json
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --workload intention > $OUT/${TAG}_bench_intention_8gpu.json 2> $OUT/${TAG}_bench_intention_8gpu.err
echo "intention rc=$?"; cut -c1-200 $OUT/${TAG}_bench_intention_8gpu.json
