OUT=gpurun_out; TAG=${1:-sc}; mkdir -p $OUT
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --workload contact > $OUT/${TAG}_bench_contact_8gpu.json 2> $OUT/${TAG}_bench_contact_8gpu.err
echo "contact rc=$?"; cut -c1-200 $OUT/${TAG}_bench_contact_8gpu.json
