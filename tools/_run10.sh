OUT=gpurun_out; TAG=${1:-r1o}
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > $OUT/${TAG}_bench2.json 2> $OUT/${TAG}_bench2.err; echo "rc=$?"; cat $OUT/${TAG}_bench2.json | cut -c1-400; tail -3 $OUT/${TAG}_bench2.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > $OUT/${TAG}_ref2.json 2> $OUT/${TAG}_ref2.err; echo "rc=$?"; cat $OUT/${TAG}_ref2.json | cut -c1-200
