O=gpurun_out/final
mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 3 --warmup 3 --no-side > $O/bench_8gpu.json 2> $O/bench_8gpu.err; echo "bench rc=$?"; cut -c1-300 $O/bench_8gpu.json
