O=gpurun_out/final
mkdir -p $O
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu 2>&1 | tail -2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-side > $O/bench_2gpu.json 2> $O/bench_2gpu.err; echo "bench rc=$?"; cut -c1-300 $O/bench_2gpu.json
