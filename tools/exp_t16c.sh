O=gpurun_out/t16
mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or 2048" 2>&1 | tail -2
for c in "sea2048 0.5" "sea4096 0.3334" "sea6144 0.25" "sea8192 0.25" "sea2048 0.25" "sea4096 0.1667"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-100; done | tee $O/times_c.txt
