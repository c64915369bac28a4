timeout 300 python -m pytest tests/test_parity_gpu.py tests/test_multi_gpu.py -x -q -m gpu -k "wide or gmp_wide or fed" 2>&1 | tail -2
for c in "sea2048 0.5" "sea4096 0.3334"; do set -- $c; timeout 100 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-100; done
