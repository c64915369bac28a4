for sh in 0 808; do
  export MDZCUDA_COOP_SHAPE=$sh
  for c in "sea2048 0.5" "sea1100 0.5" "gmp1024 0.5"; do set -- $c; echo -n "shape $sh: "; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-200; done
done
export MDZCUDA_COOP_SHAPE=808
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or 2048 or gmp_1024" 2>&1 | tail -2
