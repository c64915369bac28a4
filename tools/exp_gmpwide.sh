O=gpurun_out/gmpw
mkdir -p $O
timeout 1500 python -m pytest tests/test_parity_gpu.py tests/test_dropin_gpu.py -x -q -m gpu -k "gmp or precision" > $O/parity.log 2>&1; echo "parity rc=$?"; tail -4 $O/parity.log
for c in "gmp576 1" "gmp768 1" "gmp896 1" "gmp1024 0.5" "gmp2048 0.25"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-200; done | tee $O/times.txt
python tools/resources_table.py > $O/resources.md 2> $O/resources.err; tail -24 $O/resources.md
