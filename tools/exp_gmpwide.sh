O=gpurun_out/gmpw
mkdir -p $O
timeout 1200 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "gmp" > $O/parity.log 2>&1; echo "parity rc=$?"; tail -4 $O/parity.log
for c in "gmp1024 0.25" "gmp2048 0.25" "gmp4096 0.1667"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-230; done | tee $O/times.txt
S="compute-sanitizer --error-exitcode 7"
(timeout 400 $S --tool racecheck python tools/run_case.py gmp1024 --scale 0.03 2>&1 | tail -2; timeout 300 $S --tool memcheck python tools/run_case.py gmp1024 --scale 0.04 2>&1 | tail -2) | tee $O/sanitize.txt
