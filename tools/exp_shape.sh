O=gpurun_out/final
mkdir -p $O
timeout 900 python -m pytest tests/test_parity_gpu.py tests/test_dropin_gpu.py -x -q -m gpu -k "wide or 2048 or gmp_1024 or precision" 2>&1 | tail -2
for c in "sea1100 0.5" "sea1536 0.5" "sea2048 0.5" "gmp1024 0.5" "gmp1344 0.5"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-200; done | tee $O/times_8x6.txt
S="compute-sanitizer --error-exitcode 7"
(echo "== racecheck sea1100 (8 x 6)"; timeout 400 $S --tool racecheck python tools/run_case.py sea1100 --scale 0.03 2>&1 | grep -E "SUMMARY|rep 0" | cut -c1-160; echo "== memcheck gmp1024 (8 x 6)"; timeout 400 $S --tool memcheck python tools/run_case.py gmp1024 --scale 0.03 2>&1 | grep -E "SUMMARY|rep 0" | cut -c1-160) | tee $O/sanitize_8x6.txt
