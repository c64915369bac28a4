mkdir -p gpurun_out/t16
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "wide or 2048" > gpurun_out/t16/parity.log 2>&1; echo "parity rc=$?" 
tail -3 gpurun_out/t16/parity.log
for c in "sea2048 0.25" "sea4096 0.1667" "sea6144 0.125" "sea8192 0.125" "sea1536 0.25" "sea3072 0.1667"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1; done | tee gpurun_out/t16/times.txt
S="compute-sanitizer --error-exitcode 7"
(timeout 300 $S --tool racecheck python tools/run_case.py sea2048 --scale 0.03 2>&1 | tail -3; timeout 300 $S --tool racecheck python tools/run_case.py sea4096 --scale 0.02 2>&1 | tail -3;  timeout 300 $S --tool memcheck python tools/run_case.py sea2048 --scale 0.04 2>&1 | tail -3; timeout 300 $S --tool synccheck python tools/run_case.py sea2048 --scale 0.03 2>&1 | tail -3) | tee gpurun_out/t16/sanitize.txt
timeout 600 python -m pytest tests/test_dropin_gpu.py -x -q -m gpu -k "2048 or 16384" 2>&1 | tail -3
