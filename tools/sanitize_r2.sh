#!/bin/bash
# compute-sanitizer over small renders of the kernels that are new in round 2 (16 / 32 lanes per pixel with their
# shared-memory strips, the fed / ordered pixel queue, the hybrid iteration): memcheck and racecheck
export MDZCUDA_DEBUG_POISON=0
S="compute-sanitizer --error-exitcode 7"
run() { echo "== $*"; "$@" 2>&1 | tail -4; echo "rc=$?"; }
run $S --tool memcheck python tools/run_case.py sea2048 --scale 0.04
run $S --tool racecheck python tools/run_case.py sea2048 --scale 0.03
run $S --tool racecheck python tools/run_case.py sea4096 --scale 0.02
run $S --tool synccheck python tools/run_case.py sea2048 --scale 0.03
run $S --tool racecheck python tools/run_case.py sea8192 --scale 0.02
run $S --tool racecheck python tools/run_case.py gmp1024 --scale 0.03
run $S --tool memcheck python tools/run_case.py gmp2048 --scale 0.03
run $S --tool memcheck python tools/run_case.py mini --scale 0.08 --order 1 --depth 3000
run $S --tool racecheck python tools/run_case.py mini --scale 0.05 --order 1 --depth 3000
run $S --tool memcheck python tools/run_case.py ld --scale 0.1 --order 1
run $S --tool memcheck python -m pytest "tests/test_multi_gpu.py::test_fed_plan_single_device" -q
run $S --tool memcheck python tools/run_case.py gmp320 --scale 0.06
