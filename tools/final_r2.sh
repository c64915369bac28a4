# end-of-round verification on one B200: the GPU suite, smoke(), the default bench line and the reference arm
O=gpurun_out/final
mkdir -p $O
timeout 1500 python -m pytest tests -x -q -m gpu > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > $O/bench_1gpu.json 2> $O/bench_1gpu.err; echo "bench rc=$?"; cut -c1-600 $O/bench_1gpu.json
if [ "$1" = "all" ]; then
python tools/resources_table.py > $O/resources.md 2> $O/resources.err; tail -10 $O/resources.md
S="compute-sanitizer --error-exitcode 7"
(for t in racecheck synccheck memcheck; do echo "== $t sea2048 (8 x 8)"; timeout 400 $S --tool $t python tools/run_case.py sea2048 --scale 0.03 2>&1 | grep -E "SUMMARY|rep 0" | cut -c1-160; done; echo "== racecheck gmp1024 (8 x 8)"; timeout 400 $S --tool racecheck python tools/run_case.py gmp1024 --scale 0.03 2>&1 | grep -E "SUMMARY|rep 0" | cut -c1-160) | tee $O/sanitize_8x8.txt
fi
