run() { echo "== $*"; python tools/run_case.py "$@" | cut -c1-62; }
for lib in "" exp/libv5.so; do
  export MDZCUDA_LIB=$lib; echo "#### lib=${lib:-product}"
  run gmp512 --scale 2
  run minigmp --scale 2 --order 1
  run gmp320 --scale 2
done
unset MDZCUDA_LIB
python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -1
