run() { echo "== $*"; python tools/run_case.py "$@" | cut -c1-62; }
for lib in "" exp/libv3.so; do
  export MDZCUDA_LIB=$lib; echo "#### lib=${lib:-product}"
  run mpfr320 --scale 2
  run sea320 --scale 2
  run sea256 --scale 2
  run mpfr128 --scale 2
  run mpfr80 --scale 2
  run sea96 --scale 2
  run cfg2p128 --scale 2
  run cfg2p320 --scale 2
done
