run() { echo "== $*"; python tools/run_case.py "$@" | cut -c1-88; }
for lib in "" exp/libv2.so; do
  export MDZCUDA_LIB=$lib; echo "#### lib=${lib:-product}"
  run mini --scale 2
  run mini --scale 2 --order 1
  run misi --scale 2
  run mpfr512 --scale 2
  run mpfr320 --scale 2
  run mpfr128 --scale 2
  run mpfr80 --scale 2
  run sea256 --scale 2
  run ld
done
unset MDZCUDA_LIB
run mini --scale 4 --order 1
run mini --scale 4 --order 1 --stride 8
run mini --scale 4 --order 0 --stride 8
