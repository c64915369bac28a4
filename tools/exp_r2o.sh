run() { echo "== $*"; python tools/run_case.py "$@" | cut -c1-62; }
for lib in "" exp/libv6.so; do
  export MDZCUDA_LIB=$lib; echo "#### lib=${lib:-product}"
  run mini --scale 2 --order 1
  run misi --scale 2
  run mpfr512 --scale 2
  run sea384 --scale 2
done
