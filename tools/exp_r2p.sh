run() { python tools/run_case.py "$@" | cut -c1-58; }
for lib in "" exp/libv7.so exp/libv8.so; do
  export MDZCUDA_LIB=$lib; echo "#### lib=${lib:-product}"
  run mpfr320 --scale 2
  run sea320 --scale 2
  run sea256 --scale 2
  run sea192 --scale 2
  run sea160 --scale 2
  run sea384 --scale 2
  run cfg2p320 --scale 2
done
