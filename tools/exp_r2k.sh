run() { echo "== $*"; python tools/run_case.py "$@" | cut -c1-62; }
python -m pytest tests/test_parity_gpu.py -q -x 2>&1 | tail -2
run mpfr320 --scale 2
run sea320 --scale 2
run sea256 --scale 2
run mpfr128 --scale 2
run mpfr80 --scale 2
run cfg2p128 --scale 2
run cfg2p320 --scale 2
run mini --scale 4 --order 1
run mini --scale 4 --order 1 --stride 8
run misi --scale 2
run mpfr512 --scale 2
