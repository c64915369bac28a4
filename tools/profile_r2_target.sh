#!/bin/bash
# refresh of the target kernel's capture after the last change to the hybrid iteration + the launch list of the bench command
O=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-side --no-check > $O/r2p_launches_bench.json 2> $O/r2p_launches.err
ncu --set full --clock-control none --import-source on -f -o $O/r2p_mpfr512_target -k regex:escape -c 1 python tools/run_case.py mini --order 1 > /dev/null
python tools/ncu_summary.py $O/r2p_ncu_target.json mpfr512_target=$O/r2p_mpfr512_target.ncu-rep
ncu -i $O/r2p_mpfr512_target.ncu-rep --page source --csv 2>/dev/null > $O/r2p_target_source.csv
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2p_target_source.csv')))
print(rows[0][:12])
PY
rm -f $O/*.ncu-rep
ls -la $O | tail
