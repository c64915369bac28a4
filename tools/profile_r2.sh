#!/bin/bash
# Round-2 evidence run (one B200): launch list of the bench command, ncu --set full captures of the final kernels,
# the N = 3 compaction A/B, the resources table.  Everything lands in gpurun_out/r2p_*.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r2p_launches.csv python bench.py --steps 2 --warmup 3 --no-side --no-check > $O/r2p_launches_bench.json 2> $O/r2p_launches.err
$NCU -o $O/r2p_mpfr512_target -k regex:escape -c 1 python tools/run_case.py mini --order 1 > /dev/null
$NCU -o $O/r2p_ld64_phase0 -k regex:escape -c 1 python tools/run_case.py ld > /dev/null
$NCU -o $O/r2p_ld64_phase1 -k regex:escape --launch-skip 1 -c 1 python tools/run_case.py ld > /dev/null
$NCU -o $O/r2p_mpfr320 -k regex:escape -c 1 python tools/run_case.py mpfr320 > /dev/null
$NCU -o $O/r2p_mpfr128 -k regex:escape -c 1 python tools/run_case.py mpfr128 > /dev/null
$NCU -o $O/r2p_mpfr80 -k regex:escape -c 1 python tools/run_case.py mpfr80 > /dev/null
$NCU -o $O/r2p_gmp512 -k regex:escape -c 1 python tools/run_case.py gmp512 > /dev/null
$NCU -o $O/r2p_mpfr1024 -k regex:escape -c 1 python tools/run_case.py sea1024 --scale 0.5 > /dev/null
$NCU -o $O/r2p_coop2048 -k regex:escape -c 1 python tools/run_case.py sea2048 --scale 0.125 > /dev/null
for p in 0 1; do python tools/run_case.py mpfr80 --scale 2 --park $p --reps 3 | cut -c1-70; python tools/run_case.py sea96 --scale 2 --park $p --reps 3 | cut -c1-70; done > $O/r2p_n3_parking.txt 2>&1
python tools/resources_table.py > $O/r2p_kernel_resources.md 2>&1
# the reports are ~10 MB each and gpurun brings back 64 MiB at most: summarise here, keep the numbers
python tools/ncu_summary.py $O/r2p_ncu_summary.json mpfr512_target=$O/r2p_mpfr512_target.ncu-rep ld64_cfg2_phase0=$O/r2p_ld64_phase0.ncu-rep \
    ld64_cfg2_phase1=$O/r2p_ld64_phase1.ncu-rep mpfr320_dej_960x540=$O/r2p_mpfr320.ncu-rep mpfr128_seahorse_960x540=$O/r2p_mpfr128.ncu-rep \
    mpfr80_seahorse_960x540=$O/r2p_mpfr80.ncu-rep gmp512_seahorse_960x540=$O/r2p_gmp512.ncu-rep mpfr1024_seahorse_480x270=$O/r2p_mpfr1024.ncu-rep \
    mpfr2048_warp_per_pixel_120x67=$O/r2p_coop2048.ncu-rep
rm -f $O/*.ncu-rep
ls -la $O | tail -20
