set -x
python tools/run_case.py mini --reps 2
python tools/run_case.py mini --depth 20000 --reps 2
python tools/run_case.py minioff --reps 2
python tools/run_case.py misi --reps 2
MDZCUDA_SPEC_LEVEL=0 python tools/run_case.py mini --reps 1
MDZCUDA_SPEC_LEVEL=2 python tools/run_case.py mini --reps 1
MDZCUDA_SPEC_LEVEL=3 python tools/run_case.py mini --reps 1
python tools/run_case.py mini --chunk 32 --reps 1
python tools/run_case.py mini --bps 2 --reps 1
ncu --metrics smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_ncu_mini.csv python tools/run_case.py mini --scale 0.5 --reps 1
ncu --metrics smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__warps_active.avg.per_cycle_active,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,sm__cycles_elapsed.max,gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2b_ncu_misi.csv python tools/run_case.py misi --scale 0.5 --reps 1
