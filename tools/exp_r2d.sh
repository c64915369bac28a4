for c in misi mini mpfr512; do
  for lv in 1 0 2; do
    echo "== $c SPEC_LEVEL=$lv"; MDZCUDA_SPEC_LEVEL=$lv python tools/run_case.py $c --scale 2 --reps 1 | cut -c1-90
  done
done
for lv in 1 0 2 3; do echo "== dej320 SPEC_LEVEL=$lv"; MDZCUDA_SPEC_LEVEL=$lv python tools/run_case.py mpfr320 --scale 2 --reps 1 | cut -c1-90; done
for lv in 1 0; do echo "== mini bps2 SPEC_LEVEL=$lv"; MDZCUDA_SPEC_LEVEL=$lv python tools/run_case.py mini --scale 2 --bps 2 --reps 1 | cut -c1-90; done
