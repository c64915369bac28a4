for stride in 8 4 2; do
for bps in 0 1 2 3 4; do for order in 0 1; do
  echo -n "stride $stride bps $bps order $order: "; python tools/run_case.py ld --stride $stride --bps $bps --order $order --reps 3 | cut -c1-22 | tr '\n' ' '; echo
done; done; done
