O=gpurun_out/t16
mkdir -p $O
for c in "sea2048 0.5" "sea4096 0.3334" "sea8192 0.25"; do set -- $c; timeout 300 python tools/run_case.py $1 --scale $2 --reps 2 2>&1 | tail -1 | cut -c1-120; done | tee $O/times_big.txt
for c in "sea2048 0.125" "sea4096 0.08"; do set -- $c
  timeout 600 ncu --set full --clock-control none --import-source on -f -o $O/$1 -k regex:escape -c 1 python tools/run_case.py $1 --scale $2 > /dev/null 2>&1
  python tools/ncu_summary.py $O/ncu_$1.json $1=$O/$1.ncu-rep > /dev/null 2>&1
  ncu -i $O/$1.ncu-rep --page raw --csv 2>/dev/null > $O/$1_raw.csv
  ncu -i $O/$1.ncu-rep --page source --csv 2>/dev/null > $O/$1_source.csv
done
rm -f $O/*.ncu-rep
ls -la $O
