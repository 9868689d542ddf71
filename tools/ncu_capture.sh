#!/bin/bash
# Targeted `ncu --set full` captures of the hot kernels of one bench step (run under gpurun, one GPU).
# Reports are post-processed ON the box into small CSVs; a .ncu-rep above 20 MB is dropped (gpurun_out/ <= 64 MiB).
set -u
CMD="${CMD:-python bench.py --steps 1 --warmup 3 --no-cpu --no-train}"
OUT=gpurun_out
$CMD > $OUT/plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/plain.log; exit 1; }
cap() {  # name regex skip count
  ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -f -o $OUT/$1 $CMD > $OUT/$1.log 2>&1
  ncu -i $OUT/$1.ncu-rep --page raw --csv > $OUT/$1_raw.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page source --csv > $OUT/$1_source.csv 2>/dev/null
  ncu -i $OUT/$1.ncu-rep --page details > $OUT/$1_details.txt 2>/dev/null
  sz=$(stat -c %s $OUT/$1.ncu-rep 2>/dev/null || echo 0)
  if [ "$sz" -gt 20000000 ]; then rm -f $OUT/$1.ncu-rep; fi
  ls -la $OUT/$1*
}
for spec in "$@"; do
  IFS=: read name regex skip count <<< "$spec"
  cap "$name" "$regex" "$skip" "$count"
done
