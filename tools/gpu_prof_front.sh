#!/bin/bash
# source-level capture of one front_warp launch (reduced batch), exported as CSV for tools/ncu_regions.py / ncu_region_stalls.py
TAG=${1:-pf}; NF=${2:-296}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:front_warp -c 1 -o $OUT/prof python bench.py --steps 1 --warmup 1 --filters $NF --no-cpu > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof.ncu-rep --page source --csv --print-source sass > $OUT/src.csv 2>$OUT/src.err
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>>$OUT/src.err
ls -la $OUT
