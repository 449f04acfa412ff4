#!/bin/bash
# source-level capture of one launch of each kernel matching the regex (reduced batch); CSV exports for tools/ncu_regions.py
# usage: bash tools/gpu_prof_kernels.sh tag 'regex' [filters]
TAG=${1:-pk}; K=${2:-front}; NF=${3:-296}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$K" -c 4 -o $OUT/prof python bench.py --steps 1 --warmup 1 --filters $NF --no-cpu > $OUT/ncu.log 2>&1; echo "ncu rc=$?"
ncu -i $OUT/prof.ncu-rep --page source --csv --print-source sass > $OUT/src.csv 2>$OUT/src.err
ncu -i $OUT/prof.ncu-rep --page raw --csv > $OUT/raw.csv 2>>$OUT/src.err
ls -la $OUT
