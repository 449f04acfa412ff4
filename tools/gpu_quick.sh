#!/bin/bash
# quick iteration: GPU tests + one bench line (+ optional ncu of one kernel: bash tools/gpu_quick.sh tag kernel_regex [filters])
TAG=${1:-q}; K=$2; NF=${3:-296}
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; tail -4 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 5 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; tail -c 800 $OUT/bench.err
python -c "
import json; d=json.load(open('$OUT/bench.json')); print('ms/step', d['ms_per_step'], d['roofline']['stage_ms'], 'value', d['value'], 'e2e', d['e2e']['value'])"
if [ -n "$K" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -c 1 -o $OUT/prof python bench.py --steps 1 --warmup 1 --filters $NF --no-cpu > $OUT/ncu.log 2>&1
fi
