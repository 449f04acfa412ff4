#!/bin/bash
# One gpurun call: GPU parity tests, a bench line, the ncu launch list of the same command and one full capture per kernel.
# Usage (from the repo root on the GPU box): bash tools/gpu_check.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > $OUT/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
tail -c 3000 $OUT/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
# one launch of every kernel of the step at the launch shape of the full-size run (1024 filters = one front-end sub-batch)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'propagate|front_asm|front_chol|gram_kernel|update_smem|update_info' -c 5 \
    -o $OUT/prof python bench.py --steps 1 --warmup 1 --filters 1024 --no-cpu > $OUT/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python tools/tracks_probe.py 4096 > $OUT/tracks_probe.txt 2>&1; tail -9 $OUT/tracks_probe.txt
ls -la $OUT
