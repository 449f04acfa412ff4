#!/bin/bash
# One gpurun call for the matching front: its GPU tests, the surgery / matching probe, one ncu capture of the two kernels.
TAG=${1:-m1}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_matching.py -x -q > $OUT/pytest_match.log 2>&1; echo "pytest rc=$?"; tail -25 $OUT/pytest_match.log
timeout 300 python tools/tracks_probe.py 4096 > $OUT/tracks_probe.txt 2>&1; tail -11 $OUT/tracks_probe.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'match_descriptors|verify_matches' -c 6 \
    -o $OUT/prof_match python tools/tracks_probe.py 1024 > $OUT/ncu_match.log 2>&1; echo "ncu rc=$?"
ls -la $OUT
