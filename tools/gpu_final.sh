#!/bin/bash
# One gpurun call at the end of a round: the matching-front tests / probe / ncu capture, then the full check
# (all GPU tests, bench line, ncu launch list, per-kernel ncu, surgery probe).  Usage: bash tools/gpu_final.sh <match tag> <tag>
bash tools/gpu_match.sh ${1:-m5}
bash tools/gpu_check.sh ${2:-r01v}
