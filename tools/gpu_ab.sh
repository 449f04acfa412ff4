#!/bin/bash
# A/B of an environment switch on the bench: bash tools/gpu_ab.sh tag VAR=VALUE
TAG=${1:-ab}; KV=$2
OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 600 python bench.py --steps 5 --no-cpu > $OUT/a.json 2> $OUT/a.err
env $KV timeout 600 python bench.py --steps 5 --no-cpu > $OUT/b.json 2> $OUT/b.err
env $KV timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python - <<PY
import json
for n in ("a","b"):
    try:
        d=json.load(open("$OUT/%s.json"%n)); print(n, "ms/step", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],3), "bounds", round(d["gate_bounds"]["ms_per_step"],3))
    except Exception as e:
        print(n, "failed", e, open("$OUT/%s.err"%n).read()[-500:])
PY
