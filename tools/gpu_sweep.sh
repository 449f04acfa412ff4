#!/bin/bash
# sweep of environment switches on the bench (front-end stage times only): bash tools/gpu_sweep.sh "A=1 B=2" "A=3" ...
for KV in "$@"; do
  env $KV timeout 300 python bench.py --steps 3 --no-cpu > /tmp/sw.json 2> /tmp/sw.err
  python - "$KV" <<'PY'
import json, sys
try:
    d = json.load(open("/tmp/sw.json")); sm = d["roofline"]["stage_ms"]
    print(sys.argv[1], "| step", round(d["ms_per_step"], 2), "asm", round(sm.get("front_asm", 0), 2), "chol", round(sm.get("front_chol", 0), 2), "front", round(sm["front"], 2))
except Exception as e:
    print(sys.argv[1], "failed", e, open("/tmp/sw.err").read()[-300:])
PY
done
