#!/bin/bash
# Same-box A/B of an environment switch:  bash tools/ab_env.sh VAR val1 val2 ...   (C3 bench, 200 steps, each value twice)
var=$1; shift
for rep in 1 2; do
for v in "$@"; do
  env $var=$v timeout 200 python bench.py --steps 200 --warmup 5 2>/dev/null | tail -1 > /tmp/ab.json
  python - "$var" "$v" <<'PY'
import json, sys
j = json.loads(open("/tmp/ab.json").read())
print(sys.argv[1], sys.argv[2], "ms/step", round(j["ms_per_step"], 4), "stack", round(j["stack"].get("ms_per_step", 0), 4), "launches", j["gpu_launches"])
PY
done
done
