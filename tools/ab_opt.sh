#!/bin/bash
# A/B of the backward scheduling switches (GPS_B200_OPT bits): prints ms/step of the C3 bench for each value given.
for o in "$@"; do
  GPS_B200_OPT=$o timeout 200 python bench.py --steps 100 --warmup 5 2>/dev/null > /tmp/ab_$o.json
  python - "$o" <<'PY'
import json, sys
o = sys.argv[1]
j = json.loads(open(f"/tmp/ab_{o}.json").read().strip().splitlines()[-1])
print("GPS_B200_OPT", o, "ms/step", round(j["ms_per_step"], 4), "stack", round(j["stack"].get("ms_per_step", 0), 4), "launches", j["gpu_launches"])
PY
done
