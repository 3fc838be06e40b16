#!/bin/bash
# A/B of the backward scheduling switches (GPS_B200_OPT bits) and of the attention staging (GPS_B200_ATTN_STAGE):
#   bash tools/ab_opt.sh "OPT STAGE" ...    e.g.  bash tools/ab_opt.sh "7 0" "7 1" "103 0" "103 1"
for cfg in "$@"; do
  set -- $cfg
  GPS_B200_OPT=$1 GPS_B200_ATTN_STAGE=$2 timeout 200 python bench.py --steps 200 --warmup 5 2>/dev/null > /tmp/ab.json
  python - "$1" "$2" <<'PY'
import json, sys
j = json.loads(open("/tmp/ab.json").read().strip().splitlines()[-1])
print("GPS_B200_OPT", sys.argv[1], "ATTN_STAGE", sys.argv[2], "ms/step", round(j["ms_per_step"], 4), "stack", round(j["stack"].get("ms_per_step", 0), 4),
      "launches", j["gpu_launches"], "sm_mhz", j["clocks"]["sm_mhz"])
PY
done
