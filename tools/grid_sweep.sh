#!/bin/bash
# forward-kernel grid (CTAs) vs bench value / single-call rate; experiment knob CED_FWD_GRID (0 = the library's choice)
for grid in 0 296 370 444 518; do
  for rep in 1 2; do
    CED_FWD_GRID=$grid python bench.py --steps 40 --warmup 3 --no-cpu-baseline --no-e2e --no-subrecords 2>/dev/null > /tmp/gs.json
    python - "$grid" <<'PY'
import json, sys
d = json.loads(open("/tmp/gs.json").read().strip().splitlines()[-1])
print("grid %s: value %.1f single %.1f" % (sys.argv[1], d["value"], d["single_stream"]["value"]))
PY
  done
done
