#!/bin/bash
for c in 8 16 20 40; do
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-check --no-shared-leg --e2e-chunk $c 2>/dev/null > gpurun_out/e2e_$c.json
  python - $c <<'PY'
import json, sys
c = sys.argv[1]
d = json.loads(open("gpurun_out/e2e_%s.json" % c).read())
print("chunk", c, "e2e MP/s", round(d["e2e"]["value"], 1), "ms", round(d["e2e"]["ms_per_step"], 2), "device ms", round(d["ms_per_step"], 2))
PY
done
