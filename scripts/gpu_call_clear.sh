#!/bin/bash
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "=== 8 views"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 300 python scripts/time_modes.py 8 2>&1 | tail -1
echo "=== bench"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/clr_bench_n1.json 2> gpurun_out/clr_bench_n1.err; tail -2 gpurun_out/clr_bench_n1.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/clr_bench_n1.json'))
print(d['value'], d['ms_per_step'], d['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['shared_primitives_config']['ms_per_step'], d['parity_check']['ok'])
PY
