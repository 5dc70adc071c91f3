#!/bin/bash
mkdir -p gpurun_out
scripts/try_variants.sh norec:T rec4:TP rec5:T norec:T 2>&1 | tail -12
cp .variants/rec4.so ava-256_b200/libmvpraymarch_b200.so
echo "=== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; tail -4 gpurun_out/c3_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c3_bench.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernel_ms','allreduce_ms','rank_ms_per_step')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('ref_cuda', {k:v for k,v in d['ref_cuda_baseline'].items() if k!='what'})
print('parity', d['parity_check']['ok'], d['parity_check']['fwd'], d['parity_check']['grads'])
PY
