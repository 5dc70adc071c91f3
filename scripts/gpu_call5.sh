#!/bin/bash
mkdir -p gpurun_out
scripts/try_variants.sh new:TP h64:TP h128:T new:T 2>&1 | tail -14
cp .variants/new.so ava-256_b200/libmvpraymarch_b200.so
echo "=== bench N=1"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err; tail -4 gpurun_out/c5_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/c5_bench.json').read())
print({k:d[k] for k in ('value','ms_per_step','kernel_ms')})
print('e2e', d['e2e']['value'], d['e2e']['ms_per_step'])
print('ref_cuda', {k:v for k,v in d['ref_cuda_baseline'].items() if k!='what'})
print('parity', d['parity_check']['ok'], d['parity_check']['fwd'], d['parity_check']['grads'])
PY
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-check > /dev/null 2>&1; tail -2 gpurun_out/r02_launches_bench.csv | cut -c1-200
