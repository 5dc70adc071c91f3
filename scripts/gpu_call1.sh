#!/bin/bash
# round-2 GPU call 1: parity at scale, variants, sanitizer, bench legs, fresh ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
echo "=== pytest gpu (default build)"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/c1_pytest.log
echo "=== variants"; scripts/try_variants.sh cur:T margin:T xb:T both:TP cur:T 2>&1 | tail -20
cp .variants/cur.so ava-256_b200/libmvpraymarch_b200.so
echo "=== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -3 gpurun_out/c1_bench.err; cut -c1-600 gpurun_out/c1_bench.json
echo "=== sanitizer"; SAN_TIMEOUT=400 scripts/gpu_sanitize.sh 2>&1 | tail -15
echo "=== ncu"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_ --launch-skip 4 -c 4 -f -o gpurun_out/prof_r2a python scripts/prof_step.py 2 1024 667 16384 8 2 2>&1 | tail -3
