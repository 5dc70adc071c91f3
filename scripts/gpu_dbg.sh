#!/bin/bash
mkdir -p gpurun_out
cp .variants/d64.so ava-256_b200/libmvpraymarch_b200.so
for n in 2 8; do echo "== d64 N=$n"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py $n 2>&1 | tail -12; echo "rc=${PIPESTATUS[0]}"; done
echo "== memcheck d64 1 view"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 400 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 8 python scripts/time_modes.py 1 2>&1 | tail -60
