#!/bin/bash
mkdir -p gpurun_out
for v in cap128 async; do
  cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
  echo "== $v"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py 2 2>&1 | tail -15; echo "rc=$?"
done
cp .variants/cap128.so ava-256_b200/libmvpraymarch_b200.so
echo "== memcheck cap128 1 view"; ALPHA_MU=17 ALPHA_SIGMA=6 timeout 300 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 5 python scripts/time_modes.py 1 2>&1 | tail -40
