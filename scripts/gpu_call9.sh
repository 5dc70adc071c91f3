#!/bin/bash
mkdir -p gpurun_out
scripts/try_variants.sh ord1:TP ord0:T 2>&1 | tail -8
for v in ord0 ord1; do cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so; echo "== sizes $v"; python scripts/time_sizes.py 2 10 40; done
