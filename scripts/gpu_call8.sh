#!/bin/bash
mkdir -p gpurun_out
scripts/try_variants.sh vm1:TP vm0:T lm5:TP lm4:T vm1:T 2>&1 | tail -16
for v in vm0 vm1; do cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so; echo "== sizes $v"; python scripts/time_sizes.py 2 10 40; done
