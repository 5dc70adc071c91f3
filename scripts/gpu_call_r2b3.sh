#!/bin/bash
# round 2, session 2, call 3: forward with the compositing deferred by one batch (under the next batch's gathers)
mkdir -p gpurun_out
log=gpurun_out/b3_variants.log
: > $log
for v in base d1 d1m6 d2 d2m6 base; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== $v" >> $log
    for n in 8 40; do ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py $n 2>&1 | tail -1 >> $log; done
done
for v in d1 d2; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== pytest $v" >> $log
    timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $log
done
cp .variants/base.so ava-256_b200/libmvpraymarch_b200.so
cat $log
