#!/bin/bash
# round 2, session 2, call 6: forward sample queue as a two-half ring (no compaction after a flush) vs the compacting form
mkdir -p gpurun_out
log=gpurun_out/b6_variants.log
: > $log
for v in base ring base ring; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== $v" >> $log
    for n in 8 40; do ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py $n 2>&1 | tail -1 >> $log; done
done
cp .variants/ring.so ava-256_b200/libmvpraymarch_b200.so
echo "== pytest ring" >> $log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $log
cat $log
