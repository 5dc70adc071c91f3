#!/bin/bash
# round 2, session 2, call 4: instruction-count micro-optimisations (constant slab size, integer inside test, lanemask_lt)
mkdir -p gpurun_out
log=gpurun_out/b4_variants.log
: > $log
for v in base micro base micro; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== $v" >> $log
    for n in 8 40; do ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py $n 2>&1 | tail -1 >> $log; done
done
for v in base micro; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== ncu $v" >> $log
    ALPHA_MU=17 ALPHA_SIGMA=6 timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:render_ --launch-skip 4 -c 4 python scripts/prof_step.py 2 1024 667 16384 8 2 2>&1 | grep -E "render_.*256|gpu__time|smsp__" >> $log
done
cp .variants/micro.so ava-256_b200/libmvpraymarch_b200.so
echo "== pytest micro" >> $log
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $log
cat $log
