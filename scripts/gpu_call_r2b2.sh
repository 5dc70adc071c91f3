#!/bin/bash
# round 2, session 2, call 2: L1 / shared-memory carve-out experiments (prebuilt variants under .variants/)
mkdir -p gpurun_out
log=gpurun_out/b2_variants.log
: > $log
for v in base u bc43 bc57 bc28u176 bc28u192 fc43u224 fc71 base; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== $v" >> $log
    for n in 8 40; do ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py $n 2>&1 | tail -1 >> $log; done
done
for v in base bc43 bc28u176 fc43u224; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== ncu $v" >> $log
    ALPHA_MU=17 ALPHA_SIGMA=6 timeout 300 ncu --metrics launch__occupancy_limit_shared_mem,launch__shared_mem_config_size,l1tex__t_sector_hit_rate.pct,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:render_ --launch-skip 4 -c 4 python scripts/prof_step.py 2 1024 667 16384 8 2 2>&1 | grep -E "render_|launch__|l1tex__|gpu__time|smsp__inst" >> $log
done
for v in bc28u176 fc43u224; do
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== pytest $v" >> $log
    timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $log
done
cp .variants/base.so ava-256_b200/libmvpraymarch_b200.so
cat $log
