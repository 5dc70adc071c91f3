#!/bin/bash
# Experiment driver (GPU box): for each prebuilt library variant under .variants/, time the render kernels
# (scripts/time_modes.py, bench scene) and run the GPU test suite; results accumulate in gpurun_out/variants.log.
# usage: scripts/try_variants.sh "name:T" "name:TP" ...   (T = time, P = pytest)
mkdir -p gpurun_out
log=gpurun_out/variants.log
: > $log
for spec in "$@"; do
    v=${spec%%:*}; what=${spec##*:}
    cp .variants/$v.so ava-256_b200/libmvpraymarch_b200.so
    echo "== $v" >> $log
    case $what in *T*) ALPHA_MU=17 ALPHA_SIGMA=6 timeout 60 python scripts/time_modes.py 2>&1 | tail -1 >> $log;; esac
    case $what in *P*) timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2 >> $log;; esac
done
cat $log
