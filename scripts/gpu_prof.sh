#!/bin/bash
# fresh ncu --set full capture of the four render kernels (2 views of the bench scene) -> gpurun_out/prof_$1.ncu-rep
mkdir -p gpurun_out
ALPHA_MU=17 ALPHA_SIGMA=6 timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ --launch-skip 4 -c 4 -f -o gpurun_out/prof_$1 python scripts/prof_step.py 2 1024 667 16384 8 2 2>&1 | tail -2
