#!/bin/bash
# GPU box: compute-sanitizer memcheck + racecheck (+ synccheck) over the render kernels on three small cases that cover the
# 256- and 512-entry variants, list reuse, the TMA-staged bucket scan, algo 1 and the camera-ray prologue.  Summaries land in gpurun_out/ (copy the
# ones to keep into profiles/).  The shared-memory hand-offs under test: forward sample queue / flush, backward sample ring,
# mbarrier double buffer of the bucket staging (csrc/mvp_kernels.cu).
mkdir -p gpurun_out
CS=${CS:-/usr/local/cuda/bin/compute-sanitizer}
cases="head_small many_overlaps warp_small camera_head"
for tool in memcheck racecheck synccheck; do
    out=gpurun_out/sanitizer_$tool.log
    timeout ${SAN_TIMEOUT:-600} $CS --tool $tool --print-limit 20 python scripts/run_cases.py $cases > $out 2>&1
    echo "== $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $out | tail -1)"
    grep -E "^(head_small|many_overlaps|warp_small|camera_head)" $out
done
