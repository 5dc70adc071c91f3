#!/bin/bash
# final single-GPU evidence run of the round: tests, driver-style bench, ncu captures, sanitizer
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/fin_pytest.log
echo "=== bench (driver flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; tail -2 gpurun_out/fin_bench_n1.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/fin_ref_n1.json 2>/dev/null; cut -c1-300 gpurun_out/fin_ref_n1.json
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-shared-leg > /dev/null 2>&1; tail -1 gpurun_out/r02_launches_bench.csv | cut -c1-120
echo "=== ncu full"; scripts/gpu_prof.sh r2f
echo "=== sanitizer"; SAN_TIMEOUT=400 scripts/gpu_sanitize.sh 2>&1 | tail -12
