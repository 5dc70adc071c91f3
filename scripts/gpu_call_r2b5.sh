#!/bin/bash
# round 2, session 2, call 5: what one rank of an 8-GPU run does per step (10 views), kernel by kernel; sum_views tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_epilogue.py tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/b5_pytest.log
timeout 300 python bench.py --views 10 --steps 20 --warmup 5 --no-e2e --no-check --no-cpu-baseline > gpurun_out/b5_bench_v10.json 2> gpurun_out/b5_bench_v10.err; cut -c1-200 gpurun_out/b5_bench_v10.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 150 --csv --log-file gpurun_out/b5_launches_v10.csv python bench.py --views 10 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-shared-leg > /dev/null 2>&1; tail -1 gpurun_out/b5_launches_v10.csv | cut -c1-100
