#!/bin/bash
# tests + driver-style bench + reference arm + launch list (no ncu --set full, no sanitizer)
mkdir -p gpurun_out
echo "=== pytest"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/fin_pytest.log
echo "=== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== bench (driver flags)"; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/fin_bench_n1.json 2> gpurun_out/fin_bench_n1.err; tail -2 gpurun_out/fin_bench_n1.err
echo "=== reference arm"; timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > gpurun_out/fin_ref_n1.json 2>/dev/null; cut -c1-300 gpurun_out/fin_ref_n1.json
echo "=== launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-check --no-shared-leg > /dev/null 2>&1; tail -1 gpurun_out/r02_launches_bench.csv | cut -c1-120
