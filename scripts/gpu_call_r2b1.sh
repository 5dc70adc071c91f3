#!/bin/bash
# round 2, session 2, call 1: camera-ray fusion on the GPU (tests first), kernel timings with / without it, full suite, short bench
mkdir -p gpurun_out
echo "=== camera tests"; timeout 600 python -m pytest tests/test_gpu_camera_rays.py tests/test_gpu_raydirs_and_properties.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/b1_camera_tests.log
echo "=== time_modes"; for c in 0 1 0 1; do CAMERA=$c ALPHA_MU=17 ALPHA_SIGMA=6 timeout 120 python scripts/time_modes.py 2>&1 | tail -1; done | tee gpurun_out/b1_time_modes.log
echo "=== pytest all"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/b1_pytest.log
echo "=== bench"; timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/b1_bench.json 2> gpurun_out/b1_bench.err; tail -3 gpurun_out/b1_bench.err; cut -c1-400 gpurun_out/b1_bench.json
