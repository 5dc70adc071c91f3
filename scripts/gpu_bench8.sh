#!/bin/bash
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/c12_bench8.json 2> gpurun_out/c12_bench8.err
tail -2 gpurun_out/c12_bench8.err | cut -c1-200
python - <<'PY'
import json
d = json.loads(open("gpurun_out/c12_bench8.json").read().strip().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step", "kernel_ms", "allreduce_ms")})
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "cam", d["e2e_camera_inputs"]["value"], d["e2e_camera_inputs"]["ms_per_step"])
print("shared", d["shared_primitives_config"]["ms_per_step"])
PY
