"""Runs named parity cases (tests/helpers.py CASES) forward + backward through the public op on cuda:0 and prints a checksum
per case -- the workload for compute-sanitizer (scripts/gpu_sanitize.sh) and quick smoke runs.  No oracle involved."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.helpers import build_case  # noqa: E402
from tests.test_gpu_parity import run_ours  # noqa: E402

for name in sys.argv[1:] or ["head_small", "many_overlaps", "warp_small"]:
    s, grad = build_case(name)
    out, grads = run_ours(s, grad)
    torch.cuda.synchronize()
    print("%-16s rayrgba sum %.6e  grads %s" % (name, float(out.sum()), " ".join("%.4e" % float(abs(g).sum()) for g in grads)), flush=True)
