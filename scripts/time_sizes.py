"""fwd / bwd kernel time (C-ABI, accel prebuilt) as a function of the number of views per launch: exposes the per-launch fixed cost
(tail of the slowest warps) that limits strong scaling.  python scripts/time_sizes.py [views ...]"""
import subprocess
import sys
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for n in (sys.argv[1:] or ["2", "5", "10", "20", "40"]):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "time_modes.py"), n], capture_output=True, text=True,
                         env=dict(os.environ, ALPHA_MU="17", ALPHA_SIGMA="6"))
    print((out.stdout.strip().splitlines() or [out.stderr.strip()[-300:]])[-1], flush=True)
