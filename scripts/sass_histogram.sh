#!/bin/bash
# Opcode histogram of the shipped library's SASS (per kernel + the TMA / mbarrier / vector-reduction opcodes that back the
# claims in DESIGN.md section 3).  Runs without a GPU.  usage: scripts/sass_histogram.sh > profiles/rNN_sass_opcodes.txt
set -e
cd "$(dirname "$0")/.."
LIBF=ava-256_b200/libmvpraymarch_b200.so
echo "# cuobjdump -sass $LIBF  ($(python -c "import sys; sys.path.insert(0,'.'); from ava256_b200 import lib; print(lib.LIB.mvp_build_config().decode())"))"
cuobjdump -sass $LIBF | python3 -c '
import re, sys, collections
kern = None
ops = collections.OrderedDict()
for line in sys.stdin:
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = m.group(1); ops[kern] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Za-z0-9_.]+)", line)
    if m and kern:
        ops[kern][m.group(1)] += 1
import subprocess
def demangle(n):
    try: return subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    except Exception: return n
KEY = ("UBLKCP", "UBLKPF", "SYNCS", "REDG", "RED.", "ATOMG", "ATOMS", "LDG.E.128", "LDS.128", "MUFU", "VOTE", "SHFL", "REDUX", "ACQBULK", "FENCE", "CCTL", "STL", "LDL")
for k, c in ops.items():
    tot = sum(c.values())
    name = demangle(k)
    print("\n== %s\n   %d instructions" % (name[:140], tot))
    keyops = {o: n for o, n in c.items() if any(o.startswith(p) or p in o for p in KEY)}
    print("   key opcodes: " + ", ".join("%s x%d" % (o, n) for o, n in sorted(keyops.items())))
    print("   top: " + ", ".join("%s x%d" % (o, n) for o, n in c.most_common(12)))
'
