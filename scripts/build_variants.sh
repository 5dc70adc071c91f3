#!/bin/bash
# Builds library variants into .variants/ (git-ignored, shipped to the GPU box by gpurun) for scripts/try_variants.sh.
# usage: scripts/build_variants.sh name="-DFLAG=1 -DOTHER=0" ...        e.g.
#   scripts/build_variants.sh cur="" margin="-DMVP_LIST_MARGIN=1" xb="-DMVP_XBUCKETS=1" both="-DMVP_LIST_MARGIN=1 -DMVP_XBUCKETS=1"
#   gpurun --timeout 300 -- 'scripts/try_variants.sh cur:T margin:TP xb:TP both:TP cur:T'
set -e
cd "$(dirname "$0")/.."
mkdir -p .variants
for spec in "$@"; do
    name=${spec%%=*}; flags=${spec#*=}
    MVP_NVCC_EXTRA="$flags" python ava-256_b200/build.py -f > /dev/null
    cp ava-256_b200/libmvpraymarch_b200.so .variants/$name.so
    echo "built .variants/$name.so  [$flags]"
done
python ava-256_b200/build.py -f > /dev/null      # leave the default build in place
