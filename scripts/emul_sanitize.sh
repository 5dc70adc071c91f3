#!/bin/bash
# AddressSanitizer + UBSan over the product kernels' source running on the CPU emulation (tests/emul): out-of-bounds
# accesses to the workspace / list / gradient buffers and integer UB show up here without a GPU (compute-sanitizer is
# the GPU-side counterpart).  usage: scripts/emul_sanitize.sh [-DMVP_...=1 ...]
set -e
cd "$(dirname "$0")/.."
lib=/tmp/libmvp_emul_sanitized.so
g++ -std=c++20 -O1 -g -fsanitize=address,undefined -fno-sanitize=float-cast-overflow,float-divide-by-zero -fno-omit-frame-pointer \
    -ffp-contract=off -shared -fPIC -pthread -Wno-attributes -Wno-unknown-pragmas -I/usr/local/cuda/include -Itests/emul -Iinclude \
    "$@" tests/emul/mvp_emul.cpp tests/emul/cuda_emul.cpp -o $lib
MVP_EMUL_SANITIZED_LIB=$lib MVP_EMUL_THREADS=1 ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
    LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" python scripts/emul_sanitize_run.py
