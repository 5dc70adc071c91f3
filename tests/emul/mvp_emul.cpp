// TEST INFRASTRUCTURE -- host build of the product kernels (ava-256_b200/csrc/mvp_kernels.cu and raydirs.cu, unmodified
// source) on top of the CPU emulation in cuda_emul.h.  Exports the same C-ABI entry points as the product library, taking
// host pointers; used by tests/test_emul_kernels.py to check the kernels' logic against the oracle without a GPU.
#define MVP_CPU_EMUL 1
#include "../../ava-256_b200/csrc/mvp_kernels.cu"
