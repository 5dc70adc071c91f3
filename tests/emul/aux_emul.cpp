// TEST INFRASTRUCTURE -- host build of the product's auxiliary kernels (ava-256_b200/csrc/raydirs.cu and epilogue.cu,
// unmodified source) on the CPU emulation in cuda_emul.h: thread -> element mapping, vector paths and the block
// reduction of the colour-calibration gradients run on the CPU against their PyTorch restatements.
#define MVP_CPU_EMUL 1
#include "../../ava-256_b200/csrc/raydirs.cu"
#include "../../ava-256_b200/csrc/epilogue.cu"
