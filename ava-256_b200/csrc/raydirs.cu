// raydirs.cu -- pinhole ray generation + unit-cube clip, the step right before the raymarcher
// (SURVEY.md section 8f row 1).  Replaces /root/reference/extensions/utils/utils_kernel.cu:12-52
// (compute_raydirs_forward_kernel); the reference's backward kernel is an empty stub and its Python backward returns
// None for every input (extensions/utils/utils.py:45-46), so there is nothing else to build.
//
// Compiled WITHOUT -use_fast_math, like the reference's utils extension (extensions/utils/setup.py has no such flag).
// The per-ray arithmetic is raygen.h's mvp_gen_ray -- explicitly rounded operations only, because the render kernels' prologue
// (mvp_kernels.cu, compiled WITH -use_fast_math) generates the same rays from the same function and must get the same bits.
// Pure streaming kernel: 32 B written per ray, no reads to speak of
// -> HBM-write bound; one thread per ray, x fastest for coalesced 12/12/8-byte stores.
#ifdef MVP_CPU_EMUL   // test-only host build on the CPU emulation (tests/emul/), see mvp_kernels.cu
#include "cuda_emul.h"
#else
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#include "mvpraymarch_b200.h"
#include "raygen.h"

namespace {

__global__ void __launch_bounds__(256) compute_raydirs_kernel(int N, int H, int W, const float *__restrict__ viewpos,
                                                              const float *__restrict__ viewrot, const float *__restrict__ focal,
                                                              const float *__restrict__ princpt, const float2 *__restrict__ pixelcoords,
                                                              float volradius, float *__restrict__ raypos, float *__restrict__ raydir,
                                                              float2 *__restrict__ tminmax) {
    const int w = blockIdx.x * blockDim.x + threadIdx.x;
    const int h = blockIdx.y;
    const int n = blockIdx.z;
    if (w >= W) return;
    MvpRayCam c;
    // raypos = viewpos / volradius                                           utils_kernel.cu:32
    c.ox = __fdiv_rn(viewpos[n * 3 + 0], volradius); c.oy = __fdiv_rn(viewpos[n * 3 + 1], volradius); c.oz = __fdiv_rn(viewpos[n * 3 + 2], volradius);
#pragma unroll
    for (int i = 0; i < 9; ++i) c.R[i] = viewrot[n * 9 + i];
    c.pcx = princpt[n * 2 + 0]; c.pcy = princpt[n * 2 + 1]; c.fx = focal[n * 2 + 0]; c.fy = focal[n * 2 + 1];
    const size_t r = ((size_t)n * H + h) * W + w;
    const float2 pc = pixelcoords ? __ldg(pixelcoords + r) : make_float2((float)w, (float)h);
    float dx, dy, dz, tmin, tmax;
    mvp_gen_ray(c, pc.x, pc.y, dx, dy, dz, tmin, tmax);                       // :36-46, shared with the render kernels' prologue
    const float rx = c.ox, ry = c.oy, rz = c.oz;
    raypos[r * 3 + 0] = rx; raypos[r * 3 + 1] = ry; raypos[r * 3 + 2] = rz;
    raydir[r * 3 + 0] = dx; raydir[r * 3 + 1] = dy; raydir[r * 3 + 2] = dz;
    tminmax[r] = make_float2(tmin, tmax);
}

// ---- Morton codes of (normalised) slab centres: compute_morton of the reference (mvpraymarch.cpp:106-121; bvh.cu:20-57) ----
// quantise each coordinate of a point of the unit cube to 10 bits and interleave them x | y | z from the top bit down
__device__ __forceinline__ unsigned spread10(unsigned v) {
    // 10 bits -> every third bit of 30: four shift-or-mask rounds (16, 8, 4, 2 positions), written as multiplies
    v = (v | (v << 16)) & 0xFF0000FFu;
    v = (v | (v << 8)) & 0x0F00F00Fu;
    v = (v | (v << 4)) & 0xC30C30C3u;
    v = (v | (v << 2)) & 0x49249249u;
    return v;
}
__global__ void __launch_bounds__(256) compute_morton_kernel(size_t NK, const float *__restrict__ centre, int *__restrict__ code) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NK) return;
    const float qx = fminf(fmaxf(centre[i * 3 + 0] * 1024.f, 0.f), 1023.f);
    const float qy = fminf(fmaxf(centre[i * 3 + 1] * 1024.f, 0.f), 1023.f);
    const float qz = fminf(fmaxf(centre[i * 3 + 2] * 1024.f, 0.f), 1023.f);
    code[i] = (int)((spread10((unsigned)qx) << 2) | (spread10((unsigned)qy) << 1) | spread10((unsigned)qz));
}

}  // namespace

extern "C" int mvp_compute_morton(int32_t N, int32_t K, const float *centre, int32_t *code, void *stream) {
    if (!centre || !code) return MVP_ERR_NULL;
    if (N < 1 || K < 1) return MVP_ERR_SHAPE;
    const size_t NK = (size_t)N * K;
#ifdef MVP_CPU_EMUL
    MVP_LAUNCH(compute_morton_kernel, (unsigned)((NK + 255) / 256), 256, 0, stream, NK, centre, code);
#else
    compute_morton_kernel<<<(unsigned)((NK + 255) / 256), 256, 0, (cudaStream_t)stream>>>(NK, centre, code);
#endif
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MVP_OK : (int)e;
}

extern "C" int mvp_compute_raydirs(int32_t N, int32_t H, int32_t W, const float *viewpos, const float *viewrot, const float *focal,
                                   const float *princpt, const float *pixelcoords, float volradius, float *raypos, float *raydir,
                                   float *tminmax, void *stream) {
    if (!viewpos || !viewrot || !focal || !princpt || !raypos || !raydir || !tminmax) return MVP_ERR_NULL;
    if (N < 1 || H < 1 || W < 1 || H > 65535 || N > 65535) return MVP_ERR_SHAPE;
    dim3 grid((W + 255) / 256, H, N);
#ifdef MVP_CPU_EMUL
    MVP_LAUNCH(compute_raydirs_kernel, grid, 256, 0, stream, N, H, W, viewpos, viewrot, focal, princpt,
               reinterpret_cast<const float2 *>(pixelcoords), volradius, raypos, raydir, reinterpret_cast<float2 *>(tminmax));
#else
    compute_raydirs_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(N, H, W, viewpos, viewrot, focal, princpt,
                                                                  reinterpret_cast<const float2 *>(pixelcoords), volradius, raypos,
                                                                  raydir, reinterpret_cast<float2 *>(tminmax));
#endif
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? MVP_OK : (int)e;
}
