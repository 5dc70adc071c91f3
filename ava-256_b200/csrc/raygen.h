// raygen.h -- one ray of a pinhole camera and its clip against the unit cube: compute_raydirs_forward_kernel of the reference
// (/root/reference/extensions/utils/utils_kernel.cu:32-46), as ONE device function that two translation units share:
//
//   * raydirs.cu       (no -use_fast_math, like the reference's utils extension): the stand-alone generator, writes the rays to HBM;
//   * mvp_kernels.cu   (-use_fast_math, like the reference's mvpraymarch extension): the render kernels' prologue generates the
//                      rays of its tile from the camera record instead of reading them.
//
// Both must produce the SAME bits -- a ray that differs in the last place moves samples across voxel cells -- so every operation
// below is an explicitly rounded intrinsic (no contraction, no dependence on the math flags of the translation unit): IEEE
// division, fused multiply-adds where written, a correctly rounded reciprocal square root.  The reference normalises with
// rnorm3df (<= 1 ulp); the result here is within 2 ulp of the reference's ray direction (tests: <= 2e-7 of max|raydir|).
#ifndef MVP_RAYGEN_H_
#define MVP_RAYGEN_H_

// per-view record the accel build derives from (viewpos, viewrot, focal, princpt, volradius): 64 bytes, read as 4 x float4
//   q0 = (o.x, o.y, o.z, R[0])   q1 = (R[1], R[2], R[3], R[4])   q2 = (R[5], R[6], R[7], R[8])   q3 = (princpt.x, princpt.y, focal.x, focal.y)
// with o = viewpos / volradius (utils_kernel.cu:32) and R = viewrot row-major.
struct MvpRayCam {
    float ox, oy, oz;
    float R[9];
    float pcx, pcy, fx, fy;
};

#ifdef MVP_CPU_EMUL
static inline float mvp_rsqrt_rn(float s) { return (float)(1.0 / std::sqrt((double)s)); }
#else
__device__ __forceinline__ float mvp_rsqrt_rn(float s) { return __frsqrt_rn(s); }
#endif

// pixel (px, py) -> unit direction d and the ray's parameter range inside [-1, 1]^3 (tmin clamped at 0, utils_kernel.cu:46)
__device__ __forceinline__ void mvp_gen_ray(const MvpRayCam &c, float px, float py, float &dx, float &dy, float &dz, float &tmin, float &tmax) {
    const float u = __fdiv_rn(__fsub_rn(px, c.pcx), c.fx);                                     // :36-37
    const float v = __fdiv_rn(__fsub_rn(py, c.pcy), c.fy);
    // raydir = viewrot0 * u + viewrot1 * v + viewrot2 * 1                                     :38-39
    // operation order as nvcc contracts the reference's expression (SASS of the sm_100 build: FMUL, FFMA, FADD): the rounding of
    // a direction component that nearly cancels decides the far-away clip distances of rays that run parallel to a cube face
    const float ex = __fadd_rn(__fmaf_rn(c.R[0], u, __fmul_rn(c.R[3], v)), c.R[6]);
    const float ey = __fadd_rn(__fmaf_rn(c.R[1], u, __fmul_rn(c.R[4], v)), c.R[7]);
    const float ez = __fadd_rn(__fmaf_rn(c.R[2], u, __fmul_rn(c.R[5], v)), c.R[8]);
    const float inv = mvp_rsqrt_rn(__fmaf_rn(ez, ez, __fmaf_rn(ey, ey, __fmul_rn(ex, ex))));    // :40 normalize
    dx = __fmul_rn(ex, inv); dy = __fmul_rn(ey, inv); dz = __fmul_rn(ez, inv);
    // unit-cube slab test                                                                     :42-46
    const float t1x = __fdiv_rn(__fsub_rn(-1.f, c.ox), dx), t1y = __fdiv_rn(__fsub_rn(-1.f, c.oy), dy), t1z = __fdiv_rn(__fsub_rn(-1.f, c.oz), dz);
    const float t2x = __fdiv_rn(__fsub_rn(1.f, c.ox), dx), t2y = __fdiv_rn(__fsub_rn(1.f, c.oy), dy), t2z = __fdiv_rn(__fsub_rn(1.f, c.oz), dz);
    tmin = fmaxf(fmaxf(fminf(t1x, t2x), fmaxf(fminf(t1y, t2y), fminf(t1z, t2z))), 0.f);
    tmax = fminf(fmaxf(t1x, t2x), fminf(fmaxf(t1y, t2y), fmaxf(t1z, t2z)));
}

#endif  // MVP_RAYGEN_H_
