// TEST INFRASTRUCTURE -- a small CPU emulation of the CUDA execution model, good enough to run the PRODUCT kernels'
// own source (ava-256_b200/csrc/mvp_kernels.cu, compiled with g++ and -DMVP_CPU_EMUL) in a container without a GPU,
// so that `pytest -m "not gpu"` exercises the real list-building / marching / adjoint logic against the oracle.
// Never built into, or loaded by, the product library.
//
// Model: one OS thread runs one thread block at a time; every CUDA thread of the block is a fiber (own stack, switched
// by a 10-instruction x86-64 context switch).  Fibers run round-robin and only yield inside collectives:
//   * warp collectives (__ballot_sync, __shfl*_sync, __any/__all_sync, __reduce_*_sync, __syncwarp) gather the 32 lanes of
//     a warp; exited lanes do not take part (the kernels only use full masks in warp-uniform control flow);
//   * __syncthreads / __syncthreads_or gather the block (exited threads count as arrived, as on the hardware).
// Blocks are distributed over a few OS threads; `__shared__` becomes `static thread_local` (one copy per OS thread =
// per running block); global atomics are real atomics.  Floating point is IEEE (no MUFU approximations, no FTZ), so the
// emulated kernels agree with the GPU to rounding, not bit for bit -- the same caveat as the C oracle.
#ifndef MVP_CUDA_EMUL_H_
#define MVP_CUDA_EMUL_H_

#include <cuda_runtime.h>   // vector types, dim3, cudaError_t, cudaStream_t (declarations only; nothing from libcudart is called)
#include <math_constants.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>
#include <vector>

#if !defined(__x86_64__)
#error "tests/emul/cuda_emul.h: the fiber switch is written for x86-64"
#endif

// ---- CUDA keywords --------------------------------------------------------------------------------------------------
#undef __shared__
#define __shared__ static thread_local
#undef __global__
#define __global__
#undef __device__
#define __device__
#undef __host__
#define __host__
#undef __forceinline__
#define __forceinline__ inline
#undef __launch_bounds__
#define __launch_bounds__(...)
#undef __restrict__
#define __restrict__

namespace emul {

constexpr int kMaxThreads = 1024;
constexpr size_t kStackBytes = 256 * 1024;

struct WarpSync {
    unsigned long long val[32];
    unsigned long long res[2][32];
    unsigned lanes_arrived_mask;
    unsigned res_mask[2];
    int arrived;
    unsigned gen;
};

struct Lane {
    void *sp;
    char *stack;
    uint3 tid;
    int linear;
    bool done;
};

struct Block {
    Lane lane[kMaxThreads];
    WarpSync warp[kMaxThreads / 32];
    int nthreads;
    int alive;
    int cur;
    uint3 bid;
    dim3 bdim, gdim;
    // block barrier
    int bar_arrived;
    unsigned bar_gen;
    int bar_or[2];
    int bar_acc;
    void *sched_sp;
    const std::function<void()> *body;
    char *dyn_smem;
    size_t dyn_bytes;
};

extern thread_local Block *g_blk;

extern "C" void emul_switch(void **save_sp, void *load_sp);

inline Lane &cur() { return g_blk->lane[g_blk->cur]; }
inline void yield() { Block *b = g_blk; emul_switch(&b->lane[b->cur].sp, b->sched_sp); }

void launch(dim3 grid, dim3 block, size_t dyn_smem, const std::function<void()> &body);
char *smem_anchor();

// number of lanes of warp w that have not exited
inline int warp_active(Block *b, int w) {
    int n = 0;
    const int lo = w * 32, hi = std::min(lo + 32, b->nthreads);
    for (int i = lo; i < hi; ++i) n += b->lane[i].done ? 0 : 1;
    return n;
}

// Every live lane of the warp deposits `v`; returns when all have, with everybody's values in out[0..31] and the mask of
// participating lanes.  Two result buffers (by generation parity) make back-to-back collectives safe.
inline unsigned warp_gather(unsigned long long v, unsigned long long (&out)[32]) {
    Block *b = g_blk;
    const int t = b->cur, w = t >> 5, l = t & 31;
    WarpSync &s = b->warp[w];
    const unsigned my = s.gen;
    s.val[l] = v;
    s.lanes_arrived_mask |= 1u << l;
    s.arrived++;
    if (s.arrived == warp_active(b, w)) {
        std::memcpy(s.res[my & 1], s.val, sizeof(s.val));
        s.res_mask[my & 1] = s.lanes_arrived_mask;
        s.arrived = 0;
        s.lanes_arrived_mask = 0;
        s.gen = my + 1;
    } else {
        while (s.gen == my) yield();
    }
    std::memcpy(out, s.res[my & 1], sizeof(out));
    return s.res_mask[my & 1];
}

// called by the scheduler when a lane exits: a collective the others are waiting in may now be complete
void lane_exited(Block *b, int t);

inline int block_barrier(int pred) {
    Block *b = g_blk;
    const unsigned my = b->bar_gen;
    b->bar_acc |= pred ? 1 : 0;
    b->bar_arrived++;
    if (b->bar_arrived == b->alive) {
        b->bar_or[my & 1] = b->bar_acc;
        b->bar_acc = 0;
        b->bar_arrived = 0;
        b->bar_gen = my + 1;
    } else {
        while (b->bar_gen == my) yield();
    }
    return b->bar_or[my & 1];
}

template <class T>
inline unsigned long long to_bits(T v) {
    static_assert(sizeof(T) <= 8, "collective payload too large");
    unsigned long long u = 0;
    std::memcpy(&u, &v, sizeof(T));
    return u;
}
template <class T>
inline T from_bits(unsigned long long u) {
    T v;
    std::memcpy(&v, &u, sizeof(T));
    return v;
}

}  // namespace emul

// ---- built-in variables (defined AFTER the CUDA headers: cudaLaunchConfig_t has members called gridDim / blockDim) ----
#define threadIdx (emul::cur().tid)
#define blockIdx (emul::g_blk->bid)
#define blockDim (emul::g_blk->bdim)
#define gridDim (emul::g_blk->gdim)

// ---- warp collectives --------------------------------------------------------------------------------------------
inline void __syncwarp(unsigned = 0xffffffffu) { unsigned long long o[32]; emul::warp_gather(0, o); }
inline unsigned __ballot_sync(unsigned, int pred) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(pred ? 1 : 0, o);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if (((m >> i) & 1) && o[i]) r |= 1u << i;
    return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned, int pred) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(pred ? 1 : 0, o);
    for (int i = 0; i < 32; ++i) if (((m >> i) & 1) && !o[i]) return 0;
    return 1;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
    unsigned long long o[32];
    emul::warp_gather(emul::to_bits(v), o);
    const int l = emul::g_blk->cur & 31;
    const int s = (l & ~(width - 1)) | (src & (width - 1));
    return emul::from_bits<T>(o[s]);
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lanemask, int width = 32) {
    unsigned long long o[32];
    emul::warp_gather(emul::to_bits(v), o);
    const int l = emul::g_blk->cur & 31;
    int s = l ^ lanemask;
    if ((s & ~(width - 1)) != (l & ~(width - 1))) s = l;
    return emul::from_bits<T>(o[s]);
}
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta, int width = 32) {
    unsigned long long o[32];
    emul::warp_gather(emul::to_bits(v), o);
    const int l = emul::g_blk->cur & 31;
    int s = l - (int)delta;
    if (s < 0 || (s & ~(width - 1)) != (l & ~(width - 1))) s = l;
    return emul::from_bits<T>(o[s]);
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
    unsigned long long o[32];
    emul::warp_gather(emul::to_bits(v), o);
    const int l = emul::g_blk->cur & 31;
    int s = l + (int)delta;
    if ((s & ~(width - 1)) != (l & ~(width - 1))) s = l;
    return emul::from_bits<T>(o[s]);
}
inline int __reduce_min_sync(unsigned, int v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    int r = v;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r = std::min(r, emul::from_bits<int>(o[i]));
    return r;
}
inline int __reduce_max_sync(unsigned, int v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    int r = v;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r = std::max(r, emul::from_bits<int>(o[i]));
    return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    unsigned r = v;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r = std::min(r, emul::from_bits<unsigned>(o[i]));
    return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    unsigned r = v;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r = std::max(r, emul::from_bits<unsigned>(o[i]));
    return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    unsigned r = 0;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r |= emul::from_bits<unsigned>(o[i]);
    return r;
}
inline unsigned __reduce_and_sync(unsigned, unsigned v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    unsigned r = 0xffffffffu;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r &= emul::from_bits<unsigned>(o[i]);
    return r;
}
inline int __reduce_add_sync(unsigned, int v) {
    unsigned long long o[32];
    const unsigned m = emul::warp_gather(emul::to_bits(v), o);
    int r = 0;
    for (int i = 0; i < 32; ++i) if ((m >> i) & 1) r += emul::from_bits<int>(o[i]);
    return r;
}
inline void __syncthreads() { emul::block_barrier(0); }
inline void __threadfence_block() {}
inline int __syncthreads_or(int pred) { return emul::block_barrier(pred); }

// ---- scalar intrinsics ---------------------------------------------------------------------------------------------
template <class T> inline T __ldg(const T *p) { return *p; }
template <class T> inline void __stcs(T *p, const T &v) { *p = v; }
template <class T> inline T __ldcs(const T *p) { return *p; }
inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
inline float __fmul_rn(float a, float b) { volatile float r = a * b; return r; }
inline float __fadd_rn(float a, float b) { volatile float r = a + b; return r; }
inline float __fsub_rn(float a, float b) { volatile float r = a - b; return r; }
inline float __fdiv_rn(float a, float b) { return a / b; }
// cvt.*.s32.f32 saturates and maps NaN to 0
inline int emul_sat_int(float x) { return x != x ? 0 : x >= 2147483648.f ? 0x7fffffff : x <= -2147483648.f ? (-0x7fffffff - 1) : (int)x; }
inline int __float2int_rd(float x) { return emul_sat_int(std::floor(x)); }
inline int __float2int_rn(float x) { return emul_sat_int(std::nearbyint(x)); }
inline int __float2int_rz(float x) { return emul_sat_int(std::trunc(x)); }
inline int __float_as_int(float x) { int i; std::memcpy(&i, &x, 4); return i; }
inline float __int_as_float(int i) { float x; std::memcpy(&x, &i, 4); return x; }
inline unsigned __float_as_uint(float x) { unsigned i; std::memcpy(&i, &x, 4); return i; }
inline float __uint_as_float(unsigned i) { float x; std::memcpy(&x, &i, 4); return x; }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
inline float __expf(float x) { return std::exp(x); }
inline float __powf(float a, float b) { return std::pow(a, b); }
inline float __fdividef(float a, float b) { return a / b; }
inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const char *)p - emul::smem_anchor()); }
inline void *__cvta_shared_to_generic(size_t off) { return emul::smem_anchor() + (ptrdiff_t)(int32_t)(uint32_t)off; }

using std::ceil;
using std::floor;
using std::isfinite;
using std::isnan;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline size_t min(size_t a, size_t b) { return a < b ? a : b; }
inline size_t max(size_t a, size_t b) { return a > b ? a : b; }
inline float min(float a, float b) { return std::fmin(a, b); }
inline float max(float a, float b) { return std::fmax(a, b); }

// ---- atomics on "global" memory (blocks run on several OS threads) -------------------------------------------------
inline float atomicAdd(float *p, float v) {
    std::atomic_ref<float> a(*p);
    float old = a.load(std::memory_order_relaxed);
    while (!a.compare_exchange_weak(old, old + v, std::memory_order_relaxed)) {}
    return old;
}
inline int atomicAdd(int *p, int v) { return std::atomic_ref<int>(*p).fetch_add(v, std::memory_order_relaxed); }
inline unsigned atomicAdd(unsigned *p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_add(v, std::memory_order_relaxed); }
inline int atomicOr(int *p, int v) { return std::atomic_ref<int>(*p).fetch_or(v, std::memory_order_relaxed); }
inline unsigned atomicOr(unsigned *p, unsigned v) { return std::atomic_ref<unsigned>(*p).fetch_or(v, std::memory_order_relaxed); }
inline int atomicMax(int *p, int v) {
    std::atomic_ref<int> a(*p);
    int old = a.load(std::memory_order_relaxed);
    while (old < v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
    return old;
}
inline int atomicMin(int *p, int v) {
    std::atomic_ref<int> a(*p);
    int old = a.load(std::memory_order_relaxed);
    while (old > v && !a.compare_exchange_weak(old, v, std::memory_order_relaxed)) {}
    return old;
}

// ---- launches / runtime ----------------------------------------------------------------------------------------------
#define MVP_LAUNCH(kern, grid, block, smem, st, ...) emul::launch(dim3(grid), dim3(block), (size_t)(smem), [&] { kern(__VA_ARGS__); })
#define MVP_EMUL_DYN_SMEM(type, name) type *name = reinterpret_cast<type *>(emul::g_blk->dyn_smem)
// "device" memory is host memory; nothing below touches libcudart
#define cudaMemsetAsync(p, v, n, st) (std::memset((p), (v), (n)), cudaSuccess)
#define cudaGetLastError() cudaSuccess
#define cudaGetErrorString(e) "CUDA error (emulated build)"
inline float rnorm3df(float a, float b, float c) { return 1.f / std::sqrt(a * a + b * b + c * c); }

#endif  // MVP_CUDA_EMUL_H_
