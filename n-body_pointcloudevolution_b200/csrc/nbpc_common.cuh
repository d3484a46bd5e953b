// nbpc_common.cuh - shared plumbing for libnbpc.so (sm_100a).
//
// Two build modes:
//   * default: real CUDA (nvcc -gencode arch=compute_100a,code=sm_100a).
//   * -DNBPC_HOST_EMU: TEST INFRASTRUCTURE ONLY.  The barrier-free "v0" kernels are compiled
//     as plain C++ and each launch becomes a sequential loop over blocks/threads, so index
//     algebra can be checked on a machine without a GPU (tests/test_host_emulation.py loads
//     tests/_emu/libnbpc_emu.so directly; the product package never does).
#pragma once
#include <stddef.h>
#include <stdint.h>
#include <math.h>
#include <string.h>
#include <stdio.h>

#include <string>

#include "../../include/nbpc.h"

#ifdef NBPC_HOST_EMU
// ------------------------------------------------------------------ host emulation shim
struct dim3 {
    unsigned x, y, z;
    dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float4 { float x, y, z, w; };
struct float2 { float x, y; };
static inline float4 make_float4(float x, float y, float z, float w) { float4 r = {x, y, z, w}; return r; }
extern thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;
typedef void *cudaStream_t;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T *p) { return *p; }
static inline int atomicAdd(int *p, int v) { int o = *p; *p = o + v; return o; }
static inline unsigned atomicAdd(unsigned *p, unsigned v) { unsigned o = *p; *p = o + v; return o; }
static inline unsigned atomicMin(unsigned *p, unsigned v) { unsigned o = *p; if (v < o) *p = v; return o; }
static inline unsigned atomicMax(unsigned *p, unsigned v) { unsigned o = *p; if (v > o) *p = v; return o; }
static inline int atomicOr(int *p, int v) { int o = *p; *p = o | v; return o; }
static inline double __dmul_rn(double a, double b) { return a * b; }   // built with -ffp-contract=off
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline float __fmul_rn(float a, float b) { return a * b; }
static inline float __fadd_rn(float a, float b) { return a + b; }
static inline float __fdiv_rn(float a, float b) { return a / b; }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
template <class T> static inline T nbpc_min(T a, T b) { return a < b ? a : b; }
template <class T> static inline T nbpc_max(T a, T b) { return a > b ? a : b; }
template <class K, class... A>
static inline void nbpc_emu_launch(K kern, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    for (unsigned bz = 0; bz < grid.z; ++bz) for (unsigned by = 0; by < grid.y; ++by) for (unsigned bx = 0; bx < grid.x; ++bx)
        for (unsigned tz = 0; tz < block.z; ++tz) for (unsigned ty = 0; ty < block.y; ++ty) for (unsigned tx = 0; tx < block.x; ++tx) {
            blockIdx = dim3(bx, by, bz); threadIdx = dim3(tx, ty, tz);
            kern(args...);
        }
}
#define NBPC_LAUNCH_N(name, kern, grid, block, smem, stream, ...) nbpc_emu_launch(kern, dim3(grid), dim3(block), __VA_ARGS__)
#define NBPC_LAUNCH(kern, grid, block, smem, stream, ...) NBPC_LAUNCH_N(#kern, kern, grid, block, smem, stream, __VA_ARGS__)
static inline int nbpc_memset_async(void *p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
static inline int nbpc_launch_status() { return 0; }
#else
// ------------------------------------------------------------------ real CUDA
#include <cuda_runtime.h>
template <class T> __host__ __device__ __forceinline__ T nbpc_min(T a, T b) { return a < b ? a : b; }
template <class T> __host__ __device__ __forceinline__ T nbpc_max(T a, T b) { return a > b ? a : b; }
// Every kernel launch goes through this macro: it counts launches (nbpc_launch_count) and, when
// profiling is enabled (nbpc_prof_enable), brackets the launch with CUDA events on the launching
// stream so that bench.py can report per-kernel device times (nbpc_prof_report).
extern int g_nbpc_prof_on;
extern unsigned long long g_nbpc_launches;
void nbpc_prof_pre(const char *name, cudaStream_t stream);
void nbpc_prof_post(cudaStream_t stream);
#define NBPC_LAUNCH_N(name, kern, grid, block, smem, stream, ...)       \
    do {                                                                \
        ++g_nbpc_launches;                                              \
        if (g_nbpc_prof_on) nbpc_prof_pre(name, stream);                \
        kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);       \
        if (g_nbpc_prof_on) nbpc_prof_post(stream);                     \
    } while (0)
#define NBPC_LAUNCH(kern, grid, block, smem, stream, ...) NBPC_LAUNCH_N(#kern, kern, grid, block, smem, stream, __VA_ARGS__)
static inline int nbpc_memset_async(void *p, int v, size_t n, cudaStream_t s) {
    return cudaMemsetAsync(p, v, n, s) == cudaSuccess ? 0 : 1;
}
#endif

// per-device caches (occupancy / function attributes belong to a device): slot of the current device in a table of
// NBPC_MAX_DEVICES entries (devices beyond the table share the last slot and are re-configured on every call by callers
// that check `slot < NBPC_MAX_DEVICES - 1`)
#define NBPC_MAX_DEVICES 64
static inline int nbpc_device_slot() {
#ifdef NBPC_HOST_EMU
    return 0;
#else
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return dev < 0 ? 0 : (dev >= NBPC_MAX_DEVICES ? NBPC_MAX_DEVICES - 1 : dev);
#endif
}

// kernel name carrying the layer widths ("gl_edge_out_kernel[k=32,q=16]"); only built while profiling
struct NbpcKName {
    char buf[96];
    const char *base;
    NbpcKName(const char *b, int k, int q) : base(b) {
        buf[0] = 0;
#ifndef NBPC_HOST_EMU
        if (g_nbpc_prof_on) snprintf(buf, sizeof(buf), "%s[k=%d,q=%d]", b, k, q);
#endif
    }
    const char *c_str() const { return buf[0] ? buf : base; }
};

// arithmetic of the edge-level GEMMs (nbpc_set_math_mode / NBPC_MATH)
extern int g_nbpc_math_mode;

// ------------------------------------------------------------------ errors
void nbpc_set_error(const std::string &msg);
int nbpc_check_launch(const char *where);   // cudaGetLastError() -> NBPC_ELAUNCH (no sync)
int nbpc_require_sm100();                   // NBPC_EARCH unless the current device is sm_100

#define NBPC_ARG(cond, msg)                                                       \
    do {                                                                          \
        if (!(cond)) {                                                            \
            nbpc_set_error(std::string(__func__) + ": invalid argument: " + (msg)); \
            return NBPC_EINVAL;                                                   \
        }                                                                         \
    } while (0)

#define NBPC_TRY(expr)            \
    do {                          \
        int _rc = (expr);         \
        if (_rc != NBPC_OK) return _rc; \
    } while (0)

static inline size_t nbpc_align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }
static inline int nbpc_cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// bump allocator over a caller-owned workspace
struct NbpcArena {
    char *base;
    size_t off, cap;
    NbpcArena(void *p, size_t n) : base((char *)p), off(0), cap(n) {}
    template <class T> T *take(size_t count) {
        size_t bytes = nbpc_align_up(count * sizeof(T));
        T *r = (T *)(base ? base + off : nullptr);
        off += bytes;
        return r;
    }
    bool ok() const { return off <= cap; }
};
