// gb_hd.h -- one include for the device headers.  Under nvcc this is just <cuda_runtime.h>.
//
// With -DGB_HOSTSIM the CUDA keywords and the handful of intrinsics the emulator uses are given plain C++
// meanings, so that g++ can compile the *device* headers unchanged into tests/hostsim (a debugging harness that
// steps one env at a time on the CPU and is compared with the oracle by the `not gpu` tests).  That harness is
// test infrastructure only: libgbenv.so is never built with GB_HOSTSIM and has no host execution path.
#pragma once
#include <stdint.h>

#if !defined(GB_HOSTSIM)
#include <cuda_runtime.h>
#define GB_NOINLINE __noinline__
// raw PRMT (generic mode): unlike __byte_perm the selector is not masked with 0x7777 first
__device__ __forceinline__ uint32_t gb_prmt(uint32_t a, uint32_t b, uint32_t s) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(s));
    return r;
}
#else
#include <stddef.h>
#include <string.h>
#define __VECTOR_TYPES_H__
struct uint2 { uint32_t x, y; };
struct uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { uint2 r = {x, y}; return r; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { uint4 r = {x, y, z, w}; return r; }
struct gb_dim3 { int x, y, z; };
static gb_dim3 threadIdx = {0, 0, 0}, blockIdx = {0, 0, 0}, blockDim = {1, 1, 1}, gridDim = {1, 1, 1};
#define __device__
#define __host__
#define __global__ static
#define __constant__ static
#define __shared__ static
#define __forceinline__ inline
#define GB_NOINLINE __attribute__((noinline))
#define __restrict__
#define __launch_bounds__(...)
static inline void __syncthreads() {}
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) {}
static inline unsigned __activemask() { return 1u; }
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s) {
    uint64_t src = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; i++) {
        uint32_t n = (s >> (4 * i)) & 0xF, byte = (uint32_t)(src >> (8 * (n & 7))) & 0xFF;
        if (n & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
        r |= byte << (8 * i);
    }
    return r;
}
static inline uint32_t gb_prmt(uint32_t a, uint32_t b, uint32_t s) { return __byte_perm(a, b, s); }
static inline uint32_t __brev(uint32_t x) {
    uint32_t r = 0;
    for (int i = 0; i < 32; i++) r |= ((x >> i) & 1u) << (31 - i);
    return r;
}
static inline int __clz(uint32_t x) { return x ? __builtin_clz(x) : 32; }
static inline int __ffs(uint32_t x) { return __builtin_ffs((int)x); }
static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
template <typename T> static inline T atomicAdd(T *p, T v) { T o = *p; *p = o + v; return o; }
#endif
