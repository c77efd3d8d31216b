/*
 * tests/hostsim/hostsim.h — TEST SCAFFOLDING ONLY (never built by build(), never loaded by the
 * package).  Lets g++ compile libsmatrix_b200/csrc/smx_kernels.cu and smx_host.c unchanged so
 * that the CPU test-suite can exercise the HOST LOGIC (round loop, growth, chunking, phases)
 * without a GPU: every "kernel" runs as a sequential loop over a grid of 1-thread blocks with a
 * warp width of 1.  It checks logic, not concurrency and not performance; the product library
 * is the nvcc build and refuses to run without a CUDA device.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
using std::sqrt;

#define SMX_WARP 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __shared__ static thread_local /* two host threads may run 'kernels' at once */
#define __restrict__
#define __launch_bounds__(...)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct uint3_sim { unsigned x, y, z; };
extern thread_local uint3_sim blockIdx, threadIdx;
extern thread_local dim3 gridDim, blockDim;

template <class F>
static inline void smx_sim_launch(dim3 g, dim3 b, F f) {
  gridDim = g;
  blockDim = b;
  for (unsigned bz = 0; bz < g.z; bz++)
    for (unsigned by = 0; by < g.y; by++)
      for (unsigned bx = 0; bx < g.x; bx++)
        for (unsigned tx = 0; tx < b.x; tx++) {
          blockIdx = {bx, by, bz};
          threadIdx = {tx, 0, 0};
          f();
        }
}
#define SMX_LAUNCH(kern, grid, block, stream, ...) \
  smx_sim_launch(dim3(grid), dim3(block), [&] { kern(__VA_ARGS__); })

typedef unsigned long long sim_ull;
template <class T> static inline T atomicCAS(T* p, T cmp, T val) { T o = *p; if (o == cmp) *p = val; return o; }
template <class T> static inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }
template <class T> static inline T atomicSub(T* p, T v) { T o = *p; *p = o - v; return o; }
template <class T> static inline T atomicOr(T* p, T v) { T o = *p; *p = o | v; return o; }
template <class T> static inline T atomicAnd(T* p, T v) { T o = *p; *p = o & v; return o; }
template <class T> static inline T atomicMax(T* p, T v) { T o = *p; if (v > o) *p = v; return o; }
template <class T> static inline T atomicMin(T* p, T v) { T o = *p; if (v < o) *p = v; return o; }
template <class T> static inline T atomicExch(T* p, T v) { T o = *p; *p = v; return o; }

template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
static inline unsigned __activemask() { return 1u; }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
template <class T> static inline T __ldcg(const T* p) { return *p; }
