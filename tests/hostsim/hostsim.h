/*
 * tests/hostsim/hostsim.h — TEST SCAFFOLDING ONLY (never built by build(), never loaded by the
 * package).  Lets g++ compile libsmatrix_b200/csrc/smx_kernels.cu and smx_host.c unchanged so
 * that the CPU test-suite can exercise the HOST LOGIC (round loop, growth, chunking, phases)
 * without a GPU: every "kernel" runs as a sequential loop over a grid of 1-thread blocks with a
 * warp width of 1.  It checks logic, not concurrency and not performance; the product library
 * is the nvcc build and refuses to run without a CUDA device.
 * With -DSMX_SIM_WARP32 (warp_sched.cpp) blocks have 256 threads in warps of 32: every thread is a
 * fiber and the warp / block collectives exchange values as specified, so the warp-cooperative
 * device code runs too (slowly): tests/test_hostlogic_warp32.py.
 */
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cmath>
using std::sqrt;

#ifdef SMX_SIM_WARP32 /* the lock-step 32-lane variant: warp_sched.cpp */
#define SMX_WARP 32
#else
#define SMX_WARP 1
#endif
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

#ifdef SMX_SIM_WARP32
enum { SIM_SHFL = 0, SIM_SHFL_UP, SIM_SHFL_XOR, SIM_BALLOT, SIM_MATCH, SIM_SYNCWARP, SIM_ACTIVE, SIM_BLOCK };
uint64_t smx_sim_collective(int kind, unsigned mask, uint64_t in, int arg, const void* site);
void smx_sim_run_block(unsigned nthreads, void (*entry)(void*), void* arg);
template <class F>
static inline void smx_sim_launch(dim3 g, dim3 b, F f) { /* every thread of a block is a fiber; blocks run one after the other */
  gridDim = g;
  blockDim = b;
  for (unsigned bz = 0; bz < g.z; bz++)
    for (unsigned by = 0; by < g.y; by++)
      for (unsigned bx = 0; bx < g.x; bx++) {
        blockIdx = {bx, by, bz};
        smx_sim_run_block(b.x, [](void* p) { (*(F*)p)(); }, &f);
      }
}
#else
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
#endif
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

#ifdef SMX_SIM_WARP32
template <class T> static inline uint64_t sim_bits(T v) { uint64_t b = 0; static_assert(sizeof(T) <= 8, "shuffle width"); memcpy(&b, &v, sizeof v); return b; }
template <class T> static inline T sim_from(uint64_t b) { T v; memcpy(&v, &b, sizeof v); return v; }
template <class T> static inline T __shfl_sync(unsigned m, T v, int src) { return sim_from<T>(smx_sim_collective(SIM_SHFL, m, sim_bits(v), src, 0)); }
template <class T> static inline T __shfl_up_sync(unsigned m, T v, unsigned d) { return sim_from<T>(smx_sim_collective(SIM_SHFL_UP, m, sim_bits(v), (int)d, 0)); }
template <class T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return sim_from<T>(smx_sim_collective(SIM_SHFL_XOR, m, sim_bits(v), x, 0)); }
static inline unsigned __ballot_sync(unsigned m, int p) { return (unsigned)smx_sim_collective(SIM_BALLOT, m, p != 0, 0, 0); }
static inline int __any_sync(unsigned m, int p) { return __ballot_sync(m, p) != 0u; }
template <class T> static inline unsigned __match_any_sync(unsigned m, T v) { return (unsigned)smx_sim_collective(SIM_MATCH, m, sim_bits(v), 0, 0); }
static inline void __syncwarp(unsigned m = 0xffffffffu) { (void)smx_sim_collective(SIM_SYNCWARP, m, 0, 0, 0); }
static inline void __syncthreads() { (void)smx_sim_collective(SIM_BLOCK, 0, 0, 0, 0); }
static inline void __threadfence_block() {}
/* the lanes that wait at this textual call with the same key when nothing else in the block can advance */
#define SMX_ACTIVEMASK(key) \
  ((unsigned)smx_sim_collective(SIM_ACTIVE, 0, (uint64_t)(uintptr_t)(key), 0, [] { static char tag; return (const void*)&tag; }()))
#else
template <class T> static inline T __shfl_sync(unsigned, T v, int) { return v; }
template <class T> static inline T __shfl_up_sync(unsigned, T v, unsigned) { return v; }
template <class T> static inline T __shfl_xor_sync(unsigned, T v, int) { return v; }
static inline unsigned __ballot_sync(unsigned, int p) { return p ? 1u : 0u; }
static inline int __any_sync(unsigned, int p) { return p != 0; }
template <class T> static inline unsigned __match_any_sync(unsigned, T) { return 1u; }
static inline void __syncwarp(unsigned = 0xffffffffu) {}
static inline void __syncthreads() {}
#define SMX_ACTIVEMASK(key) 1u
#endif
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline int __clz(unsigned v) { return v ? __builtin_clz(v) : 32; }
static inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
template <class T> static inline T __ldcg(const T* p) { return *p; }
