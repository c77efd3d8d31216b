// scripts/probe_foot.cu — measurement helper (not product code): random-sector request rate vs
// footprint (TLB reach?) and vs independent loads in flight per thread (outstanding-miss limit?).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__device__ __forceinline__ ull mix(ull z){ z += 0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
__device__ __forceinline__ ull ld32(const char* p){ ull c0,c1,c2,c3; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); return c0^c1^c2^c3; }
__device__ __forceinline__ ull ld64B(const char* p){ ull c0,c1; asm volatile("ld.global.cg.L2::64B.v2.u64 {%0,%1}, [%2];":"=l"(c0),"=l"(c1):"l"(p):"memory"); return c0^c1; }
template<int ILP, int MODE> __global__ void __launch_bounds__(256) k(char* buf, ull nsec, ull n, ull* sink){
  ull acc=0; ull tid=blockIdx.x*(ull)blockDim.x+threadIdx.x, nth=(ull)gridDim.x*blockDim.x;
  for(ull a=tid;a<n;a+=nth*ILP){
    const char* p[ILP];
    #pragma unroll
    for(int j=0;j<ILP;j++) p[j] = buf + (mix(a+j*nth)%nsec)*32;
    #pragma unroll
    for(int j=0;j<ILP;j++){
      if(MODE==0) acc ^= ld32(p[j]);
      if(MODE==1) acc ^= ld64B(p[j]);
      if(MODE==2) atomicAdd((unsigned*)p[j], 1u);
      if(MODE==3) atomicCAS((ull*)p[j], 0ull, a|1ull);
      if(MODE==4) { acc ^= ld32(p[j]); }
    }
    if(MODE==4){
      #pragma unroll
      for(int j=0;j<ILP;j++) atomicAdd((unsigned*)p[j]+1, (unsigned)acc|1u);   // dependent atomic on the loaded sector
    }
  }
  if(acc==0x1234567ull) *sink=acc;
}
template<int ILP,int MODE> void run(const char* name, char* buf, size_t foot, ull n, ull* sink, int blocks_per_sm){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<ILP,MODE><<<148*blocks_per_sm,256>>>(buf,foot/32,n/8,sink);
  cudaEventRecord(a); k<ILP,MODE><<<148*blocks_per_sm,256>>>(buf,foot/32,n,sink); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,a,b);
  printf("  foot=%6.2f GiB %-26s ILP=%d blocks/SM=%d %8.3f ms  %7.2f G access/s %s\n", foot/1073741824.0, name, ILP, blocks_per_sm, ms, n/(ms*1e-3)/1e9, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  size_t maxfoot = (size_t)96<<30; ull n = 1ull<<28;
  char* buf; ull* sink; if(cudaMalloc(&buf, maxfoot)!=cudaSuccess){printf("alloc failed\n");return 1;} cudaMalloc(&sink, 8); cudaMemset(buf, 0, maxfoot);
  size_t foots[] = {(size_t)64<<20,(size_t)256<<20,(size_t)1<<30,(size_t)4<<30,(size_t)16<<30,(size_t)64<<30,(size_t)96<<30};
  for(size_t f: foots){
    run<1,0>("ld256", buf, f, n, sink, 8);
    run<4,0>("ld256", buf, f, n, sink, 8);
    run<1,1>("ld128.L2::64B", buf, f, n, sink, 8);
    run<1,2>("atomicAdd", buf, f, n, sink, 8);
    run<4,2>("atomicAdd", buf, f, n, sink, 8);
    run<1,3>("atomicCAS64", buf, f, n, sink, 8);
    run<1,4>("ld256 -> atomicAdd same", buf, f, n, sink, 8);
    cudaMemset(buf, 0, f);
  }
  printf("occupancy sweep at 16 GiB\n");
  for(int bps: {1,2,4,8}){ run<1,0>("ld256", buf, (size_t)16<<30, n, sink, bps); run<4,0>("ld256", buf, (size_t)16<<30, n, sink, bps); run<8,0>("ld256", buf, (size_t)16<<30, n, sink, bps);}
  return 0;
}
