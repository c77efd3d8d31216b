// scripts/probe_gran.cu — measurement helper (not product code): how many DRAM sectors does one
// random, 32-byte-aligned access cost, by load width and cudaLimitMaxL2FetchGranularity?
// Run under: ncu --metrics dram__sectors_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,gpu__time_duration.sum
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__device__ __forceinline__ ull mix(ull z){ z += 0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
template<int MODE> __global__ void __launch_bounds__(256) k(const char* buf, ull nsec, ull n, ull* sink){
  ull acc=0; ull tid=blockIdx.x*(ull)blockDim.x+threadIdx.x, nth=(ull)gridDim.x*blockDim.x;
  for(ull a=tid;a<n;a+=nth){
    const char* p = buf + (mix(a)%nsec)*32;
    if(MODE==0){ ull c0,c1,c2,c3; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); acc^=c0^c1^c2^c3; }
    if(MODE==1){ ull c0,c1,c2,c3; asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(c0),"=l"(c1):"l"(p):"memory"); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(c2),"=l"(c3):"l"(p+16):"memory"); acc^=c0^c1^c2^c3; }
    if(MODE==2){ ull c0; asm volatile("ld.global.cg.u64 %0, [%1];":"=l"(c0):"l"(p):"memory"); acc^=c0; }
    if(MODE==3){ ull c0,c1,c2,c3; asm volatile("ld.global.ca.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); acc^=c0^c1^c2^c3; }
    if(MODE==4){ ull c0,c1,c2,c3; asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); acc^=c0^c1^c2^c3; }
    if(MODE==5){ ull c0,c1,c2,c3; asm volatile("ld.global.cg.L2::64B.v2.u64 {%0,%1}, [%2];":"=l"(c0),"=l"(c1):"l"(p):"memory"); asm volatile("ld.global.cg.v2.u64 {%0,%1}, [%2];":"=l"(c2),"=l"(c3):"l"(p+16):"memory"); acc^=c0^c1^c2^c3; }
    if(MODE==6){ atomicAdd((unsigned*)p, 1u); }
    if(MODE==7){ ull c0,c1,c2,c3; asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); acc^=c0^c1^c2^c3; atomicAdd((unsigned*)p + (c0&7), 1u); }
  }
  if(acc==0x1234567ull) *sink=acc;
}
template<int MODE> void run(const char* name, char* buf, ull nsec, ull n, ull* sink){
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<148*8,256>>>(buf,nsec,n/8,sink);
  cudaEventRecord(a); k<MODE><<<148*8,256>>>(buf,nsec,n,sink); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,a,b);
  printf("  %-34s %8.3f ms  %7.2f G access/s  err=%s\n", name, ms, n/(ms*1e-3)/1e9, cudaGetErrorString(cudaGetLastError()));
}
int main(){
  size_t foot = (size_t)16<<30; ull n = 1ull<<28;
  char* buf; ull* sink; cudaMalloc(&buf, foot); cudaMalloc(&sink, 8); cudaMemset(buf, 0, foot);
  for(int gran : {0, 32, 64, 128}){
    if(gran){ cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran); size_t v=0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity set %d -> %zu (%s)\n", gran, v, cudaGetErrorString(e)); }
    else { size_t v=0; cudaDeviceGetLimit(&v, cudaLimitMaxL2FetchGranularity); printf("L2 fetch granularity default = %zu\n", v); }
    run<0>("ld.cg.v4.u64 (256-bit)", buf, foot/32, n, sink);
    run<1>("2 x ld.cg.v2.u64 (128-bit)", buf, foot/32, n, sink);
    run<2>("ld.cg.u64 (64-bit)", buf, foot/32, n, sink);
    run<3>("ld.ca.v4.u64", buf, foot/32, n, sink);
    run<4>("ld.nc.L1::no_allocate.v4.u64", buf, foot/32, n, sink);
    run<5>("ld.cg.L2::64B.v2 + v2", buf, foot/32, n, sink);
    run<6>("atomicAdd u32", buf, foot/32, n, sink);
    run<7>("ld 256 + atomicAdd same sector", buf, foot/32, n, sink);
  }
  return 0;
}
