// scripts/probe_vmm.cu — measurement helper (not product code): where is the TLB-reach cliff for
// random sector reads, and does it depend on how the memory was allocated (cudaMalloc vs cuMemCreate
// with minimum / recommended granularity / one big physical allocation)?
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
typedef unsigned long long ull;
__device__ __forceinline__ ull mix(ull z){ z += 0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
__global__ void __launch_bounds__(256) k(const char* buf, ull nsec, ull n, ull* sink){
  ull acc=0; ull tid=blockIdx.x*(ull)blockDim.x+threadIdx.x, nth=(ull)gridDim.x*blockDim.x;
  for(ull a=tid;a<n;a+=nth){ const char* p = buf + (mix(a)%nsec)*32; ull c0,c1,c2,c3;
    asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(c0),"=l"(c1),"=l"(c2),"=l"(c3):"l"(p):"memory"); acc^=c0^c1^c2^c3; }
  if(acc==0x1234567ull) *sink=acc;
}
static double run(const char* buf, size_t foot, ull* sink){
  ull n=1ull<<27; cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  k<<<148*8,256>>>(buf,foot/32,n/8,sink); cudaEventRecord(a); k<<<148*8,256>>>(buf,foot/32,n,sink); cudaEventRecord(b); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms,a,b); return n/(ms*1e-3)/1e9;
}
#define CU(x) do{CUresult r_=(x); if(r_!=CUDA_SUCCESS){const char* s_; cuGetErrorString(r_,&s_); printf("%s failed: %s\n",#x,s_); return 1;}}while(0)
int main(){
  cudaFree(0); ull* sink; cudaMalloc(&sink,8);
  size_t foots[] = {(size_t)48<<30,(size_t)64<<30,(size_t)72<<30,(size_t)80<<30,(size_t)96<<30,(size_t)128<<30};
  { char* buf; size_t total=(size_t)128<<30; if(cudaMalloc(&buf,total)==cudaSuccess){ cudaMemset(buf,0,total);
      for(size_t f: foots) printf("cudaMalloc(one 128 GiB block)      foot=%4zu GiB  %6.2f G/s\n", f>>30, run(buf,f,sink)); cudaFree(buf);} else printf("cudaMalloc 128GiB failed\n"); }
  CUmemAllocationProp prop = {}; prop.type = CU_MEM_ALLOCATION_TYPE_PINNED; prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE; prop.location.id = 0;
  size_t gmin=0, grec=0; CU(cuMemGetAllocationGranularity(&gmin,&prop,CU_MEM_ALLOC_GRANULARITY_MINIMUM)); CU(cuMemGetAllocationGranularity(&grec,&prop,CU_MEM_ALLOC_GRANULARITY_RECOMMENDED));
  printf("granularity min=%zu recommended=%zu\n", gmin, grec);
  size_t total=(size_t)128<<30;
  for(size_t chunk : {(size_t)0, (size_t)2<<20, (size_t)512<<20, (size_t)8<<30}){
    size_t csz = chunk? chunk : total; if(csz%grec) csz = (csz/grec+1)*grec;
    CUdeviceptr va; CU(cuMemAddressReserve(&va,total,(size_t)1<<30,0,0));
    std::vector<CUmemGenericAllocationHandle> hs; size_t mapped=0; bool ok=true;
    while(mapped<total){ CUmemGenericAllocationHandle h; size_t sz = (total-mapped<csz)? total-mapped: csz; if(cuMemCreate(&h,sz,&prop,0)!=CUDA_SUCCESS){ok=false;break;} if(cuMemMap(va+mapped,sz,0,h,0)!=CUDA_SUCCESS){ok=false;break;} hs.push_back(h); mapped+=sz; if(chunk==(size_t)2<<20 && mapped>=((size_t)128<<30)) break; }
    if(!ok){ printf("VMM chunk=%zu: create/map failed at %zu GiB\n", csz, mapped>>30); }
    CUmemAccessDesc ad = {}; ad.location = prop.location; ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE; CU(cuMemSetAccess(va,mapped,&ad,1));
    cudaMemset((void*)va,0,mapped);
    for(size_t f: foots) if(f<=mapped) printf("VMM chunk=%6zu MiB (%zu handles)   foot=%4zu GiB  %6.2f G/s\n", csz>>20, hs.size(), f>>30, run((const char*)va,f,sink));
    cuMemUnmap(va,mapped); for(auto h: hs) cuMemRelease(h); cuMemAddressFree(va,total);
  }
  return 0;
}
