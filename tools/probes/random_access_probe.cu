// random_access_probe.cu -- measures the B200's random-access ceilings that bound the hash path.
// A tool (not linked into the product): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o probe random_access_probe.cu
//   * random 32-byte reads (one 256-bit load each) with different L2 fetch-granularity settings / PTX hints
//   * random 8-byte reads, random 64-bit CAS, read-then-CAS (the insert pattern)
// Output: one line per experiment with G accesses/s; ncu on this binary gives DRAM bytes per access.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__host__ __device__ inline u64 fmix64(u64 z){z^=z>>33;z*=0xFF51AFD7ED558CCDull;z^=z>>33;z*=0xC4CEB9FE1A85EC53ull;z^=z>>33;return z;}
template<int MODE> __device__ __forceinline__ void ld256(const void*p,u64(&q)[4]){
  if(MODE==0) asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==4) asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==5) asm volatile("ld.global.cv.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
  if(MODE==6) asm volatile("ld.global.L1::evict_first.L2::evict_first.v4.u64 {%0,%1,%2,%3}, [%4];":"=l"(q[0]),"=l"(q[1]),"=l"(q[2]),"=l"(q[3]):"l"(p));
}
template<int MODE,int ILP> __global__ void __launch_bounds__(256) rd32(const u64*buf,u64 nsec,u64 per,u64*sink){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x, st=fmix64(tid+77), x=0;
  for(u64 i=0;i<per;i+=ILP){ u64 q[ILP][4];
#pragma unroll
    for(int j=0;j<ILP;++j){ st=st*6364136223846793005ull+1442695040888963407ull; ld256<MODE>(buf+4*__umul64hi(fmix64(st),nsec),q[j]); }
#pragma unroll
    for(int j=0;j<ILP;++j) x^=q[j][0]^q[j][1]^q[j][2]^q[j][3]; }
  if(x==0x1234567) *sink=x;
}
template<int ILP> __global__ void __launch_bounds__(256) rd8(const u64*buf,u64 nw,u64 per,u64*sink){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x, st=fmix64(tid+77), x=0;
  for(u64 i=0;i<per;i+=ILP){ u64 q[ILP];
#pragma unroll
    for(int j=0;j<ILP;++j){ st=st*6364136223846793005ull+1442695040888963407ull; q[j]=__ldg(buf+__umul64hi(fmix64(st),nw)); }
#pragma unroll
    for(int j=0;j<ILP;++j) x^=q[j]; }
  if(x==0x1234567) *sink=x;
}
// dependent chain: the next address depends on the loaded value (like a walker); MLP only across threads
__global__ void __launch_bounds__(256) chase32(const u64*buf,u64 nsec,u64 per,u64*sink){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x, st=fmix64(tid+99);
  for(u64 i=0;i<per;++i){ u64 q[4]; ld256<0>(buf+4*__umul64hi(fmix64(st),nsec),q); st=st*6364136223846793005ull+q[0]+q[3]+1; }
  if(st==0x1234567) *sink=st;
}
template<int ILP> __global__ void __launch_bounds__(256) cas8(u64*buf,u64 nw,u64 per,u64*sink){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x, st=fmix64(tid+77), x=0;
  for(u64 i=0;i<per;i+=ILP){ u64 q[ILP];
#pragma unroll
    for(int j=0;j<ILP;++j){ st=st*6364136223846793005ull+1442695040888963407ull; q[j]=atomicCAS(buf+__umul64hi(fmix64(st),nw),0ull,st|1); }
#pragma unroll
    for(int j=0;j<ILP;++j) x^=q[j]; }
  if(x==0x1234567) *sink=x;
}
template<int ILP> __global__ void __launch_bounds__(256) rdcas(u64*buf,u64 nsec,u64 per,u64*sink){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x, st=fmix64(tid+77), x=0;
  for(u64 i=0;i<per;i+=ILP){ u64 q[ILP][4]; u64 a[ILP];
#pragma unroll
    for(int j=0;j<ILP;++j){ st=st*6364136223846793005ull+1442695040888963407ull; a[j]=4*__umul64hi(fmix64(st),nsec); ld256<4>(buf+a[j],q[j]); }
#pragma unroll
    for(int j=0;j<ILP;++j){ int s=q[j][0]==0?0:q[j][1]==0?1:q[j][2]==0?2:3; x^=atomicCAS(buf+a[j]+s,0ull,st|1); } }
  if(x==0x1234567) *sink=x;
}
template<class F> double timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms; }
int main(int argc,char**argv){
  size_t fp = (argc>1? strtoull(argv[1],0,10):1434ull)<<20;   // MiB, default = chr14 k=19 table size
  int occ = argc>2? atoi(argv[2]):8;
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr,0);
  size_t lim=0; cudaDeviceGetLimit(&lim,cudaLimitMaxL2FetchGranularity); printf("device %s SMs %d L2 %d MiB default L2 fetch granularity %zu\n",pr.name,pr.multiProcessorCount,pr.l2CacheSize>>20,lim);
  u64*buf,*sink; cudaMalloc(&buf,fp); cudaMalloc(&sink,8); cudaMemset(buf,1,fp);
  unsigned blocks=pr.multiProcessorCount*occ; u64 threads=(u64)blocks*256; u64 per=(1ull<<28)/threads/8*8;
  u64 nsec=fp/32,nw=fp/8; double tot=(double)per*threads;
  printf("footprint %zu MiB, %u blocks x 256, %llu accesses/thread, %.0f M accesses per run\n",fp>>20,blocks,per,tot/1e6);
  for(int g: {0,32,64,128}){ if(g){ cudaError_t e=cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,g); cudaDeviceGetLimit(&lim,cudaLimitMaxL2FetchGranularity); printf("set limit %d -> %s, now %zu\n",g,cudaGetErrorString(e),lim);}
    double ms=timeit([&]{rd32<0,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 nc ILP4 limit=%d: %.2f G/s (%.2f TB/s of 32B)\n",g,tot/ms/1e6,tot*32/ms/1e9);} 
  cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity,32);
  double ms;
  ms=timeit([&]{rd32<1,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 L2::64B  ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<2,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 L2::128B ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<3,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 L2::256B ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<4,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 cg ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<5,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 cv ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<6,4><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 evict_first ILP4: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<0,1><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 nc ILP1: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{rd32<0,8><<<blocks,256>>>(buf,nsec,per,sink);}); printf("rd32 nc ILP8: %.2f G/s\n",tot/ms/1e6);
  ms=timeit([&]{chase32<<<blocks,256>>>(buf,nsec,per,sink);}); printf("chase32 (dependent, %d thr/SM): %.2f G/s\n",occ*256,tot/ms/1e6);
  ms=timeit([&]{rd8<4><<<blocks,256>>>(buf,nw,per,sink);}); printf("rd8 ILP4: %.2f G/s\n",tot/ms/1e6);
  cudaMemset(buf,0,fp);
  ms=timeit([&]{rdcas<4><<<blocks,256>>>(buf,nsec,per/2,sink);}); printf("read32+CAS8 ILP4 (insert pattern): %.2f G/s\n",tot/2/ms/1e6);
  cudaMemset(buf,0,fp);
  ms=timeit([&]{cas8<4><<<blocks,256>>>(buf,nw,per/2,sink);}); printf("CAS8 ILP4: %.2f G/s\n",tot/2/ms/1e6);
  // L2-resident footprint for comparison
  u64 nsec2=(32ull<<20)/32; 
  ms=timeit([&]{rd32<0,4><<<blocks,256>>>(buf,nsec2,per,sink);}); printf("rd32 nc ILP4 32MiB footprint (L2 resident): %.2f G/s\n",tot/ms/1e6);
  cudaMemset(buf,0,32<<20);
  ms=timeit([&]{rdcas<4><<<blocks,256>>>(buf,nsec2,per/2,sink);}); printf("read32+CAS8 32MiB footprint: %.2f G/s\n",tot/2/ms/1e6);
  cudaMemset(buf,0,32<<20);
  ms=timeit([&]{cas8<4><<<blocks,256>>>(buf,(32ull<<20)/8,per/2,sink);}); printf("CAS8 32MiB footprint: %.2f G/s\n",tot/2/ms/1e6);
  printf("last error: %s\n",cudaGetErrorString(cudaGetLastError()));
  return 0;
}
