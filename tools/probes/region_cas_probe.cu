// region_cas_probe.cu -- can inserts run at L2 speed if the table is built region by region?
// For each region size R: sweep a 1.4 GB table region by region; per region (a) optionally zero it with a
// memset right before (lines become dirty-resident in L2), (b) run a kernel that CASes 0 -> value into
// random slots of the region at load factor 0.5 (first CAS on slot 0 of a random 32-byte bucket, then the
// next slots), exactly the insert pattern.  Reports G inserts/s over the whole sweep.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__host__ __device__ inline u64 fmix64(u64 z){z^=z>>33;z*=0xFF51AFD7ED558CCDull;z^=z>>33;z*=0xC4CEB9FE1A85EC53ull;z^=z>>33;return z;}
template<int ILP> __global__ void __launch_bounds__(256) ins(u64* region,u64 nbuckets,u64 n,u64 seed,u64* fails){
  u64 tid=(u64)blockIdx.x*blockDim.x+threadIdx.x; u64 T=(u64)gridDim.x*blockDim.x; u64 f=0;
  for(u64 i=tid*ILP;i<n;i+=T*ILP){
    u64 b[ILP],v[ILP],old[ILP];
#pragma unroll
    for(int j=0;j<ILP;++j){ v[j]=fmix64(seed+i+j)|1; b[j]=__umul64hi(fmix64(v[j]),nbuckets)*4; old[j]= (i+j<n)? atomicCAS(region+b[j],0ull,v[j]) : 0; }
#pragma unroll
    for(int j=0;j<ILP;++j){ int s=1; u64 bb=b[j]; while(old[j]!=0 && i+j<n){ if(s==4){s=0; bb+=4; if(bb>=nbuckets*4) bb=0;} old[j]=atomicCAS(region+bb+s,0ull,v[j]); ++s; ++f; } }
  }
  if(f) atomicAdd(fails,f);
}
int main(){
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr,0);
  const size_t total=1434ull<<20; u64*buf,*fails; cudaMalloc(&buf,total); cudaMalloc(&fails,8);
  const u64 nslots=total/8, ninsert=nslots/2;
  cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b);
  for(int zero=0;zero<2;++zero) for(size_t rmb: {8,16,32,64,128,1434}){
    size_t R=rmb<<20; if(R>total) R=total; size_t nreg=(total+R-1)/R;
    for(int rep=0;rep<2;++rep){
      if(!zero) cudaMemset(buf,0,total);
      cudaMemset(fails,0,8); cudaDeviceSynchronize();
      cudaEventRecord(a);
      for(size_t r=0;r<nreg;++r){ size_t bytes=(r+1==nreg)? total-r*R : R; u64* reg=buf+r*R/8; u64 nb=bytes/32; u64 n=nb*2;
        if(zero) cudaMemsetAsync(reg,0,bytes);
        ins<4><<<pr.multiProcessorCount*8,256>>>(reg,nb,n,0x1234+r*7919+rep,fails); }
      cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b);
      u64 hf; cudaMemcpy(&hf,fails,8,cudaMemcpyDeviceToHost);
      if(rep) printf("zero_before=%d region=%4zu MiB (%3zu launches): %.3f ms for %.1f M inserts = %.1f G/s, extra CAS %.2f per insert\n",zero,rmb,nreg,ms,ninsert/1e6,ninsert/ms/1e6,(double)hf/ninsert);
    }
  }
  printf("%s\n",cudaGetErrorString(cudaGetLastError())); return 0;
}
