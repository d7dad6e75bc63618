// gups.cu - random 8-byte access micro-benchmark: loads/s and load+CAS/s versus footprint.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a tools/gups.cu -o tools/gups
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint64_t mix(uint64_t z){ z+=0x9E3779B97F4A7C15ull; z=(z^(z>>30))*0xBF58476D1CE4E5B9ull; z=(z^(z>>27))*0x94D049BB133111EBull; return z^(z>>31);}
template<int MODE, int ILP>
__global__ void __launch_bounds__(256) k(uint64_t* a, uint64_t n, int iters, uint64_t* sink){
  uint64_t t = blockIdx.x*(uint64_t)blockDim.x+threadIdx.x, acc=0;
  for (int it=0; it<iters; ++it){
    uint64_t v[ILP]; uint64_t idx[ILP];
    #pragma unroll
    for (int j=0;j<ILP;++j){ idx[j] = mix(t*1315423911ull + it*ILP + j) % n; v[j] = __ldcg((const unsigned long long*)&a[idx[j]]); }
    #pragma unroll
    for (int j=0;j<ILP;++j){
      if (MODE==1) { uint64_t p = atomicCAS((unsigned long long*)&a[idx[j]], (unsigned long long)v[j], (unsigned long long)(v[j]+1)); acc += p; }
      else if (MODE==2) { atomicAdd((unsigned long long*)&a[idx[j]], 1ull); acc += v[j]; }
      else if (MODE==3) { a[idx[j]] = v[j] + 1; acc += v[j]; }
      else acc += v[j];
    }
  }
  if (acc==0x1234567) *sink = acc;
}
template<int MODE,int ILP> void run(uint64_t* a, uint64_t n, const char* name){
  uint64_t* sink; cudaMalloc(&sink,8);
  int blocks=148*8, iters=256/ILP*4;
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE,ILP><<<blocks,256>>>(a,n,iters/4,sink);
  cudaEventRecord(e0); k<MODE,ILP><<<blocks,256>>>(a,n,iters,sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1);
  double ops=(double)blocks*256*iters*ILP;
  printf("  %-22s ILP=%d  %.2f G/s  (%.1f ms)\n", name, ILP, ops/ms/1e6, ms);
  cudaFree(sink);
}
int main(){
  for (double gb : {0.008, 0.03125, 0.0625, 0.25, 1.0, 8.0, 32.0, 64.0, 120.0}){
    uint64_t n=(uint64_t)(gb*(1ull<<30)/8); uint64_t* a;
    if (cudaMalloc(&a,n*8)!=cudaSuccess){ printf("alloc %.0f GB failed\n",gb); continue; }
    cudaMemset(a,0,n*8);
    printf("footprint %.4g GiB\n",gb);
    run<0,1>(a,n,"load");
    run<0,8>(a,n,"load");
    run<1,8>(a,n,"load+CAS");
    run<2,8>(a,n,"load+atomicAdd");
    run<3,8>(a,n,"load+store");
    cudaFree(a);
  }
  return 0;
}
