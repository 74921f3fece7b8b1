// scratch: isolate factor_subpanel latency
#include <cstdio>
#include "../pywfo_b200/csrc/det_lu_blocked.cuh"
namespace wfo { thread_local char g_err[512]; }
using namespace wfo;
template<int PW>
__global__ void k(double* out, long long* cyc, int n, int reps){
  extern __shared__ __align__(16) double sm[];
  __shared__ BlkShared sh;
  const int LD = blk_ld_h(n);
  double* A = sm;
  const int warp = threadIdx.x>>5, lane = threadIdx.x&31;
  for (int i=threadIdx.x;i<blk_rows(n)*LD;i+=blockDim.x){ unsigned h=i*2654435761u; A[i] = ((h>>8)&0xffff)/65536.0 - 0.5; }
  __syncthreads();
  double det=1.0; bool singular=false;
  long long t0=clock64();
  for(int r=0;r<reps;r++){
    if (warp < PW) factor_subpanel<PW>(A, LD, n, 0, 8, &sh, warp, lane, det, singular);
    __syncthreads();
  }
  long long t1=clock64();
  if(threadIdx.x==0){ cyc[0]=(t1-t0)/reps; out[0]=det; }
}
int main(){
  double* out; long long* cyc; cudaMallocManaged(&out,8); cudaMallocManaged(&cyc,8);
  int n; size_t smem;
  n=32; smem=blk_smem_bytes(n); cudaFuncSetAttribute(k<1>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  k<1><<<1,64,smem>>>(out,cyc,n,50); cudaDeviceSynchronize(); printf("PW=1 n=32: %lld cyc/subpanel (%lld per column) err=%s\n",cyc[0],cyc[0]/8,cudaGetErrorString(cudaGetLastError()));
  n=64; smem=blk_smem_bytes(n); cudaFuncSetAttribute(k<2>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  k<2><<<1,128,smem>>>(out,cyc,n,50); cudaDeviceSynchronize(); printf("PW=2 n=64: %lld cyc/subpanel (%lld per column)\n",cyc[0],cyc[0]/8);
  n=100; smem=blk_smem_bytes(n); cudaFuncSetAttribute(k<4>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  k<4><<<1,256,smem>>>(out,cyc,n,50); cudaDeviceSynchronize(); printf("PW=4 n=100: %lld cyc/subpanel (%lld per column)\n",cyc[0],cyc[0]/8);
  return 0;
}
