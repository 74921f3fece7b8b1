#include <cstdio>
#include "panel16.cuh"
namespace wfo { thread_local char g_err[512]; }
using namespace wfo;
template<int PW>
__global__ void k(double* out, long long* cyc, int n, int reps){
  extern __shared__ __align__(16) double sm[];
  __shared__ P16Shared sh;
  int ld = (n + 7) & ~7; if ((ld & 15) == 0) ld += 8;
  const int LD = ld;
  double* A = sm;
  const int warp = threadIdx.x>>5, lane = threadIdx.x&31;
  for (int i=threadIdx.x;i<((n+7)&~7)*LD;i+=blockDim.x){ unsigned h=i*2654435761u; A[i] = ((h>>8)&0xffff)/65536.0 - 0.5; }
  __syncthreads();
  double det=1.0;
  long long t0=clock64();
  for(int r=0;r<reps;r++){
    if (warp < PW) factor_panel16<PW>(A, LD, n, 0, 16, &sh, warp, lane, det);
    __syncthreads();
  }
  long long t1=clock64();
  if(threadIdx.x==0){ cyc[0]=(t1-t0)/reps; out[0]=det; cyc[1]=sh.aff_dst[30]; cyc[2]=sh.aff_dst[31]; }
}
template<int PW> void run(int n, double* out, long long* cyc){
  int ld = (n + 7) & ~7; if ((ld & 15) == 0) ld += 8;
  size_t smem = (size_t)((n+7)&~7)*ld*8;
  cudaFuncSetAttribute(k<PW>,cudaFuncAttributeMaxDynamicSharedMemorySize,(int)smem);
  k<PW><<<1,PW*32 > 64 ? PW*32 : 64,smem>>>(out,cyc,n,50); cudaDeviceSynchronize();
  printf("PW=%d n=%d: %lld cyc/panel16; column loop %lld (%lld per column), post %lld %s\n",PW,n,cyc[0],cyc[1],cyc[1]/16,cyc[2],cudaGetErrorString(cudaGetLastError()));
}
int main(){
  double* out; long long* cyc; cudaMallocManaged(&out,8); cudaMallocManaged(&cyc,64);
  run<1>(32,out,cyc); run<2>(64,out,cyc); run<3>(96,out,cyc); run<4>(100,out,cyc); run<4>(128,out,cyc);
  return 0;
}
