// micro-benchmarks for FP64 pipes on sm_100a (scratch; not part of the product)
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do{cudaError_t e=(x); if(e!=cudaSuccess){printf("ERR %s line %d\n",cudaGetErrorString(e),__LINE__); return 1;}}while(0)
__device__ __forceinline__ void dmma(double&c0,double&c1,double a,double b){
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n":"+d"(c0),"+d"(c1):"d"(a),"d"(b));}

// latency chains, one warp
__global__ void lat_kernel(long long* out, double* sink){
  double x=1.0+threadIdx.x*1e-9, y=1e-9, z=0.5;
  long long t0,t1;
  // DFMA chain
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) x=fma(x,z,y);
  }
  t1=clock64(); if(threadIdx.x==0) out[0]=t1-t0;
  // DMMA chain
  double c0=x,c1=y;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) dmma(c0,c1,z,z);
  }
  t1=clock64(); if(threadIdx.x==0) out[1]=t1-t0;
  // SHFL chain (64-bit = 2 shuffles)
  double s=c0+c1;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) s=__shfl_xor_sync(0xffffffffu,s,1+(j&15));
  }
  t1=clock64(); if(threadIdx.x==0) out[2]=t1-t0;
  // REDUX max u32 chain
  unsigned r=(unsigned)__double2hiint(s)+threadIdx.x;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) r=__reduce_max_sync(0xffffffffu,r+threadIdx.x);
  }
  t1=clock64(); if(threadIdx.x==0) out[3]=t1-t0;
  // ballot+ffs chain
  unsigned b=r;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) b=__ffs(__ballot_sync(0xffffffffu,(b+threadIdx.x)&1))+b;
  }
  t1=clock64(); if(threadIdx.x==0) out[4]=t1-t0;
  // DP reciprocal chain
  double q=s+2.0;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) q=1.0/(q+1.5);
  }
  t1=clock64(); if(threadIdx.x==0) out[5]=t1-t0;
  // 32-bit shfl chain
  unsigned u=b;
  t0=clock64();
  #pragma unroll 1
  for(int i=0;i<64;i++){
    #pragma unroll
    for(int j=0;j<16;j++) u=__shfl_sync(0xffffffffu,u,(j*7)&31)+1;
  }
  t1=clock64(); if(threadIdx.x==0) out[6]=t1-t0;
  sink[threadIdx.x]=x+c0+c1+s+r+b+q+u;
}
// syncthreads latency
__global__ void bar_kernel(long long* out){
  long long t0=clock64();
  #pragma unroll 1
  for(int i=0;i<1024;i++) __syncthreads();
  long long t1=clock64();
  if(threadIdx.x==0) out[0]=t1-t0;
}
// throughput: mode 0 dfma only, 1 dmma only, 2 both interleaved (per warp: ILP accumulators)
template<int MODE>
__global__ void __launch_bounds__(1024) thr_kernel(int iters,double* sink){
  double a[8], c[8][2];
  double x=1.0+1e-9*threadIdx.x,y=1e-9;
  #pragma unroll
  for(int i=0;i<8;i++){a[i]=i+threadIdx.x; c[i][0]=i; c[i][1]=-i;}
  for(int it=0;it<iters;it++){
    #pragma unroll
    for(int i=0;i<8;i++){
      if(MODE!=1) { a[i]=fma(a[i],x,y); if(MODE==0) a[i]=fma(a[i],y,x);} 
      if(MODE!=0) dmma(c[i][0],c[i][1],x,y);
      if(MODE==3) { a[i]=fma(a[i],y,x); a[i]=fma(a[i],x,y); a[i]=fma(a[i],y,x);} 
    }
  }
  double s=0; 
  #pragma unroll
  for(int i=0;i<8;i++) s+=a[i]+c[i][0]+c[i][1];
  if(s==1234.5) sink[0]=s;
}
template<int MODE>
float run_thr(int blocks,int threads,int iters,double*sink){
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  thr_kernel<MODE><<<blocks,threads>>>(iters,sink);
  cudaEventRecord(e0); thr_kernel<MODE><<<blocks,threads>>>(iters,sink); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms,e0,e1); return ms;
}
int main(){
  long long* out; double* sink; CK(cudaMallocManaged(&out,64*8)); CK(cudaMalloc(&sink,8192*8));
  lat_kernel<<<1,32>>>(out,sink); CK(cudaDeviceSynchronize());
  lat_kernel<<<1,32>>>(out,sink); CK(cudaDeviceSynchronize());
  const char* names[]={"DFMA dep","DMMA dep","SHFL64 dep","REDUX.MAX dep(+IADD)","ballot+ffs+add dep","DRCP(1/x)+DADD dep","SHFL32+IADD dep"};
  for(int i=0;i<7;i++) printf("%-24s %.2f cyc/op\n",names[i],out[i]/1024.0);
  for(int th=64; th<=1024; th*=2){ bar_kernel<<<1,th>>>(out); CK(cudaDeviceSynchronize()); printf("bar.sync %4d threads: %.1f cyc\n",th,out[0]/1024.0);} 
  int sms=148; int iters=2048;
  for(int wps=1; wps<=8; wps*=2){
    int threads=wps*4*32; 
    float m0=run_thr<0>(sms,threads,iters,sink), m1=run_thr<1>(sms,threads,iters,sink), m2=run_thr<2>(sms,threads,iters,sink), m3=run_thr<3>(sms,threads,iters,sink);
    double f0=2.0*16*iters*(double)sms*threads, f1=8.0*16*iters*(double)sms*threads;  // dfma flops: 16 fma/iter/thread; dmma: 8 dmma/iter/warp*512flops/32
    double fd=8*512.0/32*iters*(double)sms*threads;
    printf("warps/SMSP=%d: DFMA %.2f TF/s | DMMA %.2f TF/s | mix(1 dfma:1 dmma) t=%.3f ms vs dfma-only-equiv %.3f + dmma %.3f -> total %.2f TF/s | mix(4:1) %.2f TF/s\n", wps,
      f0/m0/1e9, fd/m1/1e9, m2, m0/2, m1, (2.0*8*iters*(double)sms*threads+fd)/m2/1e9, (2.0*32*iters*(double)sms*threads+fd)/m3/1e9);
  }
  return 0;
}
