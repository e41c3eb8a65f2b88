#include <cstdio>
#include <cuda_runtime.h>
__global__ void dfma_k(double* out, int iters) {
  double a[8]; for (int i=0;i<8;i++) a[i]=threadIdx.x*1e-3+i;
  double b=1.0000001, c=1e-9;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<8;i++) a[i]=fma(a[i],b,c);
  }
  double s=0; for(int i=0;i<8;i++) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
__global__ void dmma_k(double* out, int iters) {
  double d[8][2]; for(int i=0;i<8;i++){d[i][0]=0;d[i][1]=0;}
  double a=threadIdx.x*1e-3, b=1.0+threadIdx.x*1e-4;
  for (int it=0; it<iters; ++it) {
#pragma unroll
    for (int i=0;i<8;i++)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(d[i][0]), "+d"(d[i][1]) : "d"(a), "d"(b));
  }
  double s=0; for(int i=0;i<8;i++) s+=d[i][0]+d[i][1];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
}
int main(){
  double* out; cudaMalloc(&out, 148*8*1024*8);
  cudaEvent_t e0,e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int threads : {256, 512, 1024}) {
    int blocks=148*2; int iters=20000;
    dfma_k<<<blocks,threads>>>(out,100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dfma_k<<<blocks,threads>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms,e0,e1);
    double fl = 2.0*8*iters*(double)blocks*threads;
    printf("DFMA threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, fl/ms/1e9, ms);
    dmma_k<<<blocks,threads>>>(out,100); cudaDeviceSynchronize();
    cudaEventRecord(e0); dmma_k<<<blocks,threads>>>(out,iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms,e0,e1);
    double fl2 = 2.0*256*8*iters*(double)blocks*(threads/32);
    printf("DMMA threads=%d: %.2f TFLOP/s (%.3f ms)\n", threads, fl2/ms/1e9, ms);
  }
  return 0;
}
