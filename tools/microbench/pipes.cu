// pipes.cu — issue/pipe microbenchmarks that decide the pair-force kernel design:
// scalar FFMA vs packed FFMA2 (fma.rn.f32x2) throughput, FSETP and LDS-broadcast cost beside FFMA.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};":"=l"(r):"f"(a),"f"(b)); return r; }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c){ u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;":"=l"(d):"l"(a),"l"(b),"l"(c)); return d; }
__device__ __forceinline__ u64 fadd2(u64 a, u64 b){ u64 d; asm volatile("add.rn.f32x2 %0, %1, %2;":"=l"(d):"l"(a),"l"(b)); return d; }

__global__ void k_ffma(float* out, float a, float b){
  float x[16];
  #pragma unroll
  for(int i=0;i<16;i++) x[i]=threadIdx.x*1e-3f+i;
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<16;i++) x[i]=fmaf(x[i],a,b);
  }
  float s=0; for(int i=0;i<16;i++) s+=x[i]; if(s==1.2345f) out[0]=s;
}
__global__ void k_ffma2(float* out, float a, float b){
  u64 x[16]; u64 A=pk(a,a), B=pk(b,b);
  #pragma unroll
  for(int i=0;i<16;i++) x[i]=pk(threadIdx.x*1e-3f+i, i);
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<16;i++) x[i]=ffma2(x[i],A,B);
  }
  u64 s=0; for(int i=0;i<16;i++) s^=x[i]; if(s==12345) out[0]=1;
}
__global__ void k_fadd2(float* out, float a){
  u64 x[16]; u64 A=pk(a,a);
  #pragma unroll
  for(int i=0;i<16;i++) x[i]=pk(threadIdx.x*1e-3f+i, i);
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<16;i++) x[i]=fadd2(x[i],A);
  }
  u64 s=0; for(int i=0;i<16;i++) s^=x[i]; if(s==12345) out[0]=1;
}
// 12 FFMA + 4 FSETP/SEL-like integer ops per inner group: do ALU ops co-issue for free?
__global__ void k_mix_alu(float* out, float a, float b, int m){
  float x[12]; int y[4];
  #pragma unroll
  for(int i=0;i<12;i++) x[i]=threadIdx.x*1e-3f+i;
  #pragma unroll
  for(int i=0;i<4;i++) y[i]=threadIdx.x+i;
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<12;i++) x[i]=fmaf(x[i],a,b);
    #pragma unroll
    for(int i=0;i<4;i++) y[i]=__funnelshift_l(__float_as_int(x[i]), y[i], 1);
  }
  float s=0; for(int i=0;i<12;i++) s+=x[i]; int t=0; for(int i=0;i<4;i++) t^=y[i];
  if(s==1.2345f && t==m) out[0]=s;
}
// 12 FFMA2 (= 24 lane-FMAs) + 4 SHF
__global__ void k_mix2_alu(float* out, float a, float b, int m){
  u64 x[12]; int y[4]; u64 A=pk(a,a), B=pk(b,b);
  #pragma unroll
  for(int i=0;i<12;i++) x[i]=pk(threadIdx.x*1e-3f+i,i);
  #pragma unroll
  for(int i=0;i<4;i++) y[i]=threadIdx.x+i;
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<12;i++) x[i]=ffma2(x[i],A,B);
    #pragma unroll
    for(int i=0;i<4;i++) y[i]=__funnelshift_l((int)(x[i]>>32), y[i], 1);
  }
  u64 s=0; for(int i=0;i<12;i++) s^=x[i]; int t=0; for(int i=0;i<4;i++) t^=y[i];
  if(s==12345 && t==m) out[0]=1;
}
// FFMA + LDS.128 broadcast: 16 FFMA per 1 LDS.128 (uniform address)
__global__ void k_mix_lds(float* out, float a, int stride){
  __shared__ float4 sm[256];
  sm[threadIdx.x]=make_float4(threadIdx.x,1,2,3); __syncthreads();
  float x[16];
  #pragma unroll
  for(int i=0;i<16;i++) x[i]=threadIdx.x*1e-3f+i;
  int j=0;
  for(int it=0;it<ITERS;it++){
    float4 q=sm[j&255]; j+=stride;
    #pragma unroll
    for(int i=0;i<16;i+=4){ x[i]=fmaf(x[i],a,q.x); x[i+1]=fmaf(x[i+1],a,q.y); x[i+2]=fmaf(x[i+2],a,q.z); x[i+3]=fmaf(x[i+3],a,q.w);}
  }
  float s=0; for(int i=0;i<16;i++) s+=x[i]; if(s==1.2345f) out[0]=s;
}
// MUFU beside FFMA: 14 FFMA + 2 MUFU
__global__ void k_mix_mufu(float* out, float a, float b){
  float x[14]; float y[2];
  #pragma unroll
  for(int i=0;i<14;i++) x[i]=threadIdx.x*1e-3f+i;
  y[0]=threadIdx.x+1.f; y[1]=threadIdx.x+2.f;
  for(int it=0;it<ITERS;it++){
    #pragma unroll
    for(int i=0;i<14;i++) x[i]=fmaf(x[i],a,b);
    asm volatile("rsqrt.approx.ftz.f32 %0, %0;":"+f"(y[0]));
    asm volatile("ex2.approx.ftz.f32 %0, %0;":"+f"(y[1]));
  }
  float s=0; for(int i=0;i<14;i++) s+=x[i]; s+=y[0]+y[1]; if(s==1.2345f) out[0]=s;
}
template<typename F> float timeit(F f){ cudaEvent_t a,b; cudaEventCreate(&a); cudaEventCreate(&b); f(); f(); cudaEventRecord(a); for(int i=0;i<5;i++) f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms,a,b); return ms/5; }
int main(){
  float* out; cudaMalloc(&out,64); cudaDeviceProp p; cudaGetDeviceProperties(&p,0);
  int blocks=p.multiProcessorCount*8, thr=256; double lanes=(double)blocks*thr*ITERS;
  printf("device %s SMs %d\n", p.name, p.multiProcessorCount);
  float t;
  t=timeit([&]{k_ffma<<<blocks,thr>>>(out,1.0001f,0.5f);});  printf("FFMA   : %.3f ms  %.2f T lane-FMA/s\n", t, lanes*16/t*1e-9);
  t=timeit([&]{k_ffma2<<<blocks,thr>>>(out,1.0001f,0.5f);}); printf("FFMA2  : %.3f ms  %.2f T lane-FMA/s\n", t, lanes*32/t*1e-9);
  t=timeit([&]{k_fadd2<<<blocks,thr>>>(out,0.5f);});         printf("FADD2  : %.3f ms  %.2f T lane-ADD/s\n", t, lanes*32/t*1e-9);
  t=timeit([&]{k_mix_alu<<<blocks,thr>>>(out,1.0001f,0.5f,7);});  printf("12FFMA+4SHF : %.3f ms  %.2f T lane-FMA/s  (%.2f T instr-lanes/s)\n", t, lanes*12/t*1e-9, lanes*16/t*1e-9);
  t=timeit([&]{k_mix2_alu<<<blocks,thr>>>(out,1.0001f,0.5f,7);}); printf("12FFMA2+4SHF: %.3f ms  %.2f T lane-FMA/s\n", t, lanes*24/t*1e-9);
  t=timeit([&]{k_mix_lds<<<blocks,thr>>>(out,1.0001f,1);});  printf("16FFMA+1LDS128 bcast: %.3f ms  %.2f T lane-FMA/s\n", t, lanes*16/t*1e-9);
  t=timeit([&]{k_mix_mufu<<<blocks,thr>>>(out,1.0001f,0.5f);}); printf("14FFMA+2MUFU: %.3f ms  %.2f T lane-FMA/s\n", t, lanes*14/t*1e-9);
  return 0;
}
