// Micro-benchmark (developer tool): how fast can one warp run the attention kernel's exponential phase
// (FFMA + MUFU.EX2 + FADD + F2FP per element, 128 elements per thread) in isolation, with 1 or 2 warps per scheduler?
#include <cstdio>
#include <cuda_bf16.h>
#include <cstdint>
__device__ __forceinline__ float ex2f(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ uint32_t pack(float a,float b){__nv_bfloat162 v=__floats2bfloat162_rn(a,b); return *reinterpret_cast<uint32_t*>(&v);}
template<int MODE> __global__ void k(const float* in, uint32_t* out, long long* clk, int iters){
  float s[128];
  for(int i=0;i<128;i++) s[i]=in[(threadIdx.x*131+i)&1023];
  uint32_t acc=0; float l=0.f; float mb=in[5];
  long long t0=clock64();
  for(int it=0;it<iters;it++){
    float s0=0,s1=0,s2=0,s3=0;
#pragma unroll
    for(int i=0;i<128;i+=4){
      float p0,p1,p2,p3;
      if(MODE==0){ p0=ex2f(s[i]); p1=ex2f(s[i+1]); p2=ex2f(s[i+2]); p3=ex2f(s[i+3]); acc^=__float_as_uint(p0)^__float_as_uint(p1)^__float_as_uint(p2)^__float_as_uint(p3);}
      else {
        p0=ex2f(fmaf(s[i],1.4426950408889634f,-mb)); p1=ex2f(fmaf(s[i+1],1.4426950408889634f,-mb));
        p2=ex2f(fmaf(s[i+2],1.4426950408889634f,-mb)); p3=ex2f(fmaf(s[i+3],1.4426950408889634f,-mb));
        if(MODE>=2){ s0+=p0; s1+=p1; s2+=p2; s3+=p3; }
        if(MODE>=3){ acc^=pack(p0,p1)^pack(p2,p3); } else acc^=__float_as_uint(p0)^__float_as_uint(p1)^__float_as_uint(p2)^__float_as_uint(p3);
      }
    }
    l+=(s0+s1)+(s2+s3); mb+=1e-6f;
  }
  long long t1=clock64();
  out[blockIdx.x*blockDim.x+threadIdx.x]=acc^__float_as_uint(l);
  if(threadIdx.x==0&&blockIdx.x==0) clk[0]=t1-t0;
}
int main(){
  float* in; uint32_t* out; long long* clk; cudaMalloc(&in,4096); cudaMalloc(&out,148*1024*4); cudaMalloc(&clk,8);
  float h[1024]; for(int i=0;i<1024;i++) h[i]=-0.01f*i; cudaMemcpy(in,h,4096,cudaMemcpyHostToDevice);
  const int iters=200;
  const char* names[4]={"MUFU only","FFMA+MUFU","FFMA+MUFU+FADD","FFMA+MUFU+FADD+F2FP (kernel loop)"};
  for(int threads: {128,256,512}){
    for(int mode=0;mode<4;mode++){
      for(int rep=0;rep<2;rep++){
        if(mode==0) k<0><<<148,threads>>>(in,out,clk,iters);
        if(mode==1) k<1><<<148,threads>>>(in,out,clk,iters);
        if(mode==2) k<2><<<148,threads>>>(in,out,clk,iters);
        if(mode==3) k<3><<<148,threads>>>(in,out,clk,iters);
      }
      long long c; cudaMemcpy(&c,clk,8,cudaMemcpyDeviceToHost);
      printf("%d warps/scheduler  %-36s: %7.1f clk per 128-element tile per warp (%.2f clk/element)\n", threads/128, names[mode], (double)c/iters, (double)c/iters/128);
    }
  }
  return 0;
}
