// Micro-benchmark (developer tool): the attention exponential phase with Blackwell's packed fp32 pair instructions
// (FFMA2 / FADD2 via fma.rn.f32x2 / add.rn.f32x2) and with a share of the exponentials moved from MUFU.EX2 to a
// packed Cody-Waite + degree-3 polynomial.  PAIRS_POLY of every 8 element pairs take the polynomial.
#include <cstdio>
#include <cuda_bf16.h>
#include <cstdint>
__device__ __forceinline__ float ex2f(float x){float y; asm("ex2.approx.ftz.f32 %0, %1;":"=f"(y):"f"(x)); return y;}
__device__ __forceinline__ uint32_t pack(float a,float b){__nv_bfloat162 v=__floats2bfloat162_rn(a,b); return *reinterpret_cast<uint32_t*>(&v);}
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1){
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1), "f"(c0), "f"(c1));
}
__device__ __forceinline__ void add2(float& d0, float& d1, float a0, float a1, float b0, float b1){
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rd, ra, rb; mov.b64 {%0,%1}, rd;}"
      : "=f"(d0), "=f"(d1) : "f"(a0), "f"(a1), "f"(b0), "f"(b1));
}
// 2^x for a pair on the FMA pipe
__device__ __forceinline__ void ex2_poly2(float& y0, float& y1, float x0, float x1){
  constexpr float MAGIC = 12582912.f;
  x0 = fmaxf(x0, -125.f); x1 = fmaxf(x1, -125.f);
  float t0, t1, u0, u1, f0, f1, p0, p1;
  add2(t0, t1, x0, x1, MAGIC, MAGIC);
  add2(u0, u1, t0, t1, -MAGIC, -MAGIC);
  add2(f0, f1, x0, x1, -u0, -u1);
  fma2(p0, p1, f0, f1, 0.05520551f, 0.05520551f, 0.24261397f, 0.24261397f);
  fma2(p0, p1, p0, p1, f0, f1, 0.69325477f, 0.69325477f);
  fma2(p0, p1, p0, p1, f0, f1, 0.9999277f, 0.9999277f);
  y0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
  y1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
template<int PAIRS_POLY> __global__ void k(const float* in, uint32_t* out, long long* clk, int iters){
  float s[128];
  for(int i=0;i<128;i++) s[i]=in[(threadIdx.x*131+i)&1023];
  uint32_t acc=0; float l=0.f; float mb=in[5];
  long long t0=clock64();
  for(int it=0;it<iters;it++){
    float s0=0,s1=0,s2=0,s3=0;
#pragma unroll
    for(int i=0;i<128;i+=4){
      float x0,x1,x2,x3,p0,p1,p2,p3;
      fma2(x0,x1,s[i],s[i+1],1.4426950408889634f,1.4426950408889634f,-mb,-mb);
      fma2(x2,x3,s[i+2],s[i+3],1.4426950408889634f,1.4426950408889634f,-mb,-mb);
      const int pr = (i >> 1) & 7;                      // pair index inside a group of 8 pairs
      if (pr < PAIRS_POLY) ex2_poly2(p0,p1,x0,x1); else { p0=ex2f(x0); p1=ex2f(x1); }
      if (pr + 1 < PAIRS_POLY) ex2_poly2(p2,p3,x2,x3); else { p2=ex2f(x2); p3=ex2f(x3); }
      add2(s0,s1,s0,s1,p0,p1); add2(s2,s3,s2,s3,p2,p3);
      acc^=pack(p0,p1)^pack(p2,p3);
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
  for(int threads: {128,256}){
    for(int mode=0;mode<6;mode++){
      for(int rep=0;rep<2;rep++){
        if(mode==0) k<0><<<148,threads>>>(in,out,clk,iters);
        if(mode==1) k<1><<<148,threads>>>(in,out,clk,iters);
        if(mode==2) k<2><<<148,threads>>>(in,out,clk,iters);
        if(mode==3) k<3><<<148,threads>>>(in,out,clk,iters);
        if(mode==4) k<4><<<148,threads>>>(in,out,clk,iters);
        if(mode==5) k<5><<<148,threads>>>(in,out,clk,iters);
      }
      long long c; cudaMemcpy(&c,clk,8,cudaMemcpyDeviceToHost);
      printf("%d warps/scheduler  packed FFMA2/FADD2, %d of 8 pairs polynomial: %7.1f clk per 128-element tile per warp\n", threads/128, mode, (double)c/iters);
    }
  }
  return 0;
}
