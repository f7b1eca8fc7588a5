// Micro-benchmark (developer tool): issue cost (clocks per warp instruction, one scheduler) of the instructions the
// attention softmax is built from, including Blackwell's packed fp32 pairs.  8 independent chains per thread.
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#define REP8(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)
template<int OP> __global__ void k(const float* in, float* out, long long* clk, int iters){
  float a[16]; for(int i=0;i<16;i++) a[i]=in[(threadIdx.x+i)&63];
  const float c0=in[3], c1=in[4];
  long long t0=clock64();
  for(int it=0;it<iters;it++){
#pragma unroll
    for(int u=0;u<4;u++){
#pragma unroll
      for(int i=0;i<8;i++){
        float& x=a[2*i]; float& y=a[2*i+1];
        if(OP==0){ asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(x):"f"(c0),"f"(c1)); asm volatile("fma.rn.f32 %0,%0,%1,%2;":"+f"(y):"f"(c0),"f"(c1)); }                         // 2 FFMA
        if(OP==1){ asm volatile("{.reg .b64 ra,rb,rc; mov.b64 ra,{%0,%1}; mov.b64 rb,{%2,%2}; mov.b64 rc,{%3,%3}; fma.rn.f32x2 ra,ra,rb,rc; mov.b64 {%0,%1},ra;}":"+f"(x),"+f"(y):"f"(c0),"f"(c1)); }  // 1 FFMA2
        if(OP==2){ asm volatile("add.rn.f32 %0,%0,%1;":"+f"(x):"f"(c0)); asm volatile("add.rn.f32 %0,%0,%1;":"+f"(y):"f"(c1)); }                                                // 2 FADD
        if(OP==3){ asm volatile("{.reg .b64 ra,rb; mov.b64 ra,{%0,%1}; mov.b64 rb,{%2,%3}; add.rn.f32x2 ra,ra,rb; mov.b64 {%0,%1},ra;}":"+f"(x),"+f"(y):"f"(c0),"f"(c1)); }  // 1 FADD2
        if(OP==4){ asm volatile("max.f32 %0,%0,%1;":"+f"(x):"f"(c0)); asm volatile("max.f32 %0,%0,%1;":"+f"(y):"f"(c1)); }                                // 2 FMNMX
        if(OP==5){ asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(x)); asm volatile("ex2.approx.ftz.f32 %0,%0;":"+f"(y)); }  // 2 MUFU
        if(OP==6){ uint32_t xi=__float_as_uint(x), yi=__float_as_uint(y); asm volatile("{.reg .b32 t; shl.b32 t,%1,23; add.s32 %0,%0,t;}":"+r"(xi):"r"(__float_as_uint(c0))); asm volatile("{.reg .b32 t; shl.b32 t,%1,23; add.s32 %0,%0,t;}":"+r"(yi):"r"(__float_as_uint(c1))); x=__uint_as_float(xi); y=__uint_as_float(yi); } // 2 shift-add
        if(OP==7){ uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0,%1,%2;":"=r"(r):"f"(y),"f"(x)); x=__uint_as_float(r); }  // 1 F2FP
        if(OP==8){ asm volatile("max.f32 %0,%0,%1,%2;":"+f"(x):"f"(y),"f"(c0)); asm volatile("max.f32 %0,%0,%1,%2;":"+f"(y):"f"(x),"f"(c1)); } // 2 FMNMX3
      }
    }
  }
  long long t1=clock64();
  float s=0; for(int i=0;i<16;i++) s+=a[i];
  out[blockIdx.x*blockDim.x+threadIdx.x]=s;
  if(threadIdx.x==0&&blockIdx.x==0) clk[0]=t1-t0;
}
int main(){
  float* in; float* out; long long* clk; cudaMalloc(&in,256); cudaMalloc(&out,148*1024*4); cudaMalloc(&clk,8);
  float h[64]; for(int i=0;i<64;i++) h[i]=0.5f+0.001f*i; cudaMemcpy(in,h,256,cudaMemcpyHostToDevice);
  const int iters=2000;
  const char* names[9]={"2x FFMA","1x FFMA2","2x FADD","1x FADD2","2x FMNMX","2x MUFU.EX2","2x shift-add","1x F2FP.PACK","2x FMNMX3"};
  for(int threads: {128,256}){
    for(int op=0;op<9;op++){
      for(int rep=0;rep<2;rep++){
        switch(op){
          case 0:k<0><<<148,threads>>>(in,out,clk,iters);break; case 1:k<1><<<148,threads>>>(in,out,clk,iters);break;
          case 2:k<2><<<148,threads>>>(in,out,clk,iters);break; case 3:k<3><<<148,threads>>>(in,out,clk,iters);break;
          case 4:k<4><<<148,threads>>>(in,out,clk,iters);break; case 5:k<5><<<148,threads>>>(in,out,clk,iters);break;
          case 6:k<6><<<148,threads>>>(in,out,clk,iters);break; case 7:k<7><<<148,threads>>>(in,out,clk,iters);break;
          case 8:k<8><<<148,threads>>>(in,out,clk,iters);break;
        }
      }
      long long c; cudaMemcpy(&c,clk,8,cudaMemcpyDeviceToHost);
      printf("%d warps/scheduler  %-14s per element pair: %6.2f clk (scheduler-wide: %6.2f)\n", threads/128, names[op], (double)c/iters/32, (double)c/iters/32/(threads/128));
    }
  }
  return 0;
}
