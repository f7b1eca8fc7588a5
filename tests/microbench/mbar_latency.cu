// Micro-benchmark (developer tool): cost of testing an mbarrier whose phase is ALREADY complete, per flavour, and of the
// tcgen05 fences that follow such a wait in the kernels.  One warp, dependent chain of N waits.
#include <cstdio>
#include <cstdint>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int MODE> __device__ __forceinline__ uint32_t wait1(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  if (MODE == 0) asm volatile("{.reg .pred P; mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0,1,0,P;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (MODE == 1) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2; selp.u32 %0,1,0,P;}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  if (MODE == 2) asm volatile("{.reg .pred P; mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2, %3; selp.u32 %0,1,0,P;}" : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
  return ok;
}
template <int MODE> __global__ void k(long long* out, int iters) {
  __shared__ uint64_t bar;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar))); }
  __syncthreads();
  if (threadIdx.x == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");   // phase 0 completes
  __syncthreads();
  uint32_t acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t ok = wait1<MODE>(smem_u32(&bar), acc & 0u);      // parity 0 (complete); depends on the previous result
    acc += ok;
    if (MODE == 3) {}
  }
  long long t1 = clock64();
  long long t2 = clock64();
  for (int i = 0; i < iters; ++i) { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
  long long t3 = clock64();
  for (int i = 0; i < iters; ++i) { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
  long long t4 = clock64();
  if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t3 - t2; out[2] = t4 - t3; out[3] = acc; }
}
int main() {
  long long* d; cudaMalloc(&d, 64); long long h[4]; const int iters = 1000;
  const char* names[3] = {"test_wait", "try_wait", "try_wait + suspend hint"};
  for (int m = 0; m < 3; ++m) {
    for (int rep = 0; rep < 2; ++rep) { if (m == 0) k<0><<<1, 32>>>(d, iters); if (m == 1) k<1><<<1, 32>>>(d, iters); if (m == 2) k<2><<<1, 32>>>(d, iters); }
    cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
    printf("%-26s %6.1f clk per dependent wait (all true: %lld)   tcgen05.fence::after %5.1f clk   fence.proxy.async %5.1f clk\n", names[m], (double)h[0] / iters, h[3], (double)h[1] / iters, (double)h[2] / iters);
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
