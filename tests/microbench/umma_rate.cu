// Micro-benchmark (developer tool): issue-to-completion time of back-to-back tcgen05.mma instructions for the operand
// layouts the attention kernel uses, one CTA per SM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../lidar_layout_b200/csrc
#include <cstdio>
#include "ptx.cuh"
using namespace lidm;

// mode 0: A,B K-major SWIZZLE_128B (GEMM baseline)   1: A,B K-major SWIZZLE_64B (S = Q K^T, d=32)
// mode 2: A K-major SWIZZLE_128B, B MN-major SWIZZLE_64B (O = P V)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) k(long long* out, int reps) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) { tmem_alloc(&slot, 512); tmem_relinquish(); }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  fence_proxy_async();
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t idesc = make_idesc_bf16(128, N);
    uint64_t a, b;
    if (MODE == 0) { a = make_kmajor_desc<128>(smem_u32(smem)); b = make_kmajor_desc<128>(smem_u32(smem + 16384)); }
    else if (MODE == 1) { a = make_kmajor_desc<64>(smem_u32(smem)); b = make_kmajor_desc<64>(smem_u32(smem + 16384)); }
    else { a = make_kmajor_desc<128>(smem_u32(smem)); b = make_kmajor_desc<64>(smem_u32(smem + 16384)); idesc |= (1u << 16); }
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      // 4 K-steps of 16 per repetition, alternating two accumulators like the kernel does
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const uint64_t ak = a + (MODE == 1 ? (kk & 1) * 2 : (kk & 3) * 2);
        const uint64_t bk = b + (MODE == 2 ? (uint64_t)((kk * 1024) >> 4) : (MODE == 1 ? (kk & 1) * 2 : (kk & 3) * 2));
        umma_bf16_ss(tmem + (r & 1) * 256, ak, bk, idesc, kk != 0);
      }
    }
    const long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    const long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) { tcgen05_fence_after(); tmem_dealloc(tmem, 512); }
}

template <int MODE, int N>
void run(const char* name, long long* d, int reps) {
  cudaFuncSetAttribute(k<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  k<MODE, N><<<148, 128, 65536>>>(d, reps);
  k<MODE, N><<<148, 128, 65536>>>(d, reps);
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  cudaError_t e = cudaGetLastError();
  const int n = reps * 4;
  printf("%-44s N=%3d: issue %6.1f clk/MMA, complete %6.1f clk/MMA (ideal %5.1f) %s\n", name, N, (double)h[0] / n,
         (double)h[1] / n, N / 2.0, e == cudaSuccess ? "" : cudaGetErrorString(e));
}

int main() {
  long long* d;
  cudaMalloc(&d, 64);
  const int reps = 256;
  run<0, 128>("A,B K-major SW128", d, reps);
  run<0, 256>("A,B K-major SW128", d, reps);
  run<0, 64>("A,B K-major SW128", d, reps);
  run<0, 32>("A,B K-major SW128", d, reps);
  run<1, 128>("A,B K-major SW64 (S=QK^T)", d, reps);
  run<1, 64>("A,B K-major SW64 (S=QK^T)", d, reps);
  run<2, 32>("A K-major SW128, B MN-major SW64 (O=PV)", d, reps);
  run<2, 64>("A K-major SW128, B MN-major SW64 (O=PV)", d, reps);
  return 0;
}
