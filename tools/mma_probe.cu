// Probe (GPU dev tool, not product): issue rate of tcgen05.mma kind::f16 (M=128, K=16) from resident shared
// memory, for K-major vs MN-major SWIZZLE_128B operands and several N.  One CTA per SM, one thread issues
// `iters` x 4 MMAs (one 64-wide K-chunk each) and commits once; clocks per MMA = elapsed / count.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o build/mma_probe tools/mma_probe.cu -lcuda
#include <cstdio>
#include <vector>
#include "../mra_gan_b200/csrc/conv_tc.cuh"
namespace mra { thread_local std::string g_last_error; std::atomic<long long> g_launch_count{0}; }
using namespace mra;
using namespace mra::tc;

__device__ __forceinline__ uint64_t desc_k_sbo(uint32_t saddr, uint32_t sbo) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}
__global__ void __launch_bounds__(128, 1) mma_probe_kernel(int n, int a_mn, int b_mn, int iters, int bshift, long long* clocks,
                                                           int ashift = 0, int asbo = 1024) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                       // 24 KB (128 x 64 bf16 + slack for shifted / strided starts)
  uint8_t* sB = smem + 24576;               // 40 KB (up to 256 x 64 bf16, + slack for shifted starts)
  uint64_t* done = reinterpret_cast<uint64_t*>(smem + 24576 + 40960);
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  for (int i = threadIdx.x; i < (24576 + 40960) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async();
  if (threadIdx.x == 0) { mbar_init(done, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc(slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t idesc = make_idesc(n, a_mn, b_mn);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB) + (uint32_t)bshift * 128u;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint64_t ad = a_mn ? desc_mnmajor_sw128(a0 + k * 2048, 8192) : desc_k_sbo(a0 + ashift * 128 + k * 32, asbo);
        const uint64_t bd = b_mn ? desc_mnmajor_sw128(b0 + k * 2048, 8192) : desc_kmajor_sw128(b0 + k * 32);
        umma_f16(tmem, ad, bd, idesc, 1u);
      }
    }
    umma_commit(done);
    mbar_wait(done, 0, nullptr, 0);
    clocks[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc(tmem, 256);
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 8);
  cudaFuncSetAttribute(mma_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
  const int iters = 4000;
  printf("%4s %5s %5s %6s %12s %12s\n", "N", "A", "B", "shift", "clk/MMA", "ideal");
  for (int n : {64, 128, 256})
    for (int mode = 0; mode < 4; ++mode)
      for (int shift : {0, 1}) {
        const int a_mn = mode & 1, b_mn = mode >> 1;
        if (shift && !(b_mn)) continue;
        mma_probe_kernel<<<148, 128, 72 * 1024>>>(n, a_mn, b_mn, iters, shift, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        std::vector<long long> c(148);
        cudaMemcpy(c.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
        double avg = 0; for (auto v : c) avg += (double)v / 148;
        printf("%4d %5s %5s %6d %12.1f %12.1f\n", n, a_mn ? "MN" : "K", b_mn ? "MN" : "K", shift, avg / (iters * 4.0), 128.0 * n * 16 / 4096 / 1.0);
      }
  printf("K-major A with row-shifted start / non-1024 group stride (B K-major, N = 256):\n");
  for (int nn : {256, 128, 64})
  for (int ashift : {0, 1, 4})
    for (int asbo : {1024, 1280}) {
      mma_probe_kernel<<<148, 128, 72 * 1024>>>(nn, 0, 0, iters, 0, d, ashift, asbo);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      std::vector<long long> c(148);
      cudaMemcpy(c.data(), d, 148 * 8, cudaMemcpyDeviceToHost);
      double avg = 0; for (auto v : c) avg += (double)v / 148;
      printf("  N %3d A shift %d rows, SBO %4d B : %.1f clk/MMA\n", nn, ashift, asbo, avg / (iters * 4.0));
    }
  return 0;
}
