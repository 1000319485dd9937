// Probe (GPU dev tool, not product): does a K-major SWIZZLE_128B UMMA descriptor accept
//   (a) a start address that is 128-B aligned but not 1024-B aligned (row shift inside a swizzle atom), and
//   (b) a stride-byte-offset (8-row group stride) that is not a multiple of 1024 B,
// when the tile was written by ONE TMA box (swizzle applied on absolute smem address bits)?
// If yes, a 3-D conv can keep ONE input halo tile in shared memory and feed all 27 taps from it.
//
// A_smem: 256 rows x 64 bf16 (one 2-D TMA box, 1024-aligned).  D[m][n] = sum_k A[row(m)][k] * B[n][k]
// with row(m) = r0 + (m / 8) * S + (m % 8), for r0 in 0..9 and S in {8, 10, 12}, with base_offset = 0
// and base_offset = (start >> 7) & 7.  Prints max |D - ref| per combination.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o gpurun_out/halo_probe tools/halo_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>

#include "../mra_gan_b200/csrc/conv_tc.cuh"

namespace mra { thread_local std::string g_last_error; std::atomic<long long> g_launch_count{0}; }
using namespace mra;
using namespace mra::tc;

__device__ __forceinline__ uint64_t desc_probe(uint32_t saddr, uint32_t sbo_bytes, uint32_t base_off) {
  return (uint64_t)((saddr >> 4) & 0x3FFF) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) |
         ((uint64_t)(base_off & 7) << 49) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int r0,
             int S, int use_base_off) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // 256 rows x 128 B = 32 KB
  uint8_t* sB = smem + 256 * 128;        // 64 rows x 128 B = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* slot = reinterpret_cast<uint32_t*>(done + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(slot, 64);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0, nullptr, 0);
    tc_fence_after();
    const uint32_t a0 = smem_u32(sA) + (uint32_t)r0 * 128u;
    const uint32_t bo = use_base_off ? ((a0 >> 7) & 7u) : 0u;
    const uint32_t idesc = make_idesc(64, 0, 0);
    for (int k = 0; k < 4; ++k)
      umma_f16(tmem, desc_probe(a0 + k * 32, (uint32_t)S * 128u, bo), desc_kmajor_sw128(smem_u32(sB) + k * 32), idesc,
               (uint32_t)(k != 0));
    umma_commit(done);
  }
  mbar_wait(done, 0, nullptr, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
    tmem_wait_ld();
    for (int i = 0; i < 32; ++i) out[(warp * 32 + lane) * 64 + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 64);
}

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

int main() {
  const int RA = 256, RB = 64, K = 64;
  std::vector<float> A(RA * K), B(RB * K);
  std::vector<bf16> Ah(RA * K), Bh(RB * K);
  srand(1);
  for (int i = 0; i < RA * K; ++i) { A[i] = bf((rand() % 2001 - 1000) / 1000.f); Ah[i] = __float2bfloat16_rn(A[i]); }
  for (int i = 0; i < RB * K; ++i) { B[i] = bf((rand() % 2001 - 1000) / 1000.f); Bh[i] = __float2bfloat16_rn(B[i]); }
  bf16 *dA, *dB; float* dO;
  cudaMalloc(&dA, RA * K * 2); cudaMalloc(&dB, RB * K * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, Ah.data(), RA * K * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, Bh.data(), RB * K * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  if (make_weight_map(&tmA, dA, RA, K, RA) || make_weight_map(&tmB, dB, RB, K, RB)) { printf("tmap failed: %s\n", g_last_error.c_str()); return 1; }
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  std::vector<float> O(128 * 64);
  int bad = 0;
  for (int ubo = 0; ubo < 2; ++ubo)
    for (int S : {8, 10, 12})
      for (int r0 = 0; r0 < 10; ++r0) {
        if (r0 + 15 * S + 8 > RA) continue;
        cudaMemset(dO, 0, 128 * 64 * 4);
        probe_kernel<<<1, 128, 48 * 1024, 0>>>(tmA, tmB, dO, r0, S, ubo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cuda error %s\n", cudaGetErrorString(e)); return 2; }
        cudaMemcpy(O.data(), dO, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        double maxerr = 0;
        for (int m = 0; m < 128; ++m) {
          const int row = r0 + (m / 8) * S + (m % 8);
          for (int n = 0; n < 64; ++n) {
            double acc = 0;
            for (int k = 0; k < K; ++k) acc += (double)A[row * K + k] * B[n * K + k];
            maxerr = fmax(maxerr, fabs(acc - O[m * 64 + n]));
          }
        }
        printf("base_off=%d S=%2d r0=%d  max|err| = %.4g  %s\n", ubo, S, r0, maxerr, maxerr < 1e-2 ? "OK" : "MISMATCH");
        if (maxerr >= 1e-2) ++bad;
      }
  printf("mismatching combinations: %d\n", bad);
  return 0;
}
